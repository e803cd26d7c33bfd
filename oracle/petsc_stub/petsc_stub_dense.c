/* petsc_stub_dense.c -- second half of the container-only PETSc API stand-in (TEST INFRASTRUCTURE ONLY, see petsc_stub.h).
 *
 * What the reference's pc_sorgibbs.c, pc_chols.c (dense branch), stats.c, iact.c and the MATLRC branches of mc_sor.c /
 * pc_mcgibbs.c call beyond the sweep engine: small dense column-major matrices, MatMult on SeqAIJ / dense / MATLRC operators,
 * an exact dense solve behind KSP / MatLUFactor / MatCholeskyFactor, PETSc's MatSOR_SeqAIJ semantics (restated from its public
 * documentation), reference BLAS / LAPACK routines (dpotrf lower, dtrsv lower) and the three FFTW calls.
 * Toy bodies behind the public signatures; nothing of PETSc, LAPACK or FFTW is copied.
 */
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "fftw3.h"
#include "petsc_stub.h"

static int is_type(Mat A, const char *t) { return A && A->type && strcmp(A->type, t) == 0; }
static int is_dense(Mat A) { return A && A->d != NULL; }

/* ---- dense matrices ------------------------------------------------------------------------------------------------ */
PetscErrorCode MatCreateSeqDense(MPI_Comm comm, PetscInt m, PetscInt n, PetscScalar *data, Mat *A)
{
  Mat B        = calloc(1, sizeof(*B));
  B->hdr.comm  = comm;
  B->hdr.refct = 1;
  B->type      = MATSEQDENSE;
  B->m = B->M = m;
  B->n = B->N = n;
  B->d = calloc((size_t)(m > 0 ? m : 1) * (size_t)(n > 0 ? n : 1), sizeof(double));
  if (data) memcpy(B->d, data, sizeof(double) * (size_t)m * (size_t)n);
  *A = B;
  return 0;
}
PetscErrorCode MatCreateDense(MPI_Comm comm, PetscInt m, PetscInt n, PetscInt M, PetscInt N, PetscScalar *data, Mat *A)
{
  return MatCreateSeqDense(comm, m >= 0 ? m : M, n >= 0 ? n : N, data, A);
}
PetscErrorCode MatDenseGetArray(Mat A, PetscScalar **a)
{
  PetscCheck(is_dense(A), PETSC_COMM_SELF, PETSC_ERR_SUP, "not a dense matrix");
  *a = A->d;
  return 0;
}
PetscErrorCode MatDenseRestoreArray(Mat A, PetscScalar **a) { (void)A; *a = NULL; return 0; }
PetscErrorCode MatDenseGetArrayRead(Mat A, const PetscScalar **a)
{
  PetscCheck(is_dense(A), PETSC_COMM_SELF, PETSC_ERR_SUP, "not a dense matrix");
  *a = A->d;
  return 0;
}
PetscErrorCode MatDenseRestoreArrayRead(Mat A, const PetscScalar **a) { (void)A; *a = NULL; return 0; }
static PetscErrorCode column_vec(Mat A, PetscInt c, Vec *v)
{
  PetscCheck(is_dense(A) && c >= 0 && c < A->n, PETSC_COMM_SELF, PETSC_ERR_SUP, "bad dense column");
  return VecStubCreate(A->hdr.comm, A->m, A->M, 0, A->d + (size_t)c * A->m, v); /* aliases the column */
}
PetscErrorCode MatDenseGetColumnVecRead(Mat A, PetscInt c, Vec *v) { return column_vec(A, c, v); }
PetscErrorCode MatDenseRestoreColumnVecRead(Mat A, PetscInt c, Vec *v) { (void)A; (void)c; return VecDestroy(v); }
PetscErrorCode MatDenseGetColumnVecWrite(Mat A, PetscInt c, Vec *v) { return column_vec(A, c, v); }
PetscErrorCode MatDenseRestoreColumnVecWrite(Mat A, PetscInt c, Vec *v) { (void)A; (void)c; return VecDestroy(v); }

PetscErrorCode MatZeroEntries(Mat A)
{
  if (is_dense(A)) memset(A->d, 0, sizeof(double) * (size_t)A->m * A->n);
  else if (A->a) memset(A->a, 0, sizeof(double) * (size_t)A->i[A->m]);
  return 0;
}
PetscErrorCode MatDuplicate(Mat A, MatDuplicateOption o, Mat *B)
{
  if (is_dense(A)) {
    PetscCall(MatCreateSeqDense(A->hdr.comm, A->m, A->n, o == MAT_COPY_VALUES ? A->d : NULL, B));
    return 0;
  }
  PetscCheck(is_type(A, MATSEQAIJ), PETSC_COMM_SELF, PETSC_ERR_SUP, "MatDuplicate: seqaij or dense only");
  PetscCall(MatStubCreateSeqAIJ(A->hdr.comm, A->m, A->n, A->i, A->j, A->a, B));
  if (o != MAT_COPY_VALUES) memset((*B)->a, 0, sizeof(double) * (size_t)A->i[A->m]);
  return 0;
}
static double entry(Mat A, PetscInt r, PetscInt c)
{
  if (is_dense(A)) return A->d[(size_t)r + (size_t)c * A->m];
  for (PetscInt k = A->i[r]; k < A->i[r + 1]; ++k)
    if (A->j[k] == c) return A->a[k];
  return 0.0;
}
PetscErrorCode MatConvert(Mat A, MatType t, MatReuse r, Mat *B)
{
  (void)r;
  if (strcmp(t, MATSEQDENSE) == 0 || strcmp(t, MATDENSE) == 0) {
    PetscCall(MatCreateSeqDense(A->hdr.comm, A->m, A->n, NULL, B));
    for (PetscInt i = 0; i < A->m; ++i)
      for (PetscInt j = 0; j < A->n; ++j) (*B)->d[(size_t)i + (size_t)j * A->m] = entry(A, i, j);
    return 0;
  }
  if (strcmp(t, MATAIJ) == 0 || strcmp(t, MATSEQAIJ) == 0) { /* dense -> aij keeps every entry (explicit zeros included) */
    PetscInt *ii = malloc(sizeof(PetscInt) * (size_t)(A->m + 1)), *jj = malloc(sizeof(PetscInt) * (size_t)A->m * A->n + 1);
    double   *aa = malloc(sizeof(double) * (size_t)A->m * A->n + 8);
    PetscInt  nz = 0;
    for (PetscInt i = 0; i < A->m; ++i) {
      ii[i] = nz;
      for (PetscInt j = 0; j < A->n; ++j) {
        jj[nz] = j;
        aa[nz] = entry(A, i, j);
        ++nz;
      }
    }
    ii[A->m] = nz;
    PetscCall(MatStubCreateSeqAIJ(A->hdr.comm, A->m, A->n, ii, jj, aa, B));
    free(ii); free(jj); free(aa);
    return 0;
  }
  return PetscStubError(PETSC_ERR_SUP, __FILE__, __LINE__, "MatConvert to %s is not emulated", t);
}
PetscErrorCode MatAXPY(Mat Y, PetscScalar a, Mat X, MatStructure s)
{
  (void)s;
  if (is_dense(Y)) {
    for (PetscInt j = 0; j < Y->n; ++j)
      for (PetscInt i = 0; i < Y->m; ++i) Y->d[(size_t)i + (size_t)j * Y->m] += a * entry(X, i, j);
    return 0;
  }
  /* aij += a * X: rebuild the row lists over the union pattern */
  Mat D;
  PetscCall(MatConvert(Y, MATSEQDENSE, MAT_INITIAL_MATRIX, &D));
  for (PetscInt j = 0; j < Y->n; ++j)
    for (PetscInt i = 0; i < Y->m; ++i) D->d[(size_t)i + (size_t)j * Y->m] += a * entry(X, i, j);
  PetscInt nz = 0;
  for (PetscInt i = 0; i < Y->m; ++i)
    for (PetscInt j = 0; j < Y->n; ++j)
      if (D->d[(size_t)i + (size_t)j * Y->m] != 0.0 || entry(Y, i, j) != 0.0) ++nz;
  free(Y->j); free(Y->a);
  Y->j = malloc(sizeof(PetscInt) * (size_t)(nz + 1));
  Y->a = malloc(sizeof(double) * (size_t)(nz + 1));
  nz   = 0;
  for (PetscInt i = 0; i < Y->m; ++i) {
    Y->i[i] = nz;
    for (PetscInt j = 0; j < Y->n; ++j) {
      const double v = D->d[(size_t)i + (size_t)j * Y->m];
      if (v != 0.0) { Y->j[nz] = j; Y->a[nz] = v; ++nz; }
    }
  }
  Y->i[Y->m] = nz;
  return MatDestroy(&D);
}
PetscErrorCode MatNorm(Mat A, NormType t, PetscReal *nrm)
{
  PetscCheck(t == NORM_FROBENIUS && is_dense(A), PETSC_COMM_SELF, PETSC_ERR_SUP, "MatNorm: Frobenius norm of a dense matrix only");
  double s = 0.0;
  for (size_t q = 0; q < (size_t)A->m * A->n; ++q) s += A->d[q] * A->d[q];
  *nrm = sqrt(s);
  return 0;
}
PetscErrorCode MatShift(Mat A, PetscScalar s)
{
  if (is_dense(A)) {
    for (PetscInt i = 0; i < (A->m < A->n ? A->m : A->n); ++i) A->d[(size_t)i + (size_t)i * A->m] += s;
    return 0;
  }
  for (PetscInt r = 0; r < A->m; ++r)
    for (PetscInt k = A->i[r]; k < A->i[r + 1]; ++k)
      if (A->j[k] == r) A->a[k] += s;
  return 0;
}
PetscErrorCode MatDiagonalSet(Mat A, Vec d, InsertMode m)
{
  PetscCheck(is_dense(A), PETSC_COMM_SELF, PETSC_ERR_SUP, "MatDiagonalSet: dense only");
  for (PetscInt i = 0; i < A->m; ++i) {
    double *p = &A->d[(size_t)i + (size_t)i * A->m];
    *p        = m == ADD_VALUES ? *p + d->a[i] : d->a[i];
  }
  return 0;
}
PetscErrorCode MatDiagonalScale(Mat A, Vec l, Vec r)
{
  PetscCheck(is_type(A, MATSEQAIJ), PETSC_COMM_SELF, PETSC_ERR_SUP, "MatDiagonalScale: seqaij only");
  for (PetscInt i = 0; i < A->m; ++i)
    for (PetscInt k = A->i[i]; k < A->i[i + 1]; ++k) A->a[k] *= (l ? l->a[i] : 1.0) * (r ? r->a[A->j[k]] : 1.0);
  return 0;
}
PetscErrorCode MatSetOption(Mat A, MatOption o, PetscBool v) { (void)A; (void)o; (void)v; return 0; }
PetscErrorCode MatGetInfo(Mat A, MatInfoType t, MatInfo *info)
{
  (void)t;
  info->nz_used = info->nz_allocated = is_dense(A) ? (double)A->m * A->n : (A->i ? (double)A->i[A->m] : 0.0);
  info->memory  = 0;
  return 0;
}

/* ---- products ------------------------------------------------------------------------------------------------------ */
static void mult_add(Mat A, const double *x, double *y) /* y += A x */
{
  if (is_dense(A)) {
    for (PetscInt j = 0; j < A->n; ++j)
      for (PetscInt i = 0; i < A->m; ++i) y[i] += A->d[(size_t)i + (size_t)j * A->m] * x[j];
  } else {
    for (PetscInt i = 0; i < A->m; ++i) {
      double s = 0.0;
      for (PetscInt k = A->i[i]; k < A->i[i + 1]; ++k) s += A->a[k] * x[A->j[k]];
      y[i] += s;
    }
  }
}
static void mult_transpose_add(Mat A, const double *x, double *y) /* y += A^T x */
{
  if (is_dense(A)) {
    for (PetscInt j = 0; j < A->n; ++j) {
      double s = 0.0;
      for (PetscInt i = 0; i < A->m; ++i) s += A->d[(size_t)i + (size_t)j * A->m] * x[i];
      y[j] += s;
    }
  } else {
    for (PetscInt i = 0; i < A->m; ++i)
      for (PetscInt k = A->i[i]; k < A->i[i + 1]; ++k) y[A->j[k]] += A->a[k] * x[i];
  }
}
PetscErrorCode MatMultAdd(Mat A, Vec x, Vec y, Vec z)
{
  double *tmp = calloc((size_t)(A->m > 0 ? A->m : 1), sizeof(double));
  if (is_type(A, MATLRC)) { /* (A + U diag(c) U^T) x */
    Mat     U = A->lrc_U;
    double *w = calloc((size_t)U->n, sizeof(double));
    mult_add(A->lrc_A, x->a, tmp);
    mult_transpose_add(U, x->a, w);
    for (PetscInt j = 0; j < U->n; ++j) w[j] *= A->lrc_c->a[j];
    mult_add(U, w, tmp);
    free(w);
  } else mult_add(A, x->a, tmp);
  for (PetscInt i = 0; i < A->m; ++i) z->a[i] = y->a[i] + tmp[i];
  free(tmp);
  return 0;
}
PetscErrorCode MatMult(Mat A, Vec x, Vec y)
{
  Vec zero;
  PetscCall(VecDuplicate(y, &zero));
  PetscCall(VecZeroEntries(zero));
  PetscCall(MatMultAdd(A, x, zero, y));
  return VecDestroy(&zero);
}
PetscErrorCode MatMultTranspose(Mat A, Vec x, Vec y)
{
  for (PetscInt j = 0; j < A->n; ++j) y->a[j] = 0.0;
  mult_transpose_add(A, x->a, y->a);
  return 0;
}
static PetscErrorCode dense_product(Mat A, int ta, Mat B, int tb, Mat *C) /* C = op(A) op(B), dense result */
{
  const PetscInt m = ta ? A->n : A->m, kk = ta ? A->m : A->n, n = tb ? B->m : B->n;
  PetscCheck((tb ? B->n : B->m) == kk, PETSC_COMM_SELF, PETSC_ERR_ARG_WRONG, "product dimensions do not match");
  PetscCall(MatCreateSeqDense(A->hdr.comm, m, n, NULL, C));
  for (PetscInt j = 0; j < n; ++j)
    for (PetscInt i = 0; i < m; ++i) {
      double s = 0.0;
      for (PetscInt k = 0; k < kk; ++k) s += (ta ? entry(A, k, i) : entry(A, i, k)) * (tb ? entry(B, j, k) : entry(B, k, j));
      (*C)->d[(size_t)i + (size_t)j * m] = s;
    }
  return 0;
}
PetscErrorCode MatTransposeMatMult(Mat A, Mat B, MatReuse r, PetscReal f, Mat *C) { (void)r; (void)f; return dense_product(A, 1, B, 0, C); }
PetscErrorCode MatMatMult(Mat A, Mat B, MatReuse r, PetscReal f, Mat *C) { (void)r; (void)f; return dense_product(A, 0, B, 0, C); }
PetscErrorCode MatMatTransposeMult(Mat A, Mat B, MatReuse r, PetscReal f, Mat *C) { (void)r; (void)f; return dense_product(A, 0, B, 1, C); }

/* ---- MATLRC -------------------------------------------------------------------------------------------------------- */
PetscErrorCode MatCreateLRC(Mat A, Mat U, Vec c, Mat V, Mat *N)
{
  PetscCheck(V == NULL || V == U, PETSC_COMM_SELF, PETSC_ERR_SUP, "MatCreateLRC: V = U only");
  Mat B        = calloc(1, sizeof(*B));
  B->hdr.comm  = A->hdr.comm;
  B->hdr.refct = 1;
  B->type      = MATLRC;
  B->m = B->M = A->m;
  B->n = B->N = A->n;
  B->lrc_A = A; B->lrc_U = U; B->lrc_c = c;
  *N = B;
  return 0;
}
PetscErrorCode MatLRCGetMats(Mat A, Mat *base, Mat *U, Vec *c, Mat *V)
{
  PetscCheck(is_type(A, MATLRC), PETSC_COMM_SELF, PETSC_ERR_SUP, "not a MATLRC matrix");
  if (base) *base = A->lrc_A;
  if (U) *U = A->lrc_U;
  if (c) *c = A->lrc_c;
  if (V) *V = A->lrc_U;
  return 0;
}

/* ---- dense factorisations (the "factor" matrix F keeps the factored copy) ----------------------------------------------- */
static int lu_factor(int n, double *a, int *piv)
{
  for (int k = 0; k < n; ++k) {
    int p = k;
    for (int i = k + 1; i < n; ++i)
      if (fabs(a[i + (size_t)k * n]) > fabs(a[p + (size_t)k * n])) p = i;
    piv[k] = p;
    if (a[p + (size_t)k * n] == 0.0) return k + 1;
    if (p != k)
      for (int j = 0; j < n; ++j) { double t = a[k + (size_t)j * n]; a[k + (size_t)j * n] = a[p + (size_t)j * n]; a[p + (size_t)j * n] = t; }
    for (int i = k + 1; i < n; ++i) {
      const double l = a[i + (size_t)k * n] /= a[k + (size_t)k * n];
      for (int j = k + 1; j < n; ++j) a[i + (size_t)j * n] -= l * a[k + (size_t)j * n];
    }
  }
  return 0;
}
static void lu_solve(int n, const double *a, const int *piv, double *x)
{
  for (int k = 0; k < n; ++k) {
    if (piv[k] != k) { double t = x[k]; x[k] = x[piv[k]]; x[piv[k]] = t; }
    for (int i = k + 1; i < n; ++i) x[i] -= a[i + (size_t)k * n] * x[k];
  }
  for (int k = n - 1; k >= 0; --k) {
    x[k] /= a[k + (size_t)k * n];
    for (int i = 0; i < k; ++i) x[i] -= a[i + (size_t)k * n] * x[k];
  }
}
PetscErrorCode MatFactorInfoInitialize(MatFactorInfo *info) { info->fill = 1; info->dtcol = 0; return 0; }
PetscErrorCode MatGetFactor(Mat A, MatSolverType st, MatFactorType ft, Mat *F)
{
  (void)st;
  Mat B        = calloc(1, sizeof(*B));
  B->hdr.comm  = A->hdr.comm;
  B->hdr.refct = 1;
  B->type      = "factor";
  B->m = B->M = A->m;
  B->n = B->N = A->n;
  B->fac_kind  = ft == MAT_FACTOR_LU ? 1 : 2;
  *F = B;
  return 0;
}
PetscErrorCode MatGetOrdering(Mat A, MatOrderingType t, IS *r, IS *c)
{
  (void)t; /* every ordering is the natural one here: the dense factorisations do not need a fill-reducing permutation */
  PetscCall(ISCreateStride(A->hdr.comm, A->m, 0, 1, r));
  PetscCall(ISCreateStride(A->hdr.comm, A->m, 0, 1, c));
  return 0;
}
PetscErrorCode MatLUFactorSymbolic(Mat F, Mat A, IS r, IS c, const MatFactorInfo *info) { (void)F; (void)A; (void)r; (void)c; (void)info; return 0; }
static PetscErrorCode load_dense(Mat F, Mat A)
{
  const int n = A->m;
  free(F->fac);
  F->fac = malloc(sizeof(double) * (size_t)n * n + 8);
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) F->fac[i + (size_t)j * n] = entry(A, i, j);
  return 0;
}
PetscErrorCode MatLUFactorNumeric(Mat F, Mat A, const MatFactorInfo *info)
{
  (void)info;
  PetscCall(load_dense(F, A));
  free(F->piv);
  F->piv = malloc(sizeof(int) * (size_t)(A->m + 1));
  PetscCheck(lu_factor(A->m, F->fac, F->piv) == 0, PETSC_COMM_SELF, PETSC_ERR_LIB, "singular matrix");
  return 0;
}
PetscErrorCode MatCholeskyFactorSymbolic(Mat F, Mat A, IS perm, const MatFactorInfo *info) { (void)F; (void)A; (void)perm; (void)info; return 0; }
PetscErrorCode MatCholeskyFactorNumeric(Mat F, Mat A, const MatFactorInfo *info)
{
  (void)info;
  PetscCall(load_dense(F, A));
  PetscBLASInt n = A->m, linfo = 0;
  LAPACKpotrf_("L", &n, F->fac, &n, &linfo);
  PetscCheck(linfo == 0, PETSC_COMM_SELF, PETSC_ERR_MAT_CH_ZRPVT, "Cholesky factorisation failed at pivot %d", linfo);
  return 0;
}
PetscErrorCode MatMatSolve(Mat F, Mat B, Mat X)
{
  PetscCheck(F->fac_kind == 1 && F->fac && is_dense(B) && is_dense(X), PETSC_COMM_SELF, PETSC_ERR_SUP, "MatMatSolve: dense LU only");
  if (X != B) memcpy(X->d, B->d, sizeof(double) * (size_t)B->m * B->n);
  for (PetscInt j = 0; j < X->n; ++j) lu_solve(F->m, F->fac, F->piv, X->d + (size_t)j * X->m);
  return 0;
}
PetscErrorCode MatForwardSolve(Mat F, Vec b, Vec x) /* L x = b */
{
  PetscCheck(F->fac_kind == 2 && F->fac, PETSC_COMM_SELF, PETSC_ERR_SUP, "MatForwardSolve: dense Cholesky only");
  PetscBLASInt n = F->m, one = 1;
  if (x != b) memcpy(x->a, b->a, sizeof(double) * (size_t)n);
  BLAStrsv_("L", "N", "N", &n, F->fac, &n, x->a, &one);
  return 0;
}
PetscErrorCode MatBackwardSolve(Mat F, Vec b, Vec x) /* L^T x = b */
{
  PetscCheck(F->fac_kind == 2 && F->fac, PETSC_COMM_SELF, PETSC_ERR_SUP, "MatBackwardSolve: dense Cholesky only");
  PetscBLASInt n = F->m, one = 1;
  if (x != b) memcpy(x->a, b->a, sizeof(double) * (size_t)n);
  BLAStrsv_("L", "T", "N", &n, F->fac, &n, x->a, &one);
  return 0;
}

/* ---- KSP: an exact dense solve with the operator ------------------------------------------------------------------------- */
struct _p_KSP {
  struct _p_PetscObject hdr;
  Mat                   A;
};
PetscErrorCode KSPCreate(MPI_Comm comm, KSP *ksp)
{
  *ksp            = calloc(1, sizeof(**ksp));
  (*ksp)->hdr.comm = comm;
  return 0;
}
PetscErrorCode KSPSetOperators(KSP ksp, Mat A, Mat P) { (void)P; ksp->A = A; return 0; }
PetscErrorCode KSPDestroy(KSP *ksp) { free(*ksp); *ksp = NULL; return 0; }
static PetscErrorCode ksp_solve_columns(KSP ksp, int ncols, const double *b, double *x, int ld)
{
  Mat F;
  PetscCall(MatGetFactor(ksp->A, MATSOLVERPETSC, MAT_FACTOR_LU, &F));
  PetscCall(MatLUFactorNumeric(F, ksp->A, NULL));
  for (int j = 0; j < ncols; ++j) {
    if (x != b) memcpy(x + (size_t)j * ld, b + (size_t)j * ld, sizeof(double) * (size_t)F->m);
    lu_solve(F->m, F->fac, F->piv, x + (size_t)j * ld);
  }
  return MatDestroy(&F);
}
PetscErrorCode KSPMatSolve(KSP ksp, Mat B, Mat X)
{
  PetscCheck(is_dense(B) && is_dense(X), PETSC_COMM_SELF, PETSC_ERR_SUP, "KSPMatSolve: dense right-hand sides only");
  return ksp_solve_columns(ksp, B->n, B->d, X->d, B->m);
}
PetscErrorCode KSPSolve(KSP ksp, Vec b, Vec x) { return ksp_solve_columns(ksp, 1, b->a, x->a, b->n); }

/* ---- MatSOR on SeqAIJ ------------------------------------------------------------------------------ */
/* x_i <- (1 - omega) x_i + omega (b_i - sum_{j != i} a_ij x_j) / (a_ii + shift), rows ascending (forward) / descending (backward);
 * the local sweep types coincide with the global ones on one rank; its * lits directional passes. */
PetscErrorCode MatSOR(Mat A, Vec b, PetscReal omega, MatSORType flag, PetscReal shift, PetscInt its, PetscInt lits, Vec x)
{
  PetscCheck(is_type(A, MATSEQAIJ), PETSC_COMM_SELF, PETSC_ERR_SUP, "MatSOR: seqaij only (PETSc's MatSOR_MPIAIJ has no true parallel sweep, src/pc_sorgibbs.c:239-252)");
  if (flag & SOR_ZERO_INITIAL_GUESS) memset(x->a, 0, sizeof(double) * (size_t)A->m);
  const int fwd = (flag & SOR_FORWARD_SWEEP) || (flag & SOR_LOCAL_FORWARD_SWEEP), bwd = (flag & SOR_BACKWARD_SWEEP) || (flag & SOR_LOCAL_BACKWARD_SWEEP);
  for (PetscInt it = 0; it < its * lits; ++it) {
    for (int pass = 0; pass < 2; ++pass) {
      if ((pass == 0 && !fwd) || (pass == 1 && !bwd)) continue;
      for (PetscInt q = 0; q < A->m; ++q) {
        const PetscInt i = pass == 0 ? q : A->m - 1 - q;
        double         sum = b->a[i], d = 0.0;
        for (PetscInt k = A->i[i]; k < A->i[i + 1]; ++k) {
          if (A->j[k] == i) d = A->a[k];
          else sum -= A->a[k] * x->a[A->j[k]];
        }
        x->a[i] = (1.0 - omega) * x->a[i] + omega * sum / (d + shift);
      }
    }
  }
  return 0;
}

/* ---- Vec extras ------------------------------------------------------------------------------------------------------------ */
PetscErrorCode VecGetOwnershipRange(Vec v, PetscInt *lo, PetscInt *hi)
{
  if (lo) *lo = v->rstart;
  if (hi) *hi = v->rstart + v->n;
  return 0;
}
PetscErrorCode VecGetLocalVector(Vec v, Vec w) { if (w->own) free(w->a); w->a = v->a; w->own = 0; return 0; }
PetscErrorCode VecRestoreLocalVector(Vec v, Vec w) { (void)v; (void)w; return 0; }
PetscErrorCode VecGetLocalVectorRead(Vec v, Vec w) { return VecGetLocalVector(v, w); }
PetscErrorCode VecRestoreLocalVectorRead(Vec v, Vec w) { (void)v; (void)w; return 0; }
PetscErrorCode PetscStrcmp(const char *a, const char *b, PetscBool *flg)
{
  *flg = (a && b && strcmp(a, b) == 0) ? PETSC_TRUE : PETSC_FALSE;
  return 0;
}
PetscErrorCode PetscOptionsInt(const char *opt, const char *text, const char *man, PetscInt cur, PetscInt *v, PetscBool *set)
{
  (void)text; (void)man;
  PetscReal r   = (PetscReal)cur;
  PetscBool got = PETSC_FALSE;
  PetscCall(PetscOptionsGetReal(NULL, NULL, opt, &r, &got));
  if (got) *v = (PetscInt)r;
  if (set) *set = got;
  return 0;
}

/* ---- PC objects used as members of other PCs (pc_sorgibbs.c creates a PCPARSOR for MPIAIJ operators) ------------------------ */
static void stub_direct_register(void);
PetscErrorCode PCCreate(MPI_Comm comm, PC *pc)
{
  *pc              = calloc(1, sizeof(**pc));
  (*pc)->hdr.comm  = comm;
  (*pc)->hdr.refct = 1;
  return 0;
}
PetscErrorCode PCSetType(PC pc, PCType type)
{
  stub_direct_register();
  Mat keep = pc->pmat;
  PC  made = NULL;
  Mat dummy;
  PetscInt z[1] = {0};
  PetscCall(MatStubCreateSeqAIJ(pc->hdr.comm, 0, 0, z, z, (double *)z, &dummy));
  PetscCall(PCStubCreate(type, keep ? keep : dummy, &made));
  pc->data = made->data;
  memcpy(pc->ops, made->ops, sizeof(pc->ops));
  memcpy(pc->hdr.composed, made->hdr.composed, sizeof(pc->hdr.composed));
  free(made);
  return MatDestroy(&dummy);
}
/* built-in exact solver ("cholesky" / "lu" of PETSc proper): y = pmat^-1 x by the dense LU of KSPSolve */
static PetscErrorCode PCApply_StubDirect(PC pc, Vec x, Vec y)
{
  KSP ksp;
  PetscCall(KSPCreate(pc->hdr.comm, &ksp));
  PetscCall(KSPSetOperators(ksp, pc->pmat, pc->pmat));
  PetscCall(KSPSolve(ksp, x, y));
  return KSPDestroy(&ksp);
}
static PetscErrorCode PCCreate_StubDirect(PC pc)
{
  pc->ops->apply = PCApply_StubDirect;
  return 0;
}
static void stub_direct_register(void)
{
  static int done = 0;
  if (done) return;
  done = 1;
  PCRegister("cholesky", PCCreate_StubDirect);
  PCRegister("lu", PCCreate_StubDirect);
}
PetscErrorCode PCSetOperators(PC pc, Mat A, Mat P) { pc->mat = A; pc->pmat = P; return 0; }
PetscErrorCode PCApply(PC pc, Vec x, Vec y)
{
  PetscCheck(pc->ops->apply, PETSC_COMM_SELF, PETSC_ERR_SUP, "PCApply: the PC type has no apply");
  return pc->ops->apply(pc, x, y);
}
PetscErrorCode PCApplyRichardson(PC pc, Vec b, Vec y, Vec w, PetscReal rtol, PetscReal abstol, PetscReal dtol, PetscInt its, PetscBool guesszero, PetscInt *outits, PCRichardsonConvergedReason *reason)
{
  PetscCheck(pc->ops->applyrichardson, PETSC_COMM_SELF, PETSC_ERR_SUP, "PCApplyRichardson: the PC type has no applyrichardson");
  return pc->ops->applyrichardson(pc, b, y, w, rtol, abstol, dtol, its, guesszero, outits, reason);
}
PetscErrorCode PCSetFromOptions(PC pc) { return pc->ops->setfromoptions ? pc->ops->setfromoptions(pc, NULL) : 0; }
PetscErrorCode PCReset(PC pc) { return pc->ops->reset ? pc->ops->reset(pc) : 0; }
PetscErrorCode PCSetOptionsPrefix(PC pc, const char *prefix)
{
  pc->hdr.prefix = prefix ? strdup(prefix) : NULL; /* a few bytes per PC, never freed: test infrastructure */
  return 0;
}
PetscErrorCode PCAppendOptionsPrefix(PC pc, const char *prefix)
{
  const char *old = pc->hdr.prefix ? pc->hdr.prefix : "";
  char       *s   = malloc(strlen(old) + strlen(prefix) + 1);
  strcpy(s, old);
  strcat(s, prefix);
  pc->hdr.prefix = s;
  return 0;
}
PetscErrorCode PCGetOptionsPrefix(PC pc, const char **prefix)
{
  *prefix = pc->hdr.prefix;
  return 0;
}
PetscErrorCode PCSetUp(PC pc) { return pc->ops->setup ? pc->ops->setup(pc) : 0; }
PetscErrorCode PCDestroy(PC *pc)
{
  if (*pc) return PCStubDestroy(pc);
  return 0;
}

/* ---- BLAS / LAPACK: the two routines src/pc_chols.c calls, column-major, lower triangle ------------------------------------ */
void LAPACKpotrf_(const char *uplo, const PetscBLASInt *n, PetscScalar *a, const PetscBLASInt *lda, PetscBLASInt *info)
{
  const int N = *n, ld = *lda;
  *info = 0;
  if (uplo[0] != 'L' && uplo[0] != 'l') { *info = -1; return; }
  for (int j = 0; j < N; ++j) { /* left-looking column Cholesky: a_jj, then the column below it */
    double d = a[j + (size_t)j * ld];
    for (int k = 0; k < j; ++k) d -= a[j + (size_t)k * ld] * a[j + (size_t)k * ld];
    if (!(d > 0.0)) { *info = j + 1; return; }
    d = sqrt(d);
    a[j + (size_t)j * ld] = d;
    for (int i = j + 1; i < N; ++i) {
      double s = a[i + (size_t)j * ld];
      for (int k = 0; k < j; ++k) s -= a[i + (size_t)k * ld] * a[j + (size_t)k * ld];
      a[i + (size_t)j * ld] = s / d;
    }
  }
}
void BLAStrsv_(const char *uplo, const char *trans, const char *diag, const PetscBLASInt *n, const PetscScalar *a, const PetscBLASInt *lda, PetscScalar *x, const PetscBLASInt *incx)
{
  (void)uplo; (void)diag;
  const int N = *n, ld = *lda, inc = *incx;
  if (trans[0] == 'N' || trans[0] == 'n') { /* L x = b: column sweep of the reference dtrsv */
    for (int j = 0; j < N; ++j) {
      if (x[j * inc] != 0.0) {
        x[j * inc] /= a[j + (size_t)j * ld];
        const double t = x[j * inc];
        for (int i = j + 1; i < N; ++i) x[i * inc] -= t * a[i + (size_t)j * ld];
      }
    }
  } else { /* L^T x = b */
    for (int j = N - 1; j >= 0; --j) {
      double t = x[j * inc];
      for (int i = N - 1; i > j; --i) t -= a[i + (size_t)j * ld] * x[i * inc];
      x[j * inc] = t / a[j + (size_t)j * ld];
    }
  }
}

/* ---- FFTW: unnormalised DFT, out[k] = sum_j in[j] exp(sign 2 pi i j k / n) ------------------------------------------------- */
struct fftw_plan_s {
  int           n, sign;
  fftw_complex *in, *out;
};
fftw_plan fftw_plan_dft_1d(int n, fftw_complex *in, fftw_complex *out, int sign, unsigned flags)
{
  (void)flags;
  fftw_plan p = malloc(sizeof(*p));
  p->n = n; p->sign = sign; p->in = in; p->out = out;
  return p;
}
void fftw_destroy_plan(fftw_plan p) { free(p); }
void fftw_execute(const fftw_plan p)
{
  const int     n = p->n;
  fftw_complex *w = malloc(sizeof(fftw_complex) * (size_t)n);
  if (n > 0 && (n & (n - 1)) == 0) { /* iterative radix-2, decimation in time */
    int bits = 0;
    while ((1 << bits) < n) ++bits;
    for (int i = 0; i < n; ++i) {
      int r = 0;
      for (int b = 0; b < bits; ++b) r |= ((i >> b) & 1) << (bits - 1 - b);
      w[r] = p->in[i];
    }
    for (int len = 2; len <= n; len <<= 1) {
      const double ang = p->sign * 2.0 * 3.14159265358979323846 / len;
      for (int s = 0; s < n; s += len)
        for (int k = 0; k < len / 2; ++k) {
          const fftw_complex tw = cos(ang * k) + sin(ang * k) * I;
          const fftw_complex u = w[s + k], v = w[s + k + len / 2] * tw;
          w[s + k]           = u + v;
          w[s + k + len / 2] = u - v;
        }
    }
  } else {
    for (int k = 0; k < n; ++k) {
      fftw_complex s = 0;
      for (int j = 0; j < n; ++j) {
        const double ang = p->sign * 2.0 * 3.14159265358979323846 * (double)(((long long)j * k) % n) / n;
        s += p->in[j] * (cos(ang) + sin(ang) * I);
      }
      w[k] = s;
    }
  }
  memcpy(p->out, w, sizeof(fftw_complex) * (size_t)n);
  free(w);
}


/* ---- DMDA for src/problems.c: one rank, 2D; the operator is assembled into a dense matrix (row = j mx + i) ------------------ */
static PetscInt g_dmda_mx = 0;
PetscErrorCode DMStubCreate2d(PetscInt mx, PetscInt my, DM *dm)
{
  *dm       = calloc(1, sizeof(**dm));
  (*dm)->mx = mx;
  (*dm)->my = my;
  g_dmda_mx = mx;
  return 0;
}
PetscErrorCode DMDestroy(DM *dm) { free(*dm); *dm = NULL; return 0; }
PetscErrorCode DMDAGetInfo(DM dm, PetscInt *dim, PetscInt *M, PetscInt *N, PetscInt *P, PetscInt *m, PetscInt *n, PetscInt *p, PetscInt *dof, PetscInt *s, void *bx, void *by, void *bz, void *st)
{
  (void)m; (void)n; (void)p; (void)dof; (void)s; (void)bx; (void)by; (void)bz; (void)st;
  if (dim) *dim = 2;
  if (M) *M = dm->mx;
  if (N) *N = dm->my;
  if (P) *P = 1;
  return 0;
}
PetscErrorCode DMDAGetCorners(DM dm, PetscInt *xs, PetscInt *ys, PetscInt *zs, PetscInt *xm, PetscInt *ym, PetscInt *zm)
{
  if (xs) *xs = 0;
  if (ys) *ys = 0;
  if (zs) *zs = 0;
  if (xm) *xm = dm->mx;
  if (ym) *ym = dm->my;
  if (zm) *zm = 1;
  return 0;
}
PetscErrorCode MatSetValuesStencil(Mat A, PetscInt m, const MatStencil *rows, PetscInt n, const MatStencil *cols, const PetscScalar *v, InsertMode mode)
{
  PetscCheck(is_dense(A) && g_dmda_mx > 0, PETSC_COMM_SELF, PETSC_ERR_SUP, "MatSetValuesStencil: dense matrix of a stub DMDA only");
  for (PetscInt r = 0; r < m; ++r)
    for (PetscInt c = 0; c < n; ++c) {
      const PetscInt R = rows[r].j * g_dmda_mx + rows[r].i, Cc = cols[c].j * g_dmda_mx + cols[c].i;
      double        *e = A->d + R + (size_t)Cc * (size_t)A->m;
      *e               = mode == ADD_VALUES ? *e + v[r * n + c] : v[r * n + c];
    }
  return 0;
}
PetscErrorCode MatAssemblyBegin(Mat A, MatAssemblyType t) { (void)A; (void)t; return 0; }
PetscErrorCode MatAssemblyEnd(Mat A, MatAssemblyType t) { (void)A; (void)t; return 0; }
