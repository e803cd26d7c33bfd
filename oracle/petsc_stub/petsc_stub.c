/* petsc_stub.c -- toy bodies for the PETSc/MPI API slice declared in petsc_stub.h (TEST INFRASTRUCTURE ONLY).
 * See petsc_stub.h for what is real (the reference's own source files) and what is emulated here. */
#include "petsc_stub.h"

#include <pthread.h>
#include <stdarg.h>

/* ---- rank emulation ------------------------------------------------------------------------------------ */
static int               g_nranks = 1;
static __thread int      t_rank   = 0;
static pthread_barrier_t g_barrier;
static int               g_barrier_on = 0;

void PetscStubWorldBegin(int nranks)
{
  g_nranks = nranks;
  if (nranks > 1) {
    pthread_barrier_init(&g_barrier, NULL, (unsigned)nranks);
    g_barrier_on = 1;
  }
}
void PetscStubWorldEnd(void)
{
  if (g_barrier_on) pthread_barrier_destroy(&g_barrier);
  g_barrier_on = 0;
  g_nranks     = 1;
}
void PetscStubSetRank(int rank) { t_rank = rank; }
void PetscStubBarrier(void)
{
  if (g_barrier_on) pthread_barrier_wait(&g_barrier);
}
int MPI_Comm_size(MPI_Comm comm, int *size)
{
  *size = comm == MPI_COMM_SELF ? 1 : g_nranks;
  return 0;
}
int MPI_Comm_rank(MPI_Comm comm, int *rank)
{
  *rank = comm == MPI_COMM_SELF ? 0 : t_rank;
  return 0;
}

/* ---- errors / objects / logging / options ------------------------------------------------------------------ */
static char g_errmsg[512];
PetscErrorCode PetscStubError(int code, const char *file, int line, const char *fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  int k = snprintf(g_errmsg, sizeof g_errmsg, "%s:%d: ", file, line);
  vsnprintf(g_errmsg + k, sizeof g_errmsg - (size_t)k, fmt, ap);
  va_end(ap);
  return code ? code : PETSC_ERR_PLIB;
}
const char *PetscStubLastError(void) { return g_errmsg; }

MPI_Comm       PetscObjectComm(PetscObject o) { return o->comm; }
PetscErrorCode PetscObjectGetComm(PetscObject o, MPI_Comm *comm)
{
  *comm = o->comm;
  return 0;
}
PetscErrorCode PetscObjectReference(PetscObject o)
{
  o->refct++;
  return 0;
}
PetscErrorCode PetscObjectComposeFunction_Stub(PetscObject o, const char *name, void (*f)(void))
{
  for (int k = 0; k < 4; ++k)
    if (!o->composed[k].name || strcmp(o->composed[k].name, name) == 0) {
      o->composed[k].name = name;
      o->composed[k].f    = f;
      return 0;
    }
  return PETSC_ERR_PLIB;
}
void (*PetscStubQueryFunction(PetscObject o, const char *name))(void)
{
  for (int k = 0; k < 4; ++k)
    if (o->composed[k].name && strcmp(o->composed[k].name, name) == 0) return o->composed[k].f;
  return NULL;
}
int MPI_Bcast(void *buf, int count, int datatype, int root, MPI_Comm comm)
{
  (void)buf; (void)count; (void)datatype; (void)root; (void)comm;
  return 0;
}
PetscErrorCode PetscClassIdRegister(const char *name, PetscClassId *id)
{
  (void)name;
  *id = 1;
  return 0;
}
PetscErrorCode PetscLogEventRegister(const char *name, PetscClassId id, PetscLogEvent *e)
{
  (void)name;
  (void)id;
  *e = 1;
  return 0;
}

static struct { char key[64], val[64]; } g_opts[32];
static int g_nopts = 0;
PetscErrorCode PetscStubOptionsSet(const char *key, const char *value)
{
  if (g_nopts >= 32) return PETSC_ERR_PLIB;
  snprintf(g_opts[g_nopts].key, 64, "%s", key);
  snprintf(g_opts[g_nopts].val, 64, "%s", value ? value : "");
  g_nopts++;
  return 0;
}
PetscErrorCode PetscStubOptionsClear(void)
{
  g_nopts = 0;
  return 0;
}
static const char *opt_find(const char *name)
{
  for (int i = g_nopts - 1; i >= 0; --i)
    if (strcmp(g_opts[i].key, name) == 0) return g_opts[i].val;
  return NULL;
}
PetscErrorCode PetscOptionsGetReal(PetscOptions o, const char *pre, const char *name, PetscReal *v, PetscBool *set)
{
  (void)o;
  (void)pre;
  const char *s = opt_find(name);
  if (s) *v = atof(s);
  if (set) *set = s ? PETSC_TRUE : PETSC_FALSE;
  return 0;
}
PetscErrorCode PetscOptionsGetString(PetscOptions o, const char *pre, const char *name, char *buf, size_t len, PetscBool *set)
{
  (void)o;
  (void)pre;
  const char *s = opt_find(name);
  if (s) snprintf(buf, len, "%s", s);
  if (set) *set = s ? PETSC_TRUE : PETSC_FALSE;
  return 0;
}
PetscErrorCode PetscOptionsRangeReal(const char *opt, const char *text, const char *man, PetscReal cur, PetscReal *v, PetscBool *set, PetscReal lo, PetscReal hi)
{
  (void)text;
  (void)man;
  (void)cur;
  const char *s = opt_find(opt);
  if (s) {
    const double x = atof(s);
    PetscCheck(x >= lo && x <= hi, PETSC_COMM_SELF, PETSC_ERR_ARG_WRONG, "option %s = %g out of range [%g, %g]", opt, x, lo, hi);
    *v = x;
  }
  if (set) *set = s ? PETSC_TRUE : PETSC_FALSE;
  return 0;
}
PetscErrorCode PetscOptionsString(const char *opt, const char *text, const char *man, const char *cur, char *v, size_t len, PetscBool *set)
{
  (void)text;
  (void)man;
  (void)cur;
  const char *s = opt_find(opt);
  if (s) snprintf(v, len, "%s", s);
  if (set) *set = s ? PETSC_TRUE : PETSC_FALSE;
  return 0;
}
PetscErrorCode PetscOptionsBool(const char *opt, const char *text, const char *man, PetscBool cur, PetscBool *v, PetscBool *set)
{
  (void)text;
  (void)man;
  (void)cur;
  const char *s = opt_find(opt);
  if (s) *v = (s[0] == 0 || strcmp(s, "1") == 0 || strcmp(s, "true") == 0) ? PETSC_TRUE : PETSC_FALSE;
  if (set) *set = s ? PETSC_TRUE : PETSC_FALSE;
  return 0;
}
PetscErrorCode PetscViewerASCIIPrintf(PetscViewer v, const char *fmt, ...)
{
  (void)v;
  va_list ap;
  va_start(ap, fmt);
  vprintf(fmt, ap);
  va_end(ap);
  return 0;
}

/* ---- Vec --------------------------------------------------------------------------------------------------- */
PetscErrorCode VecStubCreate(MPI_Comm comm, PetscInt n, PetscInt N, PetscInt rstart, double *array, Vec *v)
{
  Vec w       = calloc(1, sizeof(*w));
  w->hdr.comm = comm;
  w->hdr.refct = 1;
  w->n        = n;
  w->N        = N;
  w->rstart   = rstart;
  if (array) w->a = array;
  else {
    w->a   = calloc((size_t)(n > 0 ? n : 1), sizeof(double));
    w->own = 1;
  }
  *v = w;
  return 0;
}
PetscErrorCode VecCreateSeq(MPI_Comm comm, PetscInt n, Vec *v) { return VecStubCreate(comm, n, n, 0, NULL, v); }
PetscErrorCode VecCreateMPIWithArray(MPI_Comm comm, PetscInt bs, PetscInt n, PetscInt N, const PetscScalar *array, Vec *v)
{
  (void)bs;
  PetscErrorCode e = VecStubCreate(comm, n, N, 0, (double *)array, v); /* layout template only (mc_sor.c:169) */
  if (!array) {
    free((*v)->a);
    (*v)->a   = NULL;
    (*v)->own = 0;
  }
  return e;
}
PetscErrorCode VecDestroy(Vec *v)
{
  if (*v && --(*v)->hdr.refct <= 0) {
    if ((*v)->own) free((*v)->a);
    free(*v);
  }
  *v = NULL;
  return 0;
}
PetscErrorCode VecDuplicate(Vec v, Vec *w) { return VecStubCreate(v->hdr.comm, v->n, v->N, v->rstart, NULL, w); }
PetscErrorCode VecGetArray(Vec v, PetscScalar **a)
{
  *a = v->a;
  return 0;
}
PetscErrorCode VecRestoreArray(Vec v, PetscScalar **a)
{
  (void)v;
  *a = NULL;
  return 0;
}
PetscErrorCode VecGetArrayRead(Vec v, const PetscScalar **a)
{
  *a = v->a;
  return 0;
}
PetscErrorCode VecRestoreArrayRead(Vec v, const PetscScalar **a)
{
  (void)v;
  *a = NULL;
  return 0;
}
PetscErrorCode VecGetSize(Vec v, PetscInt *N)
{
  *N = v->N;
  return 0;
}
PetscErrorCode VecGetLocalSize(Vec v, PetscInt *n)
{
  *n = v->n;
  return 0;
}
PetscErrorCode VecReciprocal(Vec v)
{
  for (PetscInt i = 0; i < v->n; ++i)
    if (v->a[i] != 0.0) v->a[i] = 1.0 / v->a[i];
  return 0;
}
PetscErrorCode VecScale(Vec v, PetscScalar s)
{
  for (PetscInt i = 0; i < v->n; ++i) v->a[i] *= s;
  return 0;
}
PetscErrorCode VecAXPY(Vec y, PetscScalar a, Vec x)
{
  if (a == 1.0) /* BLAS daxpy with alpha = 1: a plain add, no product rounding */
    for (PetscInt i = 0; i < y->n; ++i) y->a[i] = y->a[i] + x->a[i];
  else
    for (PetscInt i = 0; i < y->n; ++i) y->a[i] += a * x->a[i];
  return 0;
}
PetscErrorCode VecCopy(Vec x, Vec y)
{
  memcpy(y->a, x->a, sizeof(double) * (size_t)x->n);
  return 0;
}
PetscErrorCode VecZeroEntries(Vec v)
{
  memset(v->a, 0, sizeof(double) * (size_t)v->n);
  return 0;
}
PetscErrorCode VecSqrtAbs(Vec v)
{
  for (PetscInt i = 0; i < v->n; ++i) v->a[i] = sqrt(fabs(v->a[i]));
  return 0;
}
PetscErrorCode VecPointwiseMult(Vec w, Vec x, Vec y)
{
  for (PetscInt i = 0; i < w->n; ++i) w->a[i] = x->a[i] * y->a[i];
  return 0;
}

/* ---- IS / ISColoring --------------------------------------------------------------------------------------------- */
PetscErrorCode ISCreateGeneral(MPI_Comm comm, PetscInt n, const PetscInt idx[], PetscCopyMode mode, IS *is)
{
  (void)mode; /* always copies */
  IS s        = calloc(1, sizeof(*s));
  s->hdr.comm = comm;
  s->n        = n;
  s->idx      = malloc(sizeof(PetscInt) * (size_t)(n > 0 ? n : 1));
  if (n > 0) memcpy(s->idx, idx, sizeof(PetscInt) * (size_t)n);
  *is = s;
  return 0;
}
PetscErrorCode ISCreateStride(MPI_Comm comm, PetscInt n, PetscInt first, PetscInt step, IS *is)
{
  PetscInt *t = malloc(sizeof(PetscInt) * (size_t)(n > 0 ? n : 1));
  for (PetscInt i = 0; i < n; ++i) t[i] = first + i * step;
  PetscErrorCode e = ISCreateGeneral(comm, n, t, PETSC_COPY_VALUES, is);
  free(t);
  return e;
}
PetscErrorCode ISDestroy(IS *is)
{
  if (*is) {
    free((*is)->idx);
    free(*is);
  }
  *is = NULL;
  return 0;
}
PetscErrorCode ISGetLocalSize(IS is, PetscInt *n)
{
  *n = is->n;
  return 0;
}
PetscErrorCode ISGetIndices(IS is, const PetscInt **idx)
{
  *idx = is->idx;
  return 0;
}
PetscErrorCode ISRestoreIndices(IS is, const PetscInt **idx)
{
  (void)is;
  *idx = NULL;
  return 0;
}
/* IS_COLORING_LOCAL semantics: colour c's IS lists the LOCAL row ids with that colour, ascending */
PetscErrorCode ISColoringCreate(MPI_Comm comm, PetscInt ncolors, PetscInt n, const ISColoringValue colors[], PetscCopyMode mode, ISColoring *isc)
{
  ISColoring c = calloc(1, sizeof(*c));
  c->comm      = comm;
  c->ncolors   = ncolors;
  c->n         = n;
  c->colors    = malloc(sizeof(ISColoringValue) * (size_t)(n > 0 ? n : 1));
  if (n > 0) memcpy(c->colors, colors, sizeof(ISColoringValue) * (size_t)n);
  if (mode == PETSC_OWN_POINTER) free((void *)colors);
  c->is = calloc((size_t)ncolors, sizeof(IS));
  PetscInt *tmp = malloc(sizeof(PetscInt) * (size_t)(n > 0 ? n : 1));
  for (PetscInt k = 0; k < ncolors; ++k) {
    PetscInt cnt = 0;
    for (PetscInt i = 0; i < n; ++i)
      if (c->colors[i] == k) tmp[cnt++] = i;
    ISCreateGeneral(MPI_COMM_SELF, cnt, tmp, PETSC_COPY_VALUES, &c->is[k]);
  }
  free(tmp);
  *isc = c;
  return 0;
}
PetscErrorCode ISColoringSetType(ISColoring isc, ISColoringType t)
{
  (void)isc;
  (void)t;
  return 0;
}
PetscErrorCode ISColoringGetIS(ISColoring isc, PetscCopyMode mode, PetscInt *n, IS *iss[])
{
  (void)mode;
  if (n) *n = isc->ncolors;
  if (iss) *iss = isc->is;
  return 0;
}
PetscErrorCode ISColoringRestoreIS(ISColoring isc, PetscCopyMode mode, IS *iss[])
{
  (void)isc;
  (void)mode;
  if (iss) *iss = NULL;
  return 0;
}
PetscErrorCode ISColoringDestroy(ISColoring *isc)
{
  if (*isc) {
    for (PetscInt k = 0; k < (*isc)->ncolors; ++k) ISDestroy(&(*isc)->is[k]);
    free((*isc)->is);
    free((*isc)->colors);
    free(*isc);
  }
  *isc = NULL;
  return 0;
}

/* ---- Mat ------------------------------------------------------------------------------------------------------ */
PetscErrorCode MatStubCreateSeqAIJ(MPI_Comm comm, PetscInt m, PetscInt n, const PetscInt *i, const PetscInt *j, const double *a, Mat *A)
{
  Mat B       = calloc(1, sizeof(*B));
  B->hdr.comm = comm;
  B->hdr.refct = 1;
  B->type     = MATSEQAIJ;
  B->m = B->M = m;
  B->n = B->N = n;
  const size_t nnz = (size_t)i[m];
  B->i = malloc(sizeof(PetscInt) * (size_t)(m + 1));
  B->j = malloc(sizeof(PetscInt) * (nnz ? nnz : 1));
  B->a = malloc(sizeof(double) * (nnz ? nnz : 1));
  memcpy(B->i, i, sizeof(PetscInt) * (size_t)(m + 1));
  memcpy(B->j, j, sizeof(PetscInt) * nnz);
  memcpy(B->a, a, sizeof(double) * nnz);
  *A = B;
  return 0;
}
PetscErrorCode MatStubCreateMPIAIJ(PetscInt m, PetscInt M, PetscInt rstart, Mat Ad, Mat Ao, const PetscInt *colmap, PetscInt ncolmap, Mat *A)
{
  Mat B       = calloc(1, sizeof(*B));
  B->hdr.comm = MPI_COMM_WORLD;
  B->hdr.refct = 1;
  B->type     = MATMPIAIJ;
  B->m = B->n = m;
  B->M = B->N = M;
  B->rstart   = rstart;
  B->Ad       = Ad;
  B->Ao       = Ao;
  B->colmap   = malloc(sizeof(PetscInt) * (size_t)(ncolmap > 0 ? ncolmap : 1));
  if (ncolmap > 0) memcpy(B->colmap, colmap, sizeof(PetscInt) * (size_t)ncolmap);
  *A = B;
  return 0;
}
PetscErrorCode MatStubInjectColoring(Mat A, PetscInt ncolors, const ISColoringValue *colors)
{
  A->inject_colors  = colors;
  A->inject_ncolors = ncolors;
  return 0;
}
PetscErrorCode MatDestroy(Mat *A)
{
  if (*A && --(*A)->hdr.refct <= 0) {
    if ((*A)->Ad) MatDestroy(&(*A)->Ad);
    if ((*A)->Ao) MatDestroy(&(*A)->Ao);
    free((*A)->i);
    free((*A)->j);
    free((*A)->a);
    free((*A)->colmap);
    free((*A)->d);
    free((*A)->fac);
    free((*A)->piv);
    free(*A);
  }
  *A = NULL;
  return 0;
}
PetscErrorCode MatGetType(Mat A, MatType *t)
{
  *t = A->type;
  return 0;
}
PetscErrorCode MatGetSize(Mat A, PetscInt *M, PetscInt *N)
{
  if (M) *M = A->M;
  if (N) *N = A->N;
  return 0;
}
PetscErrorCode MatGetLocalSize(Mat A, PetscInt *m, PetscInt *n)
{
  if (m) *m = A->m;
  if (n) *n = A->n;
  return 0;
}
PetscErrorCode MatGetOwnershipRange(Mat A, PetscInt *lo, PetscInt *hi)
{
  if (lo) *lo = A->rstart;
  if (hi) *hi = A->rstart + A->m;
  return 0;
}
PetscErrorCode MatMPIAIJGetSeqAIJ(Mat A, Mat *Ad, Mat *Ao, const PetscInt **colmap)
{
  PetscCheck(strcmp(A->type, MATMPIAIJ) == 0, PETSC_COMM_SELF, PETSC_ERR_SUP, "not an mpiaij matrix");
  if (Ad) *Ad = A->Ad;
  if (Ao) *Ao = A->Ao;
  if (colmap) *colmap = A->colmap;
  return 0;
}
PetscErrorCode MatSeqAIJGetCSRAndMemType(Mat A, const PetscInt **i, const PetscInt **j, PetscScalar **a, PetscMemType *mt)
{
  PetscCheck(strcmp(A->type, MATSEQAIJ) == 0, PETSC_COMM_SELF, PETSC_ERR_SUP, "not a seqaij matrix");
  if (i) *i = A->i;
  if (j) *j = A->j;
  if (a) *a = A->a;
  if (mt) *mt = PETSC_MEMTYPE_HOST;
  return 0;
}
PetscErrorCode MatGetDiagonal(Mat A, Vec d)
{
  if (A->d) {
    for (PetscInt r = 0; r < A->m; ++r) d->a[r] = A->d[(size_t)r + (size_t)r * A->m];
    return 0;
  }
  Mat S = strcmp(A->type, MATMPIAIJ) == 0 ? A->Ad : A;
  for (PetscInt r = 0; r < S->m; ++r) {
    d->a[r] = 0.0;
    for (PetscInt k = S->i[r]; k < S->i[r + 1]; ++k)
      if (S->j[k] == r) d->a[r] = S->a[k];
  }
  return 0;
}
PetscErrorCode MatCreateVecs(Mat A, Vec *right, Vec *left)
{
  if (right) PetscCall(VecStubCreate(A->hdr.comm, A->n, A->N, A->rstart, NULL, right));
  if (left) PetscCall(VecStubCreate(A->hdr.comm, A->m, A->M, A->rstart, NULL, left));
  return 0;
}
/* MatMult*, MATLRC, dense matrices, factorisations, KSP, MatSOR: petsc_stub_dense.c */

/* ---- MatColoring: hands back the colouring the driver injected (PETSc's JP is randomised and rank dependent) ---- */
struct _p_MatColoring {
  Mat A;
};
PetscErrorCode MatColoringCreate(Mat A, MatColoring *mc)
{
  *mc      = calloc(1, sizeof(**mc));
  (*mc)->A = A;
  return 0;
}
PetscErrorCode MatColoringSetDistance(MatColoring mc, PetscInt d)
{
  (void)mc;
  PetscCheck(d == 1, PETSC_COMM_SELF, PETSC_ERR_SUP, "petsc_stub: distance-1 colourings only");
  return 0;
}
PetscErrorCode MatColoringSetType(MatColoring mc, MatColoringType t)
{
  (void)mc;
  (void)t;
  return 0;
}
PetscErrorCode MatColoringApply(MatColoring mc, ISColoring *isc)
{
  Mat A = mc->A;
  PetscCheck(A->inject_colors, PETSC_COMM_SELF, PETSC_ERR_SUP, "petsc_stub: no colouring injected (MatStubInjectColoring)");
  return ISColoringCreate(A->hdr.comm, A->inject_ncolors, A->m, A->inject_colors, PETSC_COPY_VALUES, isc);
}
PetscErrorCode MatColoringDestroy(MatColoring *mc)
{
  free(*mc);
  *mc = NULL;
  return 0;
}

/* ---- VecScatter: gather global entries ix of the distributed vector x into the sequential y ---------------------- */
struct _p_VecScatter {
  PetscInt  n;
  PetscInt *idx;
};
PetscErrorCode VecScatterCreate(Vec x, IS ix, Vec y, IS iy, VecScatter *sct)
{
  (void)x;
  PetscCheck(iy == NULL && y->n == ix->n, PETSC_COMM_SELF, PETSC_ERR_SUP, "petsc_stub: only global -> sequential gathers");
  VecScatter s = calloc(1, sizeof(*s));
  s->n         = ix->n;
  s->idx       = malloc(sizeof(PetscInt) * (size_t)(ix->n > 0 ? ix->n : 1));
  if (ix->n > 0) memcpy(s->idx, ix->idx, sizeof(PetscInt) * (size_t)ix->n);
  *sct = s;
  return 0;
}
PetscErrorCode VecScatterBegin(VecScatter sct, Vec x, Vec y, InsertMode im, ScatterMode sm)
{
  PetscCheck(im == INSERT_VALUES && sm == SCATTER_FORWARD, PETSC_COMM_SELF, PETSC_ERR_SUP, "petsc_stub: forward insert only");
  PetscStubBarrier(); /* every rank has finished writing the previous colour */
  const double *global = x->a - x->rstart; /* the rank threads' vectors are slices of one array (ref_driver.c) */
  for (PetscInt k = 0; k < sct->n; ++k) y->a[k] = global[sct->idx[k]];
  return 0;
}
PetscErrorCode VecScatterEnd(VecScatter sct, Vec x, Vec y, InsertMode im, ScatterMode sm)
{
  (void)sct; (void)x; (void)y; (void)im; (void)sm;
  PetscStubBarrier(); /* nobody overwrites y values before every rank has gathered */
  return 0;
}
PetscErrorCode VecScatterDestroy(VecScatter *sct)
{
  if (*sct) {
    free((*sct)->idx);
    free(*sct);
  }
  *sct = NULL;
  return 0;
}

/* ---- PetscRandom: rander48 (48-bit LCG, value X 2^-48), seed layout of PetscRandomSeed_Rander48 -------------------- */
struct _p_PetscRandom {
  struct _p_PetscObject hdr;
  PetscInt64            seed;
  unsigned long long    x;
};
static void rander48_seed(PetscRandom r) { r->x = 0x330EULL | (((unsigned long long)r->seed & 0xFFFFULL) << 16) | ((((unsigned long long)r->seed >> 16) & 0xFFFFULL) << 32); }
PetscErrorCode PetscRandomCreate(MPI_Comm comm, PetscRandom *r)
{
  PetscRandom p = calloc(1, sizeof(*p));
  p->hdr.comm   = comm;
  p->hdr.refct  = 1;
  p->seed       = 0x12345678; /* PETSc's default seed */
  rander48_seed(p);
  *r = p;
  return 0;
}
PetscErrorCode PetscRandomSetFromOptions(PetscRandom r)
{
  (void)r;
  return 0;
}
PetscErrorCode PetscRandomSetSeed(PetscRandom r, PetscInt64 seed)
{
  r->seed = seed;
  return 0;
}
PetscErrorCode PetscRandomGetSeed(PetscRandom r, PetscInt64 *seed)
{
  *seed = r->seed;
  return 0;
}
PetscErrorCode PetscRandomSeed(PetscRandom r)
{
  rander48_seed(r);
  return 0;
}
PetscErrorCode PetscRandomGetValueReal(PetscRandom r, PetscReal *v)
{
  r->x = (0x5DEECE66DULL * r->x + 0xBULL) & 0xFFFFFFFFFFFFULL;
  *v   = ldexp((double)(r->x & 0xFFFF), -48) + ldexp((double)((r->x >> 16) & 0xFFFF), -32) + ldexp((double)((r->x >> 32) & 0xFFFF), -16);
  return 0;
}
PetscErrorCode PetscRandomDestroy(PetscRandom *r)
{
  if (*r && --(*r)->hdr.refct <= 0) free(*r);
  *r = NULL;
  return 0;
}

/* ---- PC registry --------------------------------------------------------------------------------------------------- */
static struct { const char *name; PetscErrorCode (*create)(PC); } g_pcs[16];
static int g_npcs = 0;
PetscErrorCode PCRegister(const char *name, PetscErrorCode (*create)(PC))
{
  for (int i = 0; i < g_npcs; ++i)
    if (strcmp(g_pcs[i].name, name) == 0) {
      g_pcs[i].create = create;
      return 0;
    }
  if (g_npcs >= 16) return PETSC_ERR_PLIB;
  g_pcs[g_npcs].name   = name;
  g_pcs[g_npcs].create = create;
  g_npcs++;
  return 0;
}
PetscErrorCode PCStubCreate(const char *type, Mat pmat, PC *pc)
{
  for (int i = 0; i < g_npcs; ++i)
    if (strcmp(g_pcs[i].name, type) == 0) {
      PC p        = calloc(1, sizeof(*p));
      p->hdr.comm = pmat->hdr.comm;
      p->hdr.refct = 1;
      p->mat = p->pmat = pmat;
      PetscCall(g_pcs[i].create(p));
      *pc = p;
      return 0;
    }
  return PetscStubError(PETSC_ERR_SUP, __FILE__, __LINE__, "unknown PC type %s", type);
}
PetscErrorCode PCStubDestroy(PC *pc)
{
  if (*pc && --(*pc)->hdr.refct <= 0) { /* reference counted like PetscObjectDereference */
    if ((*pc)->ops->destroy) PetscCall((*pc)->ops->destroy(*pc));
    free(*pc);
  }
  *pc = NULL;
  return 0;
}
