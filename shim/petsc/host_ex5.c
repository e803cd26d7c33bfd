/* host_ex5.c -- a C host program in the shape of the reference's examples/ex5.c (and of the MCSOR part of examples/ex3.c),
 * driving the PETSc-typed MCSOR API of the shim: MCSORCreate / SetUp / SetSweepType / Apply / SetOmega / GetNumColors /
 * GetISColoring / BuildLRCCorrection / Destroy (include/parmgmc/mc_sor.h:21-30).  PETSc objects come from oracle/petsc_stub
 * (TEST INFRASTRUCTURE); the sweeps run in libparmgmc_b200.so on the GPU.
 *
 *   (1) examples/ex5.c:52-68: a forward sweep followed by a backward sweep equals one symmetric sweep, ||.||_2 < 1e-15
 *   (2) the same operator handed over as a one-rank MATMPIAIJ (diagonal block + empty off-diagonal block) gives the same sweep
 *   (3) examples/ex3.c:111-118 (-with_lr): MCSOR on a MATLRC operator A + B S B^T.  The sweep with the shim-built correction equals
 *       the sweep on A followed by y -= Bb (B^T y) with Bb from MCSORBuildLRCCorrection(MCSORApply on A), and the SOR
 *       iteration converges to the solution of (A + B S B^T) x = b
 *   (4) the colouring handed back by MCSORGetISColoring is a valid distance-1 colouring with MCSORGetNumColors colours
 */
#include "host_common.h"

typedef struct _MCSOR {
  void *ctx;
} *MCSOR;
PetscErrorCode ParMGMCInitialize(void);
PetscErrorCode ParMGMCFinalize(void);
PetscErrorCode MCSORCreate(Mat, MCSOR *);
PetscErrorCode MCSORSetUp(MCSOR);
PetscErrorCode MCSORDestroy(MCSOR *);
PetscErrorCode MCSORApply(MCSOR, Vec, Vec);
PetscErrorCode MCSORSetOmega(MCSOR, PetscReal);
PetscErrorCode MCSORSetSweepType(MCSOR, MatSORType);
PetscErrorCode MCSORGetSweepType(MCSOR, MatSORType *);
PetscErrorCode MCSORGetISColoring(MCSOR, ISColoring *);
PetscErrorCode MCSORGetNumColors(MCSOR, PetscInt *);
PetscErrorCode MCSORBuildLRCCorrection(PetscErrorCode (*det_sor)(void *, Vec, Vec), void *, Mat, Mat, Vec, Mat *);

static PetscErrorCode det_sor(void *ctx, Vec b, Vec y) { return MCSORApply((MCSOR)ctx, b, y); } /* src/mc_sor.c:546-551 */

static PetscErrorCode run(int argc, char **argv)
{
  const PetscInt nx = argc > 1 ? atoi(argv[1]) : 9, n = nx * nx; /* examples/ex5.c:41: a 9 x 9 DMDA */
  const double   omega = argc > 2 ? atof(argv[2]) : 1.0;
  Mat            A;
  MCSOR          mc;
  Vec            x, y, b;
  double         err;
  MatSORType     st;

  PetscCall(ParMGMCInitialize());
  PetscCall(host_assemble(nx, 1.0, &A)); /* MatAssembleShiftedLaplaceFD(da, 1, A), examples/ex5.c:47 */
  PetscCall(MCSORCreate(A, &mc));
  PetscCall(MCSORSetUp(mc));
  if (omega != 1.0) PetscCall(MCSORSetOmega(mc, omega));
  PetscCall(VecCreateSeq(MPI_COMM_SELF, n, &x));
  PetscCall(VecCreateSeq(MPI_COMM_SELF, n, &y));
  PetscCall(VecCreateSeq(MPI_COMM_SELF, n, &b));
  PetscCall(host_fill_uniform(b, 1));
  PetscCall(host_fill_uniform(x, 2));
  PetscCall(VecCopy(x, y));

  /* (1) */
  PetscCall(MCSORSetSweepType(mc, SOR_FORWARD_SWEEP));
  PetscCall(MCSORApply(mc, b, x));
  PetscCall(MCSORSetSweepType(mc, SOR_BACKWARD_SWEEP));
  PetscCall(MCSORApply(mc, b, x));
  PetscCall(MCSORSetSweepType(mc, SOR_SYMMETRIC_SWEEP));
  PetscCall(MCSORGetSweepType(mc, &st));
  PetscCheck(st == SOR_SYMMETRIC_SWEEP, PETSC_COMM_SELF, PETSC_ERR_PLIB, "MCSORGetSweepType");
  PetscCall(MCSORApply(mc, b, y));
  PetscCall(VecAXPY(x, -1, y));
  {
    const PetscScalar *a;
    err = 0;
    PetscCall(VecGetArrayRead(x, &a));
    for (PetscInt i = 0; i < n; ++i) err += a[i] * a[i];
    PetscCall(VecRestoreArrayRead(x, &a));
    err = sqrt(err);
  }
  printf("host_ex5: |forward+backward - symmetric|_2 = %.3g\n", err);
  PetscCheck(fabs(err) < 1e-15, MPI_COMM_WORLD, PETSC_ERR_PLIB, "Forward+Backward sweep is not the same as symmetric sweep"); /* examples/ex5.c:68 */
  PetscCheck(MCSORSetSweepType(mc, SOR_LOCAL_BACKWARD_SWEEP) == PETSC_ERR_SUP, PETSC_COMM_SELF, PETSC_ERR_PLIB, "unsupported sweep types must be refused (src/mc_sor.c:427)");

  /* (4) */
  {
    ISColoring      isc;
    PetscInt        nc, nis;
    IS             *iss;
    const PetscInt *ia, *ja;
    PetscScalar    *va;
    int            *colour;
    PetscCall(MCSORGetNumColors(mc, &nc));
    PetscCall(MCSORGetISColoring(mc, &isc));
    PetscCall(ISColoringGetIS(isc, PETSC_USE_POINTER, &nis, &iss));
    PetscCheck(nis == nc && nc >= 2, PETSC_COMM_SELF, PETSC_ERR_PLIB, "colour counts disagree: %d vs %d", (int)nis, (int)nc);
    PetscCall(PetscMalloc1(n, &colour));
    for (PetscInt r = 0; r < n; ++r) colour[r] = -1;
    for (PetscInt c = 0; c < nis; ++c) {
      const PetscInt *idx;
      PetscInt        len;
      PetscCall(ISGetLocalSize(iss[c], &len));
      PetscCall(ISGetIndices(iss[c], &idx));
      for (PetscInt k = 0; k < len; ++k) colour[idx[k]] = (int)c;
      PetscCall(ISRestoreIndices(iss[c], &idx));
    }
    PetscCall(MatSeqAIJGetCSRAndMemType(A, &ia, &ja, &va, NULL));
    for (PetscInt r = 0; r < n; ++r) {
      PetscCheck(colour[r] >= 0, PETSC_COMM_SELF, PETSC_ERR_PLIB, "row %d has no colour", (int)r);
      for (PetscInt k = ia[r]; k < ia[r + 1]; ++k) PetscCheck(ja[k] == r || colour[ja[k]] != colour[r], PETSC_COMM_SELF, PETSC_ERR_PLIB, "rows %d and %d are coupled and share a colour", (int)r, (int)ja[k]);
    }
    printf("host_ex5: MCSORGetISColoring: %d colours, valid distance-1 colouring\n", (int)nc);
    PetscCall(ISColoringRestoreIS(isc, PETSC_USE_POINTER, &iss));
    PetscCall(ISColoringDestroy(&isc));
    PetscCall(PetscFree(colour));
  }

  /* (2) */
  {
    Mat             Ad, Ao, Ampi;
    MCSOR           mcp;
    const PetscInt *ia, *ja;
    PetscScalar    *va;
    PetscInt       *zero;
    PetscCall(MatSeqAIJGetCSRAndMemType(A, &ia, &ja, &va, NULL));
    PetscCall(MatStubCreateSeqAIJ(MPI_COMM_SELF, n, n, ia, ja, va, &Ad));
    PetscCall(PetscCalloc1(n + 1, &zero));
    PetscCall(MatStubCreateSeqAIJ(MPI_COMM_SELF, n, 0, zero, zero, (double *)zero, &Ao));
    PetscCall(MatStubCreateMPIAIJ(n, n, 0, Ad, Ao, NULL, 0, &Ampi));
    PetscCall(MCSORCreate(Ampi, &mcp));
    PetscCall(MCSORSetUp(mcp));
    if (omega != 1.0) PetscCall(MCSORSetOmega(mcp, omega));
    PetscCall(MCSORSetSweepType(mcp, SOR_SYMMETRIC_SWEEP));
    PetscCall(host_fill_uniform(x, 2));
    PetscCall(MCSORApply(mcp, b, x));
    PetscCall(host_diffnorm(x, y, &err));
    printf("host_ex5: MATMPIAIJ (one rank) vs MATSEQAIJ sweep: rel diff %.3g\n", err);
    PetscCheck(err < 1e-14, PETSC_COMM_SELF, PETSC_ERR_PLIB, "MPIAIJ and SEQAIJ sweeps differ: %g", err);
    PetscCall(MCSORDestroy(&mcp));
    PetscCall(MatDestroy(&Ampi));
    PetscCall(PetscFree(zero));
  }

  /* (3) */
  {
    const PetscInt k = 3;
    Mat            B, Aop, Bb;
    Vec            S, t, u, r;
    MCSOR          mcl;
    PetscScalar   *a;
    PetscCall(MatCreateSeqDense(MPI_COMM_SELF, n, k, NULL, &B));
    PetscCall(MatDenseGetArray(B, &a));
    for (PetscInt j = 0; j < k; ++j)
      for (PetscInt i = 0; i < n; ++i) a[i + j * n] = ((i * 7 + j * 13) % 11 == 0) ? 0.5 + 0.1 * j : 0.0; /* a few "observed" nodes per column */
    PetscCall(MatDenseRestoreArray(B, &a));
    PetscCall(VecCreateSeq(MPI_COMM_SELF, k, &S));
    PetscCall(VecGetArray(S, &a));
    for (PetscInt j = 0; j < k; ++j) a[j] = 10.0 + j; /* S = Sigma^-1 */
    PetscCall(VecRestoreArray(S, &a));
    PetscCall(MatCreateLRC(A, B, S, B, &Aop)); /* examples/ex3.c:111 */
    PetscCall(MCSORCreate(Aop, &mcl));
    PetscCall(MCSORSetUp(mcl));
    /* reference quirk kept by the library: Bb is built at omega = 1 whatever the sampler's omega (src/mc_sor.c:583-593) */
    PetscCall(MCSORSetSweepType(mcl, SOR_FORWARD_SWEEP));
    PetscCall(MCSORSetSweepType(mc, SOR_FORWARD_SWEEP));
    PetscCall(MCSORSetOmega(mc, 1.0));
    PetscCall(MCSORBuildLRCCorrection(det_sor, mc, A, B, S, &Bb));
    PetscCall(host_fill_uniform(x, 5));
    PetscCall(VecCopy(x, y));
    PetscCall(MCSORApply(mcl, b, x));
    PetscCall(MCSORApply(mc, b, y));
    PetscCall(VecCreateSeq(MPI_COMM_SELF, k, &t));
    PetscCall(VecDuplicate(y, &u));
    PetscCall(MatMultTranspose(B, y, t));
    PetscCall(MatMult(Bb, t, u));
    PetscCall(VecAXPY(y, -1.0, u)); /* MCSORPostSOR_LRC, src/mc_sor.c:101-112 */
    PetscCall(host_diffnorm(x, y, &err));
    printf("host_ex5: MATLRC sweep vs sweep on A + MCSORBuildLRCCorrection: rel diff %.3g\n", err);
    PetscCheck(err < 1e-12, PETSC_COMM_SELF, PETSC_ERR_PLIB, "LRC post-correction differs: %g", err);
    /* SOR iteration on the MATLRC operator converges to (A + B S B^T)^-1 b */
    PetscCall(VecZeroEntries(x));
    for (int it = 0; it < 400; ++it) PetscCall(MCSORApply(mcl, b, x));
    PetscCall(VecDuplicate(x, &r));
    PetscCall(MatMult(Aop, x, r));
    PetscCall(host_diffnorm(r, b, &err));
    printf("host_ex5: MATLRC SOR iteration: relative residual %.3g after 400 sweeps\n", err);
    PetscCheck(err < 1e-10, PETSC_COMM_SELF, PETSC_ERR_PLIB, "SOR on the MATLRC operator has not converged: %g", err);
    PetscCall(VecDestroy(&t)); PetscCall(VecDestroy(&u)); PetscCall(VecDestroy(&r));
    PetscCall(MCSORDestroy(&mcl));
    PetscCall(MatDestroy(&Bb));
    PetscCall(MatDestroy(&Aop));
    PetscCall(MatDestroy(&B));
    PetscCall(VecDestroy(&S));
  }

  PetscCall(VecDestroy(&b));
  PetscCall(VecDestroy(&x));
  PetscCall(VecDestroy(&y));
  PetscCall(MCSORDestroy(&mc));
  PetscCall(MatDestroy(&A));
  PetscCall(ParMGMCFinalize());
  printf("host_ex5 ok\n");
  return PETSC_SUCCESS;
}

const char *PetscStubLastError(void);
int main(int argc, char **argv)
{
  PetscErrorCode e = run(argc, argv);
  if (e) fprintf(stderr, "host_ex5 failed (%d): %s\n", e, PetscStubLastError());
  return e ? 1 : 0;
}
