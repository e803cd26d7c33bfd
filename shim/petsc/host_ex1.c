/* host_ex1.c -- a C host program in the shape of the reference's examples/ex1.c, driving the PETSc shim.
 *
 * It uses only what an ex1-style program uses: ParMGMCInitialize, a PC of a ParMGMC type configured by option keys,
 * PCSetSampleCallback, and Richardson iterations through pc->ops->applyrichardson (what KSPSolve_Richardson's fast
 * path calls).  PETSc itself is not in this image, so the PETSc objects come from
 * oracle/petsc_stub (TEST INFRASTRUCTURE); the PC implementation underneath is parmgmc_b200_petsc.c ->
 * libparmgmc_b200.so -> CUDA.  Problem and check are ex1's: 9x9 shifted Laplacian, kappa = 10, b = 1, relative error
 * of the running sample mean against the solve <= 0.02 (examples/ex1.c:83-88, :109, :131-135).
 *
 *   host_ex1 <pc type> <samples> <tolerance> [option value]...
 */
#include <math.h>
#include <petsc.h>

PetscErrorCode ParMGMCInitialize(void);
PetscErrorCode ParMGMCFinalize(void);
PetscErrorCode PCSetSampleCallback(PC, PetscErrorCode (*)(PetscInt, Vec, void *), void *, PetscErrorCode (*)(void *));

typedef struct {
  double  *mean;
  PetscInt n, count;
} MeanCtx;

static PetscErrorCode accumulate(PetscInt it, Vec y, void *ctx)
{
  MeanCtx           *m = ctx;
  const PetscScalar *a;
  (void)it;
  PetscCall(VecGetArrayRead(y, &a));
  m->count++;
  for (PetscInt i = 0; i < m->n; ++i) m->mean[i] += (a[i] - m->mean[i]) / (double)m->count; /* Welford, examples/benchmark/main.cc:151-175 */
  PetscCall(VecRestoreArrayRead(y, &a));
  return PETSC_SUCCESS;
}

/* MatAssembleShiftedLaplaceFD semantics (src/problems.c:14-75) */
static PetscErrorCode assemble(PetscInt nx, double kappa, Mat *A)
{
  const PetscInt n = nx * nx;
  const double   h = 1.0 / (double)((nx - 1) * (nx - 1));
  PetscInt      *ia, *ja, nnz = 0;
  double        *va;
  PetscCall(PetscMalloc1(n + 1, &ia));
  PetscCall(PetscMalloc1(5 * n, &ja));
  PetscCall(PetscMalloc1(5 * n, &va));
  ia[0] = 0;
  for (PetscInt j = 0; j < nx; ++j)
    for (PetscInt i = 0; i < nx; ++i) {
      double d = kappa * kappa;
      if (j > 0) { ja[nnz] = i + nx * (j - 1); va[nnz++] = -h; d += h; }
      if (i > 0) { ja[nnz] = i - 1 + nx * j; va[nnz++] = -h; d += h; }
      const PetscInt dpos = nnz++;
      ja[dpos] = i + nx * j;
      if (i < nx - 1) { ja[nnz] = i + 1 + nx * j; va[nnz++] = -h; d += h; }
      if (j < nx - 1) { ja[nnz] = i + nx * (j + 1); va[nnz++] = -h; d += h; }
      va[dpos] = d;
      ia[i + nx * j + 1] = nnz;
    }
  PetscCall(MatStubCreateSeqAIJ(MPI_COMM_WORLD, n, n, ia, ja, va, A));
  PetscCall(PetscFree(ia));
  PetscCall(PetscFree(ja));
  PetscCall(PetscFree(va));
  return PETSC_SUCCESS;
}

static PetscErrorCode run(int argc, char **argv)
{
  const char *type     = argc > 1 ? argv[1] : "mcgibbs";
  const int   nsamples = argc > 2 ? atoi(argv[2]) : 1000000; /* examples/ex1.c:20: 10^6 Gibbs samples */
  const double tol     = argc > 3 ? atof(argv[3]) : 0.02;        /* examples/ex1.c:135 */
  const PetscInt nx = 9, n = nx * nx;
  Mat         A;
  PC          pc;
  Vec         b, x, w, exact;
  PetscScalar *a;
  PetscInt    outits;
  PCRichardsonConvergedReason reason;
  MeanCtx     m;

  PetscCall(ParMGMCInitialize());
  PetscCall(assemble(nx, 10.0, &A));
  for (int k = 4; k + 1 < argc; k += 2) PetscCall(PetscStubOptionsSet(argv[k], argv[k + 1]));
  PetscCall(PCStubCreate(type, A, &pc)); /* PCCreate + PCSetType + PCSetOperators */
  PetscCall(pc->ops->setfromoptions(pc, NULL));
  PetscCall(pc->ops->setup(pc));
  PetscCall(VecCreateSeq(MPI_COMM_SELF, n, &b));
  PetscCall(VecCreateSeq(MPI_COMM_SELF, n, &x));
  PetscCall(VecCreateSeq(MPI_COMM_SELF, n, &w));
  PetscCall(VecCreateSeq(MPI_COMM_SELF, n, &exact));
  PetscCall(VecGetArray(b, &a));
  for (PetscInt i = 0; i < n; ++i) a[i] = 1.0;
  PetscCall(VecRestoreArray(b, &a));

  /* the mean A^-1 b by plain Gauss-Seidel on the host (stands in for ex1's KSPSolve with a direct solver) */
  {
    const PetscInt *ia, *ja;
    PetscScalar    *va, *e;
    PetscCall(MatSeqAIJGetCSRAndMemType(A, &ia, &ja, &va, NULL));
    PetscCall(VecGetArray(exact, &e));
    for (int it = 0; it < 200; ++it)
      for (PetscInt r = 0; r < n; ++r) {
        double s = 1.0, d = 1.0;
        for (PetscInt k = ia[r]; k < ia[r + 1]; ++k)
          if (ja[k] == r) d = va[k];
          else s -= va[k] * e[ja[k]];
        e[r] = s / d;
      }
    PetscCall(VecRestoreArray(exact, &e));
  }

  /* burn-in (examples/ex1.c:124), then sampling with the mean accumulated in the callback (:127-130) */
  PetscCall(pc->ops->applyrichardson(pc, b, x, w, 0, 0, 0, 1000, PETSC_FALSE, &outits, &reason));
  m.n = n;
  m.count = 0;
  PetscCall(PetscCalloc1(n, &m.mean));
  PetscCall(PCSetSampleCallback(pc, accumulate, &m, NULL));
  PetscCall(pc->ops->applyrichardson(pc, b, x, w, 0, 0, 0, nsamples, PETSC_FALSE, &outits, &reason));
  PetscCheck(outits == nsamples && reason == PCRICHARDSON_CONVERGED_ITS && m.count == nsamples, PETSC_COMM_SELF, PETSC_ERR_PLIB, "unexpected iteration count");

  double num = 0, den = 0;
  PetscCall(VecGetArray(exact, &a));
  for (PetscInt i = 0; i < n; ++i) {
    num += (m.mean[i] - a[i]) * (m.mean[i] - a[i]);
    den += a[i] * a[i];
  }
  PetscCall(VecRestoreArray(exact, &a));
  const double rel = sqrt(num / den);
  if (pc->ops->view) PetscCall(pc->ops->view(pc, NULL));
  printf("host_ex1 %s: %d samples, relative mean error %.4g (tolerance %g)\n", type, nsamples, rel, tol);
  PetscCheck(rel <= tol, PETSC_COMM_SELF, PETSC_ERR_PLIB, "sample mean has not converged: %g", rel); /* examples/ex1.c:135 */
  PetscCall(PCStubDestroy(&pc));
  PetscCall(VecDestroy(&b)); PetscCall(VecDestroy(&x)); PetscCall(VecDestroy(&w)); PetscCall(VecDestroy(&exact));
  PetscCall(MatDestroy(&A));
  PetscCall(ParMGMCFinalize());
  return PETSC_SUCCESS;
}

const char *PetscStubLastError(void);
int main(int argc, char **argv)
{
  PetscErrorCode e = run(argc, argv);
  if (e) fprintf(stderr, "host_ex1 failed (%d): %s\n", e, PetscStubLastError());
  return e ? 1 : 0;
}
