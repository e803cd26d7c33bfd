/* host_api.c -- exercises, from C and through the PETSc shim, the rest of the reference's exported API for the sampling path
 * (include/parmgmc/parmgmc.h:40-44, pc/pc_gamgmc.h:15-18, pc/pc_chols.h:15-16, pc/woodbury.h:15-17, iact.h:14-15) on the GPU.
 * PETSc objects come from oracle/petsc_stub (TEST INFRASTRUCTURE).
 *
 *   (1) ParMGMCGetPetscRandom / VecSetRandomStandardNormal: reference-counted generator, N(0,1) moments, seed-reproducible
 *   (2) PC "woodbury" on a MATLRC operator, inner PCs by option keys (examples/benchmark/lshape.opts shape) and by
 *       PCWoodburySetSampler / PCWoodburySetSolver objects: sample mean -> (A + B S B^T)^-1 b (the acceptance of examples/ex4.c)
 *   (3) PC "mcgibbs" directly on the MATLRC operator (src/pc_mcgibbs.c:236-244): same posterior mean
 *   (4) PC "gamgmc": PCGAMGMCSetLevels, PCGAMGMCGetInternalPC, PCMGGetLevels through the composed "PCMGGetLevels_C",
 *       PCGAMGMCSetInternalPC adopting the level count of another gamgmc PC
 *   (5) PC "cholsampler": PCCholSamplerSetIsCoarseGAMG, callback only legal inside PCPreSolve / PCPostSolve for PCApply
 *       (src/pc_chols.c:267), sample indices restart at every presolve (:352-353)
 *   (6) IACT / Autocorrelation of white noise: tau ~ 1, acf[0] = 1
 */
#include "host_common.h"

PetscErrorCode ParMGMCInitialize(void);
PetscErrorCode ParMGMCFinalize(void);
PetscErrorCode ParMGMCGetPetscRandom(PetscRandom *);
PetscErrorCode VecSetRandomStandardNormal(Vec, PetscRandom);
PetscErrorCode PCSetSampleCallback(PC, PetscErrorCode (*)(PetscInt, Vec, void *), void *, PetscErrorCode (*)(void *));
PetscErrorCode PCGAMGMCSetLevels(PC, PetscInt);
PetscErrorCode PCGAMGMCGetInternalPC(PC, PC *);
PetscErrorCode PCGAMGMCSetInternalPC(PC, PC);
PetscErrorCode PCCholSamplerSetIsCoarseGAMG(PC, PetscBool);
PetscErrorCode PCWoodburySetSolver(PC, PC);
PetscErrorCode PCWoodburySetSampler(PC, PC);
PetscErrorCode Autocorrelation(PetscInt, const PetscScalar *, PetscScalar **);
PetscErrorCode IACT(PetscInt, const PetscScalar *, PetscScalar *, PetscScalar **, PetscBool *);

typedef struct {
  double  *mean;
  PetscInt n, count, last_it;
} MeanCtx;

static PetscErrorCode accumulate(PetscInt it, Vec y, void *ctx)
{
  MeanCtx           *m = ctx;
  const PetscScalar *a;
  PetscCall(VecGetArrayRead(y, &a));
  m->count++;
  m->last_it = it;
  for (PetscInt i = 0; i < m->n; ++i) m->mean[i] += (a[i] - m->mean[i]) / (double)m->count;
  PetscCall(VecRestoreArrayRead(y, &a));
  return PETSC_SUCCESS;
}

/* PCMGGetLevels as PETSc resolves it on a PC that composes "PCMGGetLevels_C" (src/pc_gamgmc.c:413) */
static PetscErrorCode HostPCMGGetLevels(PC pc, PetscInt *levels)
{
  PetscUseMethod((PetscObject)pc, "PCMGGetLevels_C", (PC, PetscInt *), (pc, levels));
  return PETSC_SUCCESS;
}

static PetscErrorCode sample_mean_check(PC pc, Vec b, Vec exact, PetscInt n, int nsamples, double tol, const char *what)
{
  Vec                         x, w;
  MeanCtx                     m = {0};
  PetscInt                    outits;
  PCRichardsonConvergedReason reason;
  double                      num = 0, den = 0;
  const PetscScalar          *e;

  PetscCall(VecCreateSeq(MPI_COMM_SELF, n, &x));
  PetscCall(VecCreateSeq(MPI_COMM_SELF, n, &w));
  PetscCall(pc->ops->applyrichardson(pc, b, x, w, 0, 0, 0, 500, PETSC_FALSE, &outits, &reason)); /* burn-in */
  m.n = n;
  PetscCall(PetscCalloc1(n, &m.mean));
  PetscCall(PCSetSampleCallback(pc, accumulate, &m, NULL));
  PetscCall(pc->ops->applyrichardson(pc, b, x, w, 0, 0, 0, nsamples, PETSC_FALSE, &outits, &reason));
  PetscCheck(outits == nsamples && reason == PCRICHARDSON_CONVERGED_ITS && m.count == nsamples && m.last_it == nsamples - 1, PETSC_COMM_SELF, PETSC_ERR_PLIB, "%s: unexpected iteration bookkeeping", what);
  PetscCall(VecGetArrayRead(exact, &e));
  for (PetscInt i = 0; i < n; ++i) {
    num += (m.mean[i] - e[i]) * (m.mean[i] - e[i]);
    den += e[i] * e[i];
  }
  PetscCall(VecRestoreArrayRead(exact, &e));
  printf("host_api: %s: %d samples, relative mean error %.4g (tolerance %g)\n", what, nsamples, sqrt(num / den), tol);
  PetscCheck(sqrt(num / den) <= tol, PETSC_COMM_SELF, PETSC_ERR_PLIB, "%s: sample mean has not converged: %g", what, sqrt(num / den));
  PetscCall(PetscFree(m.mean));
  PetscCall(VecDestroy(&x));
  PetscCall(VecDestroy(&w));
  return PETSC_SUCCESS;
}

static PetscErrorCode run(int argc, char **argv)
{
  const int      nsamples = argc > 1 ? atoi(argv[1]) : 100000;
  const double   tol      = argc > 2 ? atof(argv[2]) : 0.05;
  const PetscInt nx = 9, n = nx * nx, k = 3;
  Mat            A, B, Aop;
  Vec            S, b, exact;
  PetscScalar   *a;

  PetscCall(ParMGMCInitialize());

  /* (1) */
  {
    PetscRandom r1, r2;
    Vec         z, z2;
    double      m1 = 0, m2 = 0, m4 = 0;
    const PetscInt nz = 200001; /* odd: the last Box-Muller pair is half used (src/parmgmc.c:110) */
    PetscCall(ParMGMCGetPetscRandom(&r1));
    PetscCall(ParMGMCGetPetscRandom(&r2));
    PetscCheck(r1 == r2, PETSC_COMM_SELF, PETSC_ERR_PLIB, "ParMGMCGetPetscRandom must hand out one object per process");
    PetscCall(PetscRandomDestroy(&r2)); /* drops only the caller's reference (src/parmgmc.c:63-66) */
    PetscCall(PetscRandomSetSeed(r1, 1234));
    PetscCall(VecCreateSeq(MPI_COMM_SELF, nz, &z));
    PetscCall(VecCreateSeq(MPI_COMM_SELF, nz, &z2));
    PetscCall(VecSetRandomStandardNormal(z, r1));
    PetscCall(VecSetRandomStandardNormal(z2, r1));
    PetscCall(VecGetArray(z, &a));
    for (PetscInt i = 0; i < nz; ++i) { m1 += a[i]; m2 += a[i] * a[i]; m4 += a[i] * a[i] * a[i] * a[i]; }
    PetscCall(VecRestoreArray(z, &a));
    m1 /= nz; m2 /= nz; m4 /= nz;
    double d;
    PetscCall(host_diffnorm(z, z2, &d));
    printf("host_api: VecSetRandomStandardNormal: mean %.4f var %.4f kurtosis %.4f, consecutive calls differ by %.3g\n", m1, m2, m4 / (m2 * m2), d);
    PetscCheck(fabs(m1) < 0.01 && fabs(m2 - 1) < 0.02 && fabs(m4 / (m2 * m2) - 3) < 0.1 && d > 1.0, PETSC_COMM_SELF, PETSC_ERR_PLIB, "VecSetRandomStandardNormal moments");
    PetscCall(VecDestroy(&z)); PetscCall(VecDestroy(&z2));
    PetscCall(PetscRandomDestroy(&r1));
  }

  /* the MATLRC operator of (2), (3): A + B S B^T, examples/ex4.c / lshape.opts in miniature */
  PetscCall(host_assemble(nx, 10.0, &A));
  PetscCall(MatCreateSeqDense(MPI_COMM_SELF, n, k, NULL, &B));
  PetscCall(MatDenseGetArray(B, &a));
  for (PetscInt j = 0; j < k; ++j)
    for (PetscInt i = 0; i < n; ++i) a[i + j * n] = ((i * 7 + j * 13) % 11 == 0) ? 0.5 + 0.1 * j : 0.0;
  PetscCall(MatDenseRestoreArray(B, &a));
  PetscCall(VecCreateSeq(MPI_COMM_SELF, k, &S));
  PetscCall(VecGetArray(S, &a));
  for (PetscInt j = 0; j < k; ++j) a[j] = 200.0 + 50.0 * j;
  PetscCall(VecRestoreArray(S, &a));
  PetscCall(MatCreateLRC(A, B, S, B, &Aop));
  PetscCall(VecCreateSeq(MPI_COMM_SELF, n, &b));
  PetscCall(VecCreateSeq(MPI_COMM_SELF, n, &exact));
  PetscCall(VecGetArray(b, &a));
  for (PetscInt i = 0; i < n; ++i) a[i] = 1.0 + 0.01 * i;
  PetscCall(VecRestoreArray(b, &a));
  { /* exact posterior mean: dense solve with the stub's KSP (stands in for ex4's direct solve) */
    KSP ksp;
    Mat Ad, BS, BSBt;
    PetscCall(MatConvert(A, MATDENSE, MAT_INITIAL_MATRIX, &Ad));
    PetscCall(MatDuplicate(B, MAT_COPY_VALUES, &BS));
    { /* BS = B diag(S) */
      PetscScalar       *bs;
      const PetscScalar *sv;
      PetscCall(MatDenseGetArray(BS, &bs));
      PetscCall(VecGetArrayRead(S, &sv));
      for (PetscInt j = 0; j < k; ++j)
        for (PetscInt i = 0; i < n; ++i) bs[i + j * n] *= sv[j];
      PetscCall(VecRestoreArrayRead(S, &sv));
      PetscCall(MatDenseRestoreArray(BS, &bs));
    }
    PetscCall(MatMatTransposeMult(BS, B, MAT_INITIAL_MATRIX, 1, &BSBt));
    PetscCall(MatAXPY(Ad, 1.0, BSBt, DIFFERENT_NONZERO_PATTERN));
    PetscCall(KSPCreate(MPI_COMM_SELF, &ksp));
    PetscCall(KSPSetOperators(ksp, Ad, Ad));
    PetscCall(KSPSolve(ksp, b, exact));
    PetscCall(KSPDestroy(&ksp));
    PetscCall(MatDestroy(&Ad)); PetscCall(MatDestroy(&BS)); PetscCall(MatDestroy(&BSBt));
  }

  /* (2a) inner PCs by option keys */
  {
    PC pc;
    PetscCall(PetscStubOptionsSet("-pc_woodbury_sampler", "mcgibbs"));
    PetscCall(PetscStubOptionsSet("-pc_woodbury_solver", "cholesky"));
    PetscCall(PetscStubOptionsSet("-pc_woodbury_sampler_pc_mcgibbs_symmetric", ""));
    PetscCall(PCStubCreate("woodbury", Aop, &pc));
    PetscCall(pc->ops->setfromoptions(pc, NULL));
    PetscCall(pc->ops->setup(pc));
    PetscCall(sample_mean_check(pc, b, exact, n, nsamples, tol, "woodbury (option keys: mcgibbs symmetric / cholesky)"));
    PetscCall(PCStubDestroy(&pc));
    PetscCall(PetscStubOptionsClear());
  }
  /* (2b) inner PCs as objects (src/woodbury.c:188-214) */
  {
    PC pc, sampler, solver;
    PetscCall(PCStubCreate("woodbury", Aop, &pc));
    PetscCall(PCStubCreate("sorgibbs", A, &sampler));
    PetscCall(PCStubCreate("cholsampler", A, &solver));
    PetscCall(PCWoodburySetSampler(pc, sampler));
    PetscCall(PCWoodburySetSolver(pc, solver));
    PetscCall(PCDestroy(&sampler)); /* the woodbury PC holds its own references */
    PetscCall(PCDestroy(&solver));
    PetscCall(pc->ops->setfromoptions(pc, NULL));
    PetscCall(pc->ops->setup(pc));
    PetscCall(sample_mean_check(pc, b, exact, n, nsamples, tol, "woodbury (PCWoodburySetSampler sorgibbs / PCWoodburySetSolver cholsampler)"));
    PetscCall(PCStubDestroy(&pc));
  }
  /* (3) */
  {
    PC pc;
    PetscCall(PCStubCreate("mcgibbs", Aop, &pc));
    PetscCall(pc->ops->setfromoptions(pc, NULL));
    PetscCall(pc->ops->setup(pc));
    PetscCall(sample_mean_check(pc, b, exact, n, nsamples, tol, "mcgibbs on MATLRC"));
    PetscCall(PCStubDestroy(&pc));
  }

  /* (4) */
  {
    PC       pc, pc2, inner;
    PetscInt levels = 0;
    PetscCall(PetscStubOptionsSet("-pc_b200_grid", "9,9"));
    PetscCall(PCStubCreate("gamgmc", A, &pc));
    PetscCall(PCGAMGMCSetLevels(pc, 3));
    PetscCall(pc->ops->setfromoptions(pc, NULL));
    PetscCall(PCGAMGMCGetInternalPC(pc, &inner));
    PetscCall(HostPCMGGetLevels(inner, &levels));
    PetscCheck(levels == 3, PETSC_COMM_SELF, PETSC_ERR_PLIB, "PCMGGetLevels on the internal PC: %d", (int)levels);
    PetscCall(pc->ops->setup(pc));
    PetscCall(HostPCMGGetLevels(pc, &levels));
    PetscCheck(levels == 3, PETSC_COMM_SELF, PETSC_ERR_PLIB, "PCMGGetLevels after set-up: %d", (int)levels);
    PetscCall(PCStubCreate("gamgmc", A, &pc2));
    PetscCall(PCGAMGMCSetInternalPC(pc2, pc)); /* adopts pc's configuration; pc2 keeps a reference (src/pc_gamgmc.c:125-133) */
    PetscCall(PCGAMGMCGetInternalPC(pc2, &inner));
    PetscCheck(inner == pc, PETSC_COMM_SELF, PETSC_ERR_PLIB, "PCGAMGMCGetInternalPC must return the PC that was set");
    PetscCall(HostPCMGGetLevels(pc2, &levels));
    PetscCheck(levels == 3, PETSC_COMM_SELF, PETSC_ERR_PLIB, "PCGAMGMCSetInternalPC did not adopt the level count: %d", (int)levels);
    PetscCall(pc2->ops->setfromoptions(pc2, NULL));
    PetscCall(pc2->ops->setup(pc2));
    {
      Vec ex2; /* gamgmc samples N(A^-1 b, A^-1): mean by the stub's dense solve */
      KSP ksp;
      Mat Ad;
      PetscCall(VecDuplicate(b, &ex2));
      PetscCall(MatConvert(A, MATDENSE, MAT_INITIAL_MATRIX, &Ad));
      PetscCall(KSPCreate(MPI_COMM_SELF, &ksp));
      PetscCall(KSPSetOperators(ksp, Ad, Ad));
      PetscCall(KSPSolve(ksp, b, ex2));
      PetscCall(sample_mean_check(pc2, b, ex2, n, nsamples, tol, "gamgmc, 3 levels adopted through PCGAMGMCSetInternalPC"));
      PetscCall(KSPDestroy(&ksp));
      PetscCall(MatDestroy(&Ad));
      PetscCall(VecDestroy(&ex2));
    }
    printf("host_api: PCGAMGMCGet/SetInternalPC + PCMGGetLevels_C: %d levels\n", (int)levels);
    PetscCall(PCStubDestroy(&pc2));
    PetscCall(PCDestroy(&pc)); /* drops the creator's reference; pc2 dropped its own in PCDestroy_B200 */
    PetscCall(PetscStubOptionsClear());
  }

  /* (5) */
  {
    PC      pc;
    Vec     x;
    MeanCtx m = {0};
    PetscCall(PCStubCreate("cholsampler", A, &pc));
    PetscCall(PCCholSamplerSetIsCoarseGAMG(pc, PETSC_TRUE));
    PetscCall(PCCholSamplerSetIsCoarseGAMG(pc, PETSC_FALSE));
    PetscCall(pc->ops->setfromoptions(pc, NULL));
    PetscCall(pc->ops->setup(pc));
    PetscCall(VecCreateSeq(MPI_COMM_SELF, n, &x));
    m.n = n;
    PetscCall(PetscCalloc1(n, &m.mean));
    PetscCall(pc->ops->apply(pc, b, x)); /* no callback: legal anywhere */
    PetscCall(PCSetSampleCallback(pc, accumulate, &m, NULL));
    PetscCheck(pc->ops->apply(pc, b, x) == PETSC_ERR_SUP, PETSC_COMM_SELF, PETSC_ERR_PLIB, "PCApply with a callback outside a solve must be refused (src/pc_chols.c:267)");
    for (int solve = 0; solve < 2; ++solve) {
      PetscCall(pc->ops->presolve(pc, NULL, b, x));
      for (int it = 0; it < 5; ++it) PetscCall(pc->ops->apply(pc, b, x));
      PetscCall(pc->ops->postsolve(pc, NULL, b, x));
      PetscCheck(m.count == 5 * (solve + 1) && m.last_it == 4, PETSC_COMM_SELF, PETSC_ERR_PLIB, "cholsampler in-solve callback bookkeeping: count %d last %d", (int)m.count, (int)m.last_it);
    }
    printf("host_api: cholsampler presolve / postsolve callback bookkeeping ok\n");
    PetscCall(PetscFree(m.mean));
    PetscCall(VecDestroy(&x));
    PetscCall(PCStubDestroy(&pc));
  }

  /* (6) */
  {
    const PetscInt nq = 4096;
    PetscRandom    r;
    Vec            q;
    PetscScalar   *acf, *acf2, tau;
    PetscBool      valid;
    const PetscScalar *qa;
    PetscCall(ParMGMCGetPetscRandom(&r));
    PetscCall(VecCreateSeq(MPI_COMM_SELF, nq, &q));
    PetscCall(VecSetRandomStandardNormal(q, r));
    PetscCall(VecGetArrayRead(q, &qa));
    PetscCall(Autocorrelation(nq, qa, &acf));
    PetscCall(IACT(nq, qa, &tau, &acf2, &valid));
    PetscCall(VecRestoreArrayRead(q, &qa));
    printf("host_api: IACT of white noise: tau %.3f valid %d acf[0] %.6f acf[1] %.4f\n", tau, (int)valid, acf[0], acf[1]);
    PetscCheck(fabs(acf[0] - 1.0) < 1e-12 && fabs(acf[1]) < 0.08 && fabs(acf2[1] - acf[1]) < 1e-12 && tau > 0.7 && tau < 1.4, PETSC_COMM_SELF, PETSC_ERR_PLIB, "IACT / Autocorrelation of white noise");
    PetscCall(PetscFree(acf));
    PetscCall(PetscFree(acf2));
    PetscCall(VecDestroy(&q));
    PetscCall(PetscRandomDestroy(&r));
  }

  PetscCall(VecDestroy(&b)); PetscCall(VecDestroy(&exact));
  PetscCall(MatDestroy(&Aop)); PetscCall(MatDestroy(&B)); PetscCall(VecDestroy(&S)); PetscCall(MatDestroy(&A));
  PetscCall(ParMGMCFinalize());
  printf("host_api ok\n");
  return PETSC_SUCCESS;
}

const char *PetscStubLastError(void);
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>
static void on_segv(int sig)
{
  void *bt[32];
  const int n = backtrace(bt, 32);
  (void)sig;
  backtrace_symbols_fd(bt, n, 2);
  _exit(139);
}
int main(int argc, char **argv)
{
  setvbuf(stdout, NULL, _IONBF, 0);
  signal(SIGSEGV, on_segv);
  PetscErrorCode e = run(argc, argv);
  if (e) fprintf(stderr, "host_api failed (%d): %s\n", e, PetscStubLastError());
  return e ? 1 : 0;
}
