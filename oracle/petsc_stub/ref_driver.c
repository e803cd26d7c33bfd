/* ref_driver.c -- plain-C entry points (ctypes) that drive the reference's OWN compiled code (TEST INFRASTRUCTURE ONLY).
 *
 * Everything arithmetic below happens inside /root/reference/src/{mc_sor.c, pc_mcgibbs.c, parmgmc.c}, compiled
 * unmodified against petsc_stub.h.  This file only builds the containers (CSR -> Mat, arrays -> Vec), picks the
 * number of emulated ranks and calls the reference's public API:
 *   MCSORCreate / MCSORSetUp / MCSORSetOmega / MCSORSetSweepType / MCSORApply      (include/parmgmc/mc_sor.h:17-30)
 *   PCCreate_MulticolorGibbs -> setfromoptions / setup / applyrichardson           (src/pc_mcgibbs.c:305-326)
 *   ParMGMCGetPetscRandom / VecSetRandomStandardNormal                              (src/parmgmc.c:56-116)
 */
#include <pthread.h>

#include "parmgmc/mc_sor.h"
#include "parmgmc/parmgmc.h"
#include "parmgmc/pc/pc_mcgibbs.h"
#include "parmgmc/iact.h"
#include "parmgmc/stats.h"
#include "petsc_stub.h"

const char *PetscStubLastError(void);

/* the other PC types parmgmc.c registers are not compiled into this library */
#define ABSENT_PC(fn)                                                                                   \
  PetscErrorCode fn(PC pc)                                                                              \
  {                                                                                                     \
    (void)pc;                                                                                           \
    return PetscStubError(PETSC_ERR_SUP, __FILE__, __LINE__, #fn " is not part of oracle/_ref");        \
  }
ABSENT_PC(PCCreate_GAMGMC)  /* wraps PETSc's PCMG / PCGAMG: nothing to run without PETSc */
ABSENT_PC(PCCreate_PARSOR)  /* raw MPI point-to-point: out of scope */
PetscErrorCode PCPARSORApplySOR(PC pc, Vec b, PetscInt its, PetscBool zero, Vec y)
{
  (void)pc; (void)b; (void)its; (void)zero; (void)y;
  return PetscStubError(PETSC_ERR_SUP, __FILE__, __LINE__, "PCPARSOR is not part of oracle/_ref");
}

const char *ref_last_error(void) { return PetscStubLastError(); }

/* MCSORApply on one rank.  ncolors <= 1 / color == NULL: the reference's own 1-rank colouring (all rows colour 0,
 * mc_sor.c:397-410).  Otherwise the colouring is injected through the reference's multi-rank branch
 * (MatCreateISColoring_AIJ -> MatColoringApply, mc_sor.c:383-395) by telling it the communicator has two ranks while
 * the matrix stays SEQAIJ, so the loop that runs is still MCSORApply_SEQAIJ (mc_sor.c:241-296). */
int ref_mcsor_seq(int n, const int *rowptr, const int *col, const double *val, int ncolors, const unsigned short *color, double omega, int type, int nsweeps, const double *b, double *y)
{
  Mat   A;
  MCSOR mc;
  Vec   vb, vy;
  PetscStubWorldBegin(color && ncolors > 1 ? 2 : 1);
  PetscCall(MatStubCreateSeqAIJ(MPI_COMM_WORLD, n, n, rowptr, col, val, &A));
  if (color && ncolors > 1) PetscCall(MatStubInjectColoring(A, ncolors, color));
  PetscCall(MCSORCreate(A, &mc));
  PetscCall(MCSORSetUp(mc));
  PetscCall(MCSORSetOmega(mc, omega));
  PetscCall(MCSORSetSweepType(mc, (MatSORType)type));
  PetscCall(VecStubCreate(MPI_COMM_WORLD, n, n, 0, (double *)b, &vb));
  PetscCall(VecStubCreate(MPI_COMM_WORLD, n, n, 0, y, &vy));
  for (int s = 0; s < nsweeps; ++s) PetscCall(MCSORApply(mc, vb, vy));
  PetscCall(VecDestroy(&vb));
  PetscCall(VecDestroy(&vy));
  PetscCall(MCSORDestroy(&mc));
  PetscCall(MatDestroy(&A));
  PetscStubWorldEnd();
  return 0;
}

int ref_mcsor_num_colors(int n, const int *rowptr, const int *col, const double *val, int *out)
{
  Mat   A;
  MCSOR mc;
  PetscStubWorldBegin(1);
  PetscCall(MatStubCreateSeqAIJ(MPI_COMM_WORLD, n, n, rowptr, col, val, &A));
  PetscCall(MCSORCreate(A, &mc));
  PetscCall(MCSORSetUp(mc));
  PetscCall(MCSORGetNumColors(mc, out));
  PetscCall(MCSORDestroy(&mc));
  PetscCall(MatDestroy(&A));
  PetscStubWorldEnd();
  return 0;
}

/* ---- MCSORApply_MPIAIJ (mc_sor.c:298-381) on `nranks` emulated ranks (threads) ------------------------------------- */
typedef struct {
  int                   rank, nranks, n, ncolors, type, nsweeps, rc;
  const int            *rowptr, *col, *rowstart;
  const double         *val, *b;
  const unsigned short *color;
  double                omega, *y;
} rank_job;

/* split the rows [r0, r1) of the global CSR the way MPIAIJ stores them: diagonal block with local column ids,
 * off-diagonal block with compressed ids whose global ids (ascending) form colmap */
static int build_rank_matrix(const rank_job *j, Mat *out)
{
  const int r0 = j->rowstart[j->rank], r1 = j->rowstart[j->rank + 1], m = r1 - r0;
  int      *mark = calloc((size_t)j->n, sizeof(int)), ncm = 0;
  for (int r = r0; r < r1; ++r)
    for (int k = j->rowptr[r]; k < j->rowptr[r + 1]; ++k) {
      const int c = j->col[k];
      if ((c < r0 || c >= r1) && !mark[c]) mark[c] = 1;
    }
  int *colmap = malloc(sizeof(int) * (size_t)(j->n > 0 ? j->n : 1));
  for (int c = 0; c < j->n; ++c)
    if (mark[c]) {
      mark[c]      = ncm + 1;
      colmap[ncm++] = c;
    }
  const int nnz = j->rowptr[r1] - j->rowptr[r0];
  int      *di = calloc((size_t)m + 1, sizeof(int)), *oi = calloc((size_t)m + 1, sizeof(int));
  int      *dj = malloc(sizeof(int) * (size_t)(nnz + 1)), *oj = malloc(sizeof(int) * (size_t)(nnz + 1));
  double   *da = malloc(sizeof(double) * (size_t)(nnz + 1)), *oa = malloc(sizeof(double) * (size_t)(nnz + 1));
  int       nd = 0, no = 0;
  for (int r = r0; r < r1; ++r) {
    for (int k = j->rowptr[r]; k < j->rowptr[r + 1]; ++k) {
      const int c = j->col[k];
      if (c >= r0 && c < r1) {
        dj[nd]   = c - r0;
        da[nd++] = j->val[k];
      } else {
        oj[no]   = mark[c] - 1;
        oa[no++] = j->val[k];
      }
    }
    di[r - r0 + 1] = nd;
    oi[r - r0 + 1] = no;
  }
  Mat Ad, Ao;
  PetscCall(MatStubCreateSeqAIJ(MPI_COMM_SELF, m, m, di, dj, da, &Ad));
  PetscCall(MatStubCreateSeqAIJ(MPI_COMM_SELF, m, ncm, oi, oj, oa, &Ao));
  PetscCall(MatStubCreateMPIAIJ(m, j->n, r0, Ad, Ao, colmap, ncm, out));
  PetscCall(MatStubInjectColoring(*out, j->ncolors, j->color + r0));
  free(mark); free(colmap); free(di); free(oi); free(dj); free(oj); free(da); free(oa);
  return 0;
}

static int rank_main(rank_job *j)
{
  const int r0 = j->rowstart[j->rank], m = j->rowstart[j->rank + 1] - r0;
  Mat       A;
  MCSOR     mc;
  Vec       vb, vy;
  PetscStubSetRank(j->rank);
  PetscCall(build_rank_matrix(j, &A));
  PetscCall(MCSORCreate(A, &mc));
  PetscCall(MCSORSetUp(mc));
  PetscCall(MCSORSetOmega(mc, j->omega));
  PetscCall(MCSORSetSweepType(mc, (MatSORType)j->type));
  PetscCall(VecStubCreate(MPI_COMM_WORLD, m, j->n, r0, (double *)j->b + r0, &vb));
  PetscCall(VecStubCreate(MPI_COMM_WORLD, m, j->n, r0, j->y + r0, &vy)); /* a slice of the one shared array */
  for (int s = 0; s < j->nsweeps; ++s) PetscCall(MCSORApply(mc, vb, vy));
  PetscStubBarrier();
  PetscCall(VecDestroy(&vb));
  PetscCall(VecDestroy(&vy));
  PetscCall(MCSORDestroy(&mc));
  PetscCall(MatDestroy(&A));
  return 0;
}
static void *rank_thread(void *p)
{
  rank_job *j = p;
  j->rc       = rank_main(j);
  return NULL;
}

int ref_mcsor_mpi(int n, const int *rowptr, const int *col, const double *val, int nranks, const int *rowstart, int ncolors, const unsigned short *color, double omega, int type, int nsweeps, const double *b, double *y)
{
  if (nranks < 2) return PetscStubError(PETSC_ERR_ARG_WRONG, __FILE__, __LINE__, "ref_mcsor_mpi needs at least two ranks");
  PetscStubWorldBegin(nranks);
  pthread_t *th = malloc(sizeof(pthread_t) * (size_t)nranks);
  rank_job  *jb = malloc(sizeof(rank_job) * (size_t)nranks);
  for (int r = 0; r < nranks; ++r) {
    jb[r] = (rank_job){r, nranks, n, ncolors, type, nsweeps, 0, rowptr, col, rowstart, val, b, color, omega, y};
    pthread_create(&th[r], NULL, rank_thread, &jb[r]);
  }
  int rc = 0;
  for (int r = 0; r < nranks; ++r) {
    pthread_join(th[r], NULL);
    if (jb[r].rc) rc = jb[r].rc;
  }
  free(th);
  free(jb);
  PetscStubWorldEnd();
  return rc;
}

/* ---- PCMCGIBBS on one rank: setfromoptions + setup + applyrichardson, noise from the library's global PetscRandom ----- */
typedef int (*ref_sample_cb)(int it, const double *y, int n, void *ctx);
static ref_sample_cb g_cb;
static void         *g_cbctx;
static PetscErrorCode cb_tramp(PetscInt it, Vec y, void *ctx)
{
  (void)ctx;
  return g_cb ? g_cb(it, y->a, y->n, g_cbctx) : 0;
}

int ref_mcgibbs_richardson(int n, const int *rowptr, const int *col, const double *val, int ncolors, const unsigned short *color, const char *omega_opt, const char *sweep_opt, long long seed, int its, const double *b, double *y, ref_sample_cb cb, void *cbctx)
{
  Mat                         A;
  PC                          pc;
  Vec                         vb = NULL, vy, vw;
  PetscRandom                 pr;
  PetscInt                    outits;
  PCRichardsonConvergedReason reason;
  PetscStubWorldBegin(color && ncolors > 1 ? 2 : 1);
  PetscCall(ParMGMCInitialize());
  PetscCall(ParMGMCGetPetscRandom(&pr)); /* the process-global stream (src/parmgmc.c:38-68) */
  PetscCall(PetscRandomSetSeed(pr, seed));
  PetscCall(PetscRandomSeed(pr));
  PetscCall(PetscRandomDestroy(&pr));
  PetscCall(MatStubCreateSeqAIJ(MPI_COMM_WORLD, n, n, rowptr, col, val, &A));
  if (color && ncolors > 1) PetscCall(MatStubInjectColoring(A, ncolors, color));
  PetscCall(PetscStubOptionsClear());
  if (omega_opt && omega_opt[0]) PetscCall(PetscStubOptionsSet("-pc_mcgibbs_omega", omega_opt));
  if (sweep_opt && sweep_opt[0]) PetscCall(PetscStubOptionsSet(sweep_opt, ""));
  PetscCall(PCStubCreate(PCMCGIBBS, A, &pc));
  PetscCall(pc->ops->setfromoptions(pc, NULL));
  PetscCall(pc->ops->setup(pc));
  pc->setupcalled = PETSC_TRUE;
  g_cb            = cb;
  g_cbctx         = cbctx;
  if (cb) PetscCall(PCSetSampleCallback(pc, cb_tramp, NULL, NULL));
  if (b) PetscCall(VecStubCreate(MPI_COMM_WORLD, n, n, 0, (double *)b, &vb));
  PetscCall(VecStubCreate(MPI_COMM_WORLD, n, n, 0, y, &vy));
  PetscCall(VecStubCreate(MPI_COMM_WORLD, n, n, 0, NULL, &vw));
  PetscCall(pc->ops->applyrichardson(pc, vb, vy, vw, 0, 0, 0, its, PETSC_FALSE, &outits, &reason));
  if (outits != its || reason != PCRICHARDSON_CONVERGED_ITS) return PetscStubError(PETSC_ERR_PLIB, __FILE__, __LINE__, "unexpected outits/reason");
  if (vb) PetscCall(VecDestroy(&vb));
  PetscCall(VecDestroy(&vy));
  PetscCall(VecDestroy(&vw));
  PetscCall(PCStubDestroy(&pc));
  PetscCall(MatDestroy(&A));
  PetscCall(PetscStubOptionsClear());
  PetscCall(ParMGMCFinalize());
  PetscStubWorldEnd();
  return 0;
}

/* VecSetRandomStandardNormal (src/parmgmc.c:70-116, Box-Muller branch) on the global stream seeded with `seed`; ncalls fills of n */
int ref_normal_fill(long long seed, int n, int ncalls, double *out)
{
  PetscRandom pr;
  Vec         v;
  PetscCall(ParMGMCGetPetscRandom(&pr));
  PetscCall(PetscRandomSetSeed(pr, seed));
  PetscCall(PetscRandomSeed(pr));
  for (int c = 0; c < ncalls; ++c) {
    PetscCall(VecStubCreate(MPI_COMM_SELF, n, n, 0, out + (size_t)c * (size_t)n, &v));
    PetscCall(VecSetRandomStandardNormal(v, pr));
    PetscCall(VecDestroy(&v));
  }
  PetscCall(PetscRandomDestroy(&pr));
  PetscCall(ParMGMCFinalize());
  return 0;
}

/* ---- a generic one-rank sampler run: PCSORGIBBS (src/pc_sorgibbs.c) / PCCHOLSAMPLER (src/pc_chols.c, dense LAPACK branch) /
 *      PCMCGIBBS, optionally on a MATLRC operator A + B diag(S) B^T (k > 0): setfromoptions + setup + applyrichardson (its > 0)
 *      or apply (its == 0), noise from the library's global rander48 stream ------------------------------------------------- */
int ref_sampler_run(const char *pctype, int n, const int *rowptr, const int *col, const double *val, int k, const double *B, const double *S, int ncolors, const unsigned short *color,
                    const char *opt1, const char *val1, const char *opt2, const char *val2, long long seed, int its, const double *b, double *y, ref_sample_cb cb, void *cbctx)
{
  Mat                         A, P, U = NULL;
  PC                          pc;
  Vec                         vb = NULL, vy, vw, vS = NULL;
  PetscRandom                 pr;
  PetscInt                    outits = 0;
  PCRichardsonConvergedReason reason = PCRICHARDSON_CONVERGED_ITS;
  PetscStubWorldBegin(color && ncolors > 1 ? 2 : 1);
  PetscCall(ParMGMCInitialize());
  PetscCall(ParMGMCGetPetscRandom(&pr));
  PetscCall(PetscRandomSetSeed(pr, seed));
  PetscCall(PetscRandomSeed(pr));
  PetscCall(PetscRandomDestroy(&pr));
  PetscCall(MatStubCreateSeqAIJ(MPI_COMM_WORLD, n, n, rowptr, col, val, &A));
  if (color && ncolors > 1) PetscCall(MatStubInjectColoring(A, ncolors, color));
  P = A;
  if (k > 0) {
    PetscCall(MatCreateSeqDense(MPI_COMM_WORLD, n, k, (double *)B, &U));
    PetscCall(VecStubCreate(MPI_COMM_SELF, k, k, 0, (double *)S, &vS));
    PetscCall(MatCreateLRC(A, U, vS, NULL, &P));
  }
  PetscCall(PetscStubOptionsClear());
  if (opt1 && opt1[0]) PetscCall(PetscStubOptionsSet(opt1, val1 ? val1 : ""));
  if (opt2 && opt2[0]) PetscCall(PetscStubOptionsSet(opt2, val2 ? val2 : ""));
  PetscCall(PCStubCreate(pctype, P, &pc));
  if (pc->ops->setfromoptions) PetscCall(pc->ops->setfromoptions(pc, NULL));
  PetscCall(pc->ops->setup(pc));
  pc->setupcalled = PETSC_TRUE;
  g_cb            = cb;
  g_cbctx         = cbctx;
  if (cb) PetscCall(PCSetSampleCallback(pc, cb_tramp, NULL, NULL));
  if (b) PetscCall(VecStubCreate(MPI_COMM_WORLD, n, n, 0, (double *)b, &vb));
  else PetscCall(VecStubCreate(MPI_COMM_WORLD, n, n, 0, NULL, &vb));
  PetscCall(VecStubCreate(MPI_COMM_WORLD, n, n, 0, y, &vy));
  PetscCall(VecStubCreate(MPI_COMM_WORLD, n, n, 0, NULL, &vw));
  if (its > 0) {
    PetscCall(pc->ops->applyrichardson(pc, vb, vy, vw, 0, 0, 0, its, PETSC_FALSE, &outits, &reason));
    if (outits != its || reason != PCRICHARDSON_CONVERGED_ITS) return PetscStubError(PETSC_ERR_PLIB, __FILE__, __LINE__, "unexpected outits/reason");
  } else PetscCall(pc->ops->apply(pc, vb, vy));
  PetscCall(VecDestroy(&vb));
  PetscCall(VecDestroy(&vy));
  PetscCall(VecDestroy(&vw));
  PetscCall(PCStubDestroy(&pc));
  if (k > 0) {
    PetscCall(MatDestroy(&P));
    PetscCall(MatDestroy(&U));
    PetscCall(VecDestroy(&vS));
  }
  PetscCall(MatDestroy(&A));
  PetscCall(PetscStubOptionsClear());
  PetscCall(ParMGMCFinalize());
  PetscStubWorldEnd();
  return 0;
}

/* MCSORApply on a MATLRC operator (src/mc_sor.c:565-595 set-up, :480-544 MCSORBuildLRCCorrection, :101-112 post-correction) */
int ref_mcsor_lrc(int n, const int *rowptr, const int *col, const double *val, int k, const double *B, const double *S, int ncolors, const unsigned short *color, double omega, int type, int nsweeps, const double *b, double *y)
{
  Mat   A, U, P;
  MCSOR mc;
  Vec   vb, vy, vS;
  PetscStubWorldBegin(color && ncolors > 1 ? 2 : 1);
  PetscCall(MatStubCreateSeqAIJ(MPI_COMM_WORLD, n, n, rowptr, col, val, &A));
  if (color && ncolors > 1) PetscCall(MatStubInjectColoring(A, ncolors, color));
  PetscCall(MatCreateSeqDense(MPI_COMM_WORLD, n, k, (double *)B, &U));
  PetscCall(VecStubCreate(MPI_COMM_SELF, k, k, 0, (double *)S, &vS));
  PetscCall(MatCreateLRC(A, U, vS, NULL, &P));
  PetscCall(PetscStubOptionsClear());
  PetscCall(MCSORCreate(P, &mc));
  PetscCall(MCSORSetUp(mc));
  PetscCall(MCSORSetOmega(mc, omega));
  PetscCall(MCSORSetSweepType(mc, (MatSORType)type));
  PetscCall(VecStubCreate(MPI_COMM_WORLD, n, n, 0, (double *)b, &vb));
  PetscCall(VecStubCreate(MPI_COMM_WORLD, n, n, 0, y, &vy));
  for (int s = 0; s < nsweeps; ++s) PetscCall(MCSORApply(mc, vb, vy));
  PetscCall(VecDestroy(&vb));
  PetscCall(VecDestroy(&vy));
  PetscCall(MCSORDestroy(&mc));
  PetscCall(MatDestroy(&P));
  PetscCall(MatDestroy(&U));
  PetscCall(VecDestroy(&vS));
  PetscCall(MatDestroy(&A));
  PetscStubWorldEnd();
  return 0;
}

/* src/iact.c: Autocorrelation (FFT, zero-padded) and IACT (Sokal windowing, c = 5) */
int ref_iact(int n, const double *x, double *tau, double *acf_out, int *valid)
{
  PetscScalar *acf = NULL;
  PetscBool    v   = PETSC_FALSE;
  PetscCall(IACT(n, x, tau, acf_out ? &acf : NULL, &v));
  if (acf_out) {
    memcpy(acf_out, acf, sizeof(double) * (size_t)n);
    free(acf);
  }
  *valid = v ? 1 : 0;
  return 0;
}
int ref_autocorrelation(int n, const double *x, double *acf_out)
{
  PetscScalar *acf = NULL;
  PetscCall(Autocorrelation(n, x, &acf));
  memcpy(acf_out, acf, sizeof(double) * (size_t)n);
  free(acf);
  return 0;
}

/* src/stats.c: EstimateCovarianceMatErrors; samples[(i * chains + c) * n ...] = sample i of chain c */
int ref_cov_errors(int n, const int *rowptr, const int *col, const double *val, int chains, int samples_per_chain, const double *samples, double *errs)
{
  Mat  A;
  Vec *vs = malloc(sizeof(Vec) * (size_t)chains * (size_t)samples_per_chain);
  PetscStubWorldBegin(1);
  PetscCall(MatStubCreateSeqAIJ(MPI_COMM_SELF, n, n, rowptr, col, val, &A));
  for (int q = 0; q < chains * samples_per_chain; ++q) PetscCall(VecStubCreate(MPI_COMM_SELF, n, n, 0, (double *)samples + (size_t)q * n, &vs[q]));
  PetscCall(EstimateCovarianceMatErrors(A, chains, samples_per_chain, vs, errs));
  for (int q = 0; q < chains * samples_per_chain; ++q) PetscCall(VecDestroy(&vs[q]));
  free(vs);
  PetscCall(MatDestroy(&A));
  PetscStubWorldEnd();
  return 0;
}


/* src/problems.c MatAssembleShiftedLaplaceFD on an mx x my grid (one rank): dense column-major result, n = mx my */
#include "parmgmc/problems.h"
int ref_assemble_laplace2d(int mx, int my, double kappa, double *dense_colmajor)
{
  DM  dm;
  Mat A;
  PetscCall(DMStubCreate2d(mx, my, &dm));
  PetscCall(MatCreateSeqDense(MPI_COMM_SELF, mx * my, mx * my, NULL, &A)); /* the stub's dense Mat owns (zeroed) storage */
  PetscCall(MatAssembleShiftedLaplaceFD(dm, kappa, A));
  memcpy(dense_colmajor, A->d, sizeof(double) * (size_t)mx * my * mx * my);
  PetscCall(MatDestroy(&A));
  PetscCall(DMDestroy(&dm));
  return 0;
}
