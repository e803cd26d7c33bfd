/* parmgmc_b200_petsc.c -- PETSc shim: registers the reference's PC type names and forwards every PC operation to
 * the C ABI of libparmgmc_b200.so (include/parmgmc_b200.h).
 *
 * A host program written against ParMGMC keeps its source: it calls ParMGMCInitialize() once after
 * PetscInitialize() (examples/ex1.c:77-78 of the reference) and then only PETSc API with
 * -pc_type mcgibbs | sorgibbs | gamgmc | cholsampler.  Linking this file + libparmgmc_b200.so instead of libparmgmc.so
 * swaps the sampling hot path for the CUDA one.
 *
 * Mirrors:  src/parmgmc.c:44-54, :118-151 (registration, PCSetSampleCallback dispatch)
 *           src/pc_mcgibbs.c:305-326, src/pc_sorgibbs.c:307-324, src/pc_gamgmc.c:382-415, src/pc_chols.c:420-448
 *           (the PCCreate_* vtables: setup / apply / applyrichardson / setfromoptions / reset / destroy / view)
 *
 * PETSc is not installed in the build image, so this file is compiled there only against oracle/petsc_stub (a syntax /
 * type check, tests/test_abi_cpu.py); INTEGRATION.md lists what a maintainer verifies against a real PETSc.
 */
#include <petsc/private/pcimpl.h>
#include <petscksp.h>
#include <petscmat.h>

#include "../../include/parmgmc_b200.h"

#define PCMCGIBBS "mcgibbs"
#define PCGAMGMC "gamgmc"
#define PCSORGIBBS "sorgibbs"
#define PCCHOLSAMPLER "cholsampler"
#define PCWOODBURY "woodbury"

/* include/parmgmc/mc_sor.h:17-19 */
typedef struct _MCSOR {
  void *ctx;
} *MCSOR;

PetscClassId  PARMGMC_CLASSID;
PetscLogEvent MULTICOL_SOR, VEC_SET_RANDOM_NORMAL;

static pmg_ctx     g_ctx;  /* one device context per process = per MPI rank (replaces the parmgmc_rand singleton, src/parmgmc.c:38-42) */
static PetscRandom g_rand; /* the object ParMGMCGetPetscRandom hands out; its SEED keys the device generator */
static uint64_t    g_vec_draws; /* VecSetRandomStandardNormal calls made through this shim (the device generator's block counter) */

typedef struct {
  const char *type;
  pmg_pc      pc;
  pmg_mat     mat;
  PetscErrorCode (*scb)(PetscInt, Vec, void *);
  void *cbctx;
  PetscErrorCode (*del_scb)(void *);
  Vec cbvec; /* Vec handed to the user's callback */
  PC  internal;                     /* PCGAMGMCGet/SetInternalPC (src/pc_gamgmc.c:116-133) */
  PC  wb_solver, wb_sampler;        /* PCWoodburySetSolver / SetSampler (src/woodbury.c:188-214) */
  PetscBool is_gamg_coarse;         /* PCCholSamplerSetIsCoarseGAMG (src/pc_chols.c:398-406) */
  PetscBool in_solve, richardson;   /* PCPreSolve / PCPostSolve_CholSampler (src/pc_chols.c:344-370) */
  PetscBool in_apply;               /* inside PCApply: the shim notifies the callback itself, with the in-solve sample index */
  PetscInt  sample_index;
  pmg_mat   base;                   /* base matrix of a MATLRC operator (kept alive next to d->mat) */
} PC_B200;

#define PMGCall(call)                                                                            \
  do {                                                                                           \
    int rc_ = (call);                                                                            \
    PetscCheck(rc_ == PMG_OK, PETSC_COMM_SELF, rc_ == PMG_ERR_SUP ? PETSC_ERR_SUP : PETSC_ERR_LIB, "parmgmc_b200: %s", pmg_last_error()); \
  } while (0)

/* the option keys of SURVEY Appendix C that the core understands; values are forwarded verbatim */
static const char *const b200_option_keys[] = {"-pc_mcgibbs_omega", "-pc_mcgibbs_forward", "-pc_mcgibbs_backward", "-pc_mcgibbs_symmetric", "-pc_sorgibbs_forward", "-pc_sorgibbs_local_forward",
                                               "-pc_gamgmc_mg_type", "-gamgmc_pc_mg_levels", "-gamgmc_mg_levels_ksp_type", "-gamgmc_mg_levels_ksp_max_it", "-gamgmc_mg_levels_pc_type",
                                               "-gamgmc_mg_levels_pc_mcgibbs_omega", "-gamgmc_mg_levels_pc_mcgibbs_symmetric", "-gamgmc_mg_coarse_ksp_type", "-gamgmc_mg_coarse_ksp_max_it",
                                               "-gamgmc_mg_coarse_pc_type", "-pc_cholsampler_dense_threshold", "-pc_cholsampler_coarse_gamg", "-pc_woodbury_solver", "-pc_woodbury_sampler", "-pc_b200_coloring", "-pc_b200_noise", "-pc_b200_cycle", "-pc_b200_grid", NULL};

static PetscErrorCode PCSetFromOptions_B200(PC pc, PetscOptionItems PetscOptionsObject)
{
  PC_B200 *d = pc->data;
  char     buf[128];

  PetscFunctionBeginUser;
  (void)PetscOptionsObject;
  for (int k = 0; b200_option_keys[k]; ++k) {
    PetscBool set = PETSC_FALSE;
    PetscCall(PetscOptionsGetString(NULL, ((PetscObject)pc)->prefix, b200_option_keys[k], buf, sizeof buf, &set));
    if (set) PMGCall(pmg_pc_set_option(d->pc, b200_option_keys[k], buf));
    if (strcmp(d->type, PCWOODBURY) == 0) { /* options of the two inner PCs: -pc_woodbury_{sampler,solver}_<key> (src/woodbury.c:188-214, :250-251) */
      static const char *const inner[] = {"-pc_woodbury_sampler_", "-pc_woodbury_solver_"};
      for (int w = 0; w < 2; ++w) {
        char key[160];
        snprintf(key, sizeof key, "%s%s", inner[w], b200_option_keys[k] + 1);
        set = PETSC_FALSE;
        PetscCall(PetscOptionsGetString(NULL, ((PetscObject)pc)->prefix, key, buf, sizeof buf, &set));
        if (set) PMGCall(pmg_pc_set_option(d->pc, key, buf));
      }
    }
  }
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* One pmg_mat for one PETSc Mat.
 *   MATSEQAIJ  the CSR arrays of MatSeqAIJGetCSRAndMemType (src/mc_sor.c:142,250)                     -> pmg_mat_create_csr
 *   MATMPIAIJ  diagonal + off-diagonal block + column map of MatMPIAIJGetSeqAIJ (src/mc_sor.c:308-310),
 *              re-joined into this rank's rows with GLOBAL columns                                      -> pmg_mat_create_csr_dist
 *   MATLRC     MatLRCGetMats (src/mc_sor.c:565-571, src/pc_mcgibbs.c:236-244): base AIJ, dense B, diagonal S -> pmg_mat_create_lrc
 * PetscInt is 32-bit here; widen the copies for --with-64-bit-indices. */
static PetscErrorCode B200MatFromPetsc(Mat P, pmg_mat *out, pmg_mat *base_out)
{
  MatType type;

  PetscFunctionBeginUser;
  *out      = NULL;
  *base_out = NULL;
  PetscCall(MatGetType(P, &type));
  if (strcmp(type, MATSEQAIJ) == 0) {
    const PetscInt *ia, *ja;
    PetscScalar    *va;
    PetscInt        n;
    int64_t        *rowptr;
    PetscCall(MatSeqAIJGetCSRAndMemType(P, &ia, &ja, &va, NULL));
    PetscCall(MatGetSize(P, &n, NULL));
    PetscCall(PetscMalloc1(n + 1, &rowptr));
    for (PetscInt r = 0; r <= n; ++r) rowptr[r] = ia[r];
    PMGCall(pmg_mat_create_csr(g_ctx, n, rowptr, ja, va, out));
    PetscCall(PetscFree(rowptr));
  } else if (strcmp(type, MATMPIAIJ) == 0) {
    Mat             Ad, Ao;
    const PetscInt *colmap, *di, *dj, *oi, *oj;
    PetscScalar    *da, *oa, *val;
    PetscInt        m, M, rstart, rend;
    int64_t        *rowptr, *col, nnz = 0;
    PetscCall(MatMPIAIJGetSeqAIJ(P, &Ad, &Ao, &colmap));
    PetscCall(MatSeqAIJGetCSRAndMemType(Ad, &di, &dj, &da, NULL));
    PetscCall(MatSeqAIJGetCSRAndMemType(Ao, &oi, &oj, &oa, NULL));
    PetscCall(MatGetLocalSize(P, &m, NULL));
    PetscCall(MatGetSize(P, &M, NULL));
    PetscCall(MatGetOwnershipRange(P, &rstart, &rend));
    PetscCall(PetscMalloc1(m + 1, &rowptr));
    PetscCall(PetscMalloc1(di[m] + oi[m], &col));
    PetscCall(PetscMalloc1(di[m] + oi[m], &val));
    rowptr[0] = 0;
    for (PetscInt r = 0; r < m; ++r) { /* local columns first, ghost columns after them: the accumulation order of src/mc_sor.c:323-334 */
      for (PetscInt k = di[r]; k < di[r + 1]; ++k) { col[nnz] = (int64_t)dj[k] + rstart; val[nnz++] = da[k]; }
      for (PetscInt k = oi[r]; k < oi[r + 1]; ++k) { col[nnz] = colmap[oj[k]]; val[nnz++] = oa[k]; }
      rowptr[r + 1] = nnz;
    }
    PMGCall(pmg_mat_create_csr_dist(g_ctx, M, rstart, m, rowptr, col, val, out));
    PetscCall(PetscFree(rowptr));
    PetscCall(PetscFree(col));
    PetscCall(PetscFree(val));
  } else if (strcmp(type, MATLRC) == 0) {
    Mat                A, B;
    Vec                S;
    const PetscScalar *barr, *sarr;
    PetscInt           k;
    pmg_mat            unused;
    PetscCall(MatLRCGetMats(P, &A, &B, &S, NULL));
    PetscCall(B200MatFromPetsc(A, base_out, &unused));
    PetscCall(MatGetSize(B, NULL, &k));
    PetscCall(MatDenseGetArrayRead(B, &barr)); /* column-major, leading dimension = local rows (MatDenseGetLDA on a real PETSc) */
    PetscCall(VecGetArrayRead(S, &sarr));
    PMGCall(pmg_mat_create_lrc(*base_out, (int)k, barr, sarr, out));
    PetscCall(VecRestoreArrayRead(S, &sarr));
    PetscCall(MatDenseRestoreArrayRead(B, &barr));
  } else SETERRQ(PetscObjectComm((PetscObject)P), PETSC_ERR_SUP, "parmgmc_b200 shim: operator type %s (seqaij | mpiaij | lrc)", type);
  PetscFunctionReturn(PETSC_SUCCESS);
}

static PetscErrorCode B200DropMats(PC_B200 *d)
{
  PetscFunctionBeginUser;
  if (d->mat) PMGCall(pmg_mat_destroy(d->mat));
  if (d->base) PMGCall(pmg_mat_destroy(d->base));
  d->mat = d->base = NULL;
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* PCSetUp_*: hand the operator to the device */
static PetscErrorCode PCSetUp_B200(PC pc)
{
  PC_B200 *d = pc->data;

  PetscFunctionBeginUser;
  PetscCall(B200DropMats(d));
  PetscCall(B200MatFromPetsc(pc->pmat, &d->mat, &d->base));
  if (strcmp(d->type, PCWOODBURY) == 0) { /* src/woodbury.c:149: both inner PCs are mandatory; objects given through the setters name their type */
    if (d->wb_sampler && d->wb_sampler->data) PMGCall(pmg_pc_set_option(d->pc, "-pc_woodbury_sampler", ((PC_B200 *)d->wb_sampler->data)->type));
    if (d->wb_solver && d->wb_solver->data) PMGCall(pmg_pc_set_option(d->pc, "-pc_woodbury_solver", ((PC_B200 *)d->wb_solver->data)->type));
  }
  PMGCall(pmg_pc_set_operator(d->pc, d->mat));
  PMGCall(pmg_pc_setup(d->pc));
  PetscFunctionReturn(PETSC_SUCCESS);
}

static int B200SampleTrampoline(int64_t it, const double *y_host, int64_t n, void *ctx)
{
  PC_B200     *d = ctx;
  PetscScalar *a;
  if (!d->scb || d->in_apply) return 0;
  if (VecGetArray(d->cbvec, &a)) return 1;
  memcpy(a, y_host, sizeof(double) * (size_t)n);
  if (VecRestoreArray(d->cbvec, &a)) return 1;
  return (int)d->scb((PetscInt)it, d->cbvec, d->cbctx);
}

static PetscErrorCode PCApplyRichardson_B200(PC pc, Vec b, Vec y, Vec w, PetscReal rtol, PetscReal abstol, PetscReal dtol, PetscInt its, PetscBool guesszero, PetscInt *outits, PCRichardsonConvergedReason *reason)
{
  PC_B200           *d = pc->data;
  const PetscScalar *barr = NULL;
  PetscScalar       *yarr;
  int64_t            oits;
  int                r;

  PetscFunctionBeginUser;
  (void)w; (void)rtol; (void)abstol; (void)dtol; /* every reference sampler ignores them (src/pc_mcgibbs.c:157-160) */
  if (d->scb && !d->cbvec) PetscCall(VecDuplicate(y, &d->cbvec));
  if (b) PetscCall(VecGetArrayRead(b, &barr));
  PetscCall(VecGetArray(y, &yarr));
  d->richardson = PETSC_TRUE;
  PMGCall(pmg_pc_apply_richardson(d->pc, barr, yarr, its, guesszero ? 1 : 0, &oits, &r));
  d->richardson = PETSC_FALSE;
  PetscCall(VecRestoreArray(y, &yarr));
  if (b) PetscCall(VecRestoreArrayRead(b, &barr));
  *outits = (PetscInt)oits;
  *reason = (PCRichardsonConvergedReason)r;
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* PCApply_SORGibbs (src/pc_sorgibbs.c:105-113) / PCApply_CholSampler (src/pc_chols.c:262-291) */
static PetscErrorCode PCApply_B200(PC pc, Vec x, Vec y)
{
  PC_B200           *d = pc->data;
  const PetscScalar *xarr;
  PetscScalar       *yarr;

  PetscFunctionBeginUser;
  PetscCall(VecGetArrayRead(x, &xarr));
  PetscCall(VecGetArray(y, &yarr));
  /* src/pc_chols.c:267: outside Richardson a callback is only legal between PCPreSolve and PCPostSolve */
  PetscCheck(strcmp(d->type, PCCHOLSAMPLER) != 0 || d->richardson || !d->scb || d->in_solve, PetscObjectComm((PetscObject)pc), PETSC_ERR_SUP, "Setting a sample callback is only supported for Cholesky sampler during KSPSolve");
  d->in_apply = PETSC_TRUE;
  PMGCall(pmg_pc_apply(d->pc, xarr, yarr));
  d->in_apply = PETSC_FALSE;
  PetscCall(VecRestoreArray(y, &yarr));
  PetscCall(VecRestoreArrayRead(x, &xarr));
  if (d->in_solve && d->scb) PetscCall(d->scb(d->sample_index++, y, d->cbctx)); /* PCCholSamplerNotifySample inside a KSPSolve */
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* PCPreSolve_CholSampler / PCPostSolve_CholSampler (src/pc_chols.c:344-370) */
static PetscErrorCode PCPreSolve_B200(PC pc, KSP ksp, Vec b, Vec x)
{
  PC_B200 *d = pc->data;
  (void)ksp; (void)b; (void)x;
  d->in_solve     = PETSC_TRUE;
  d->sample_index = 0;
  return PETSC_SUCCESS;
}
static PetscErrorCode PCPostSolve_B200(PC pc, KSP ksp, Vec b, Vec x)
{
  PC_B200 *d = pc->data;
  (void)ksp; (void)b; (void)x;
  d->in_solve = PETSC_FALSE;
  return PETSC_SUCCESS;
}

static PetscErrorCode PCReset_B200(PC pc)
{
  PC_B200 *d = pc->data;

  PetscFunctionBeginUser;
  PMGCall(pmg_pc_reset(d->pc));
  PetscCall(B200DropMats(d));
  PetscCall(VecDestroy(&d->cbvec));
  PetscFunctionReturn(PETSC_SUCCESS);
}

static PetscErrorCode PCDestroy_B200(PC pc)
{
  PC_B200 *d = pc->data;

  PetscFunctionBeginUser;
  PetscCall(PCReset_B200(pc));
  if (d->del_scb) PetscCall(d->del_scb(d->cbctx));
  PetscCall(PCDestroy(&d->internal));
  PetscCall(PCDestroy(&d->wb_solver));
  PetscCall(PCDestroy(&d->wb_sampler));
  PMGCall(pmg_pc_destroy(d->pc));
  PetscCall(PetscFree(d));
  pc->data = NULL;
  PetscFunctionReturn(PETSC_SUCCESS);
}

static PetscErrorCode PCView_B200(PC pc, PetscViewer viewer)
{
  PC_B200 *d = pc->data;
  char     buf[4096];

  PetscFunctionBeginUser;
  PMGCall(pmg_pc_view(d->pc, buf, sizeof buf));
  PetscCall(PetscViewerASCIIPrintf(viewer, "%s", buf));
  PetscFunctionReturn(PETSC_SUCCESS);
}

static PetscErrorCode PCSetSampleCallback_B200(PC pc, PetscErrorCode (*cb)(PetscInt, Vec, void *), void *ctx, PetscErrorCode (*deleter)(void *))
{
  PC_B200 *d = pc->data;

  PetscFunctionBeginUser;
  if (d->del_scb) { /* src/pc_mcgibbs.c:295-298 */
    PetscCall(d->del_scb(d->cbctx));
    d->del_scb = NULL;
  }
  d->scb     = cb;
  d->cbctx   = ctx;
  d->del_scb = deleter;
  PMGCall(pmg_pc_set_sample_callback(d->pc, cb ? B200SampleTrampoline : NULL, d, NULL));
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* include/parmgmc/parmgmc.h:40: how a PC type publishes its callback setter */
PetscErrorCode PCRegisterSetSampleCallback(PC pc, PetscErrorCode (*set)(PC, PetscErrorCode (*)(PetscInt, Vec, void *), void *, PetscErrorCode (*)(void *)))
{
  PetscFunctionBeginUser;
  PetscCall(PetscObjectComposeFunction((PetscObject)pc, "PCSetSampleCallback_C", set));
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* "PCMGGetLevels_C" of a gamgmc PC (src/pc_gamgmc.c:358-366, :413): what PCMGGetLevels(pc, &l) resolves to */
static PetscErrorCode PCGAMGMCGetLevels_B200(PC pc, PetscInt *levels)
{
  int l = 0;

  PetscFunctionBeginUser;
  PMGCall(pmg_pc_gamgmc_get_levels(((PC_B200 *)pc->data)->pc, &l));
  *levels = (PetscInt)l;
  PetscFunctionReturn(PETSC_SUCCESS);
}

static PetscErrorCode PCCreate_B200(PC pc, const char *type)
{
  PC_B200 *d;

  PetscFunctionBeginUser;
  PetscCall(PetscNew(&d));
  d->type = type;
  PMGCall(pmg_pc_create(g_ctx, type, &d->pc));
  pc->data                 = d;
  pc->ops->setup           = PCSetUp_B200;
  pc->ops->destroy         = PCDestroy_B200;
  pc->ops->applyrichardson = PCApplyRichardson_B200;
  pc->ops->setfromoptions  = PCSetFromOptions_B200;
  pc->ops->reset           = PCReset_B200;
  pc->ops->view            = PCView_B200;
  if (strcmp(type, PCSORGIBBS) == 0 || strcmp(type, PCCHOLSAMPLER) == 0) pc->ops->apply = PCApply_B200;
  if (strcmp(type, PCCHOLSAMPLER) == 0) { /* src/pc_chols.c:443-444 */
    pc->ops->presolve  = PCPreSolve_B200;
    pc->ops->postsolve = PCPostSolve_B200;
  }
  PetscCall(PCRegisterSetSampleCallback(pc, PCSetSampleCallback_B200));
  if (strcmp(type, PCGAMGMC) == 0) PetscCall(PetscObjectComposeFunction((PetscObject)pc, "PCMGGetLevels_C", PCGAMGMCGetLevels_B200)); /* src/pc_gamgmc.c:413 */
  PetscFunctionReturn(PETSC_SUCCESS);
}

PetscErrorCode PCCreate_MulticolorGibbs(PC pc) { return PCCreate_B200(pc, PCMCGIBBS); }
PetscErrorCode PCCreate_SORGibbs(PC pc) { return PCCreate_B200(pc, PCSORGIBBS); }
PetscErrorCode PCCreate_GAMGMC(PC pc) { return PCCreate_B200(pc, PCGAMGMC); }
PetscErrorCode PCCreate_CholSampler(PC pc) { return PCCreate_B200(pc, PCCHOLSAMPLER); }
PetscErrorCode PCCreate_Woodbury(PC pc) { return PCCreate_B200(pc, PCWOODBURY); } /* src/woodbury.c:288-302 */

/* include/parmgmc/pc/pc_mcgibbs.h:17-18, pc_gamgmc.h:16 */
PetscErrorCode PCMulticolorGibbsSetOmega(PC pc, PetscReal omega)
{
  PetscFunctionBeginUser;
  PMGCall(pmg_pc_mcgibbs_set_omega(((PC_B200 *)pc->data)->pc, omega));
  PetscFunctionReturn(PETSC_SUCCESS);
}
PetscErrorCode PCMulticolorGibbsSetSweepType(PC pc, MatSORType type)
{
  PetscFunctionBeginUser;
  PMGCall(pmg_pc_mcgibbs_set_sweep_type(((PC_B200 *)pc->data)->pc, (int)type));
  PetscFunctionReturn(PETSC_SUCCESS);
}
PetscErrorCode PCGAMGMCSetLevels(PC pc, PetscInt levels)
{
  PetscFunctionBeginUser;
  PMGCall(pmg_pc_gamgmc_set_levels(((PC_B200 *)pc->data)->pc, (int)levels));
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* include/parmgmc/pc/pc_gamgmc.h:17-18.  The reference hands out its inner PCMG so that a caller can configure levels,
 * smoothers and interpolations on it before PCSetUp.  Here the hierarchy lives on the device behind the gamgmc PC itself, so
 * the "internal PC" is a configuration carrier: a PC set through PCGAMGMCSetInternalPC is kept (reference-counted, returned
 * by Get) and, when it is one of this library's PCs, its level count is adopted; without one, Get returns the gamgmc PC, on
 * which PCMGGetLevels works through the composed "PCMGGetLevels_C". */
PetscErrorCode PCGAMGMCGetInternalPC(PC pc, PC *mg)
{
  PC_B200 *d = pc->data;

  PetscFunctionBeginUser;
  if (mg) *mg = d->internal ? d->internal : pc;
  PetscFunctionReturn(PETSC_SUCCESS);
}
PetscErrorCode PCGAMGMCSetInternalPC(PC pc, PC mg)
{
  PC_B200 *d = pc->data;

  PetscFunctionBeginUser;
  PetscCall(PCDestroy(&d->internal));
  d->internal = mg;
  PetscCall(PetscObjectReference((PetscObject)mg));
  if (mg->ops->applyrichardson == PCApplyRichardson_B200 && strcmp(((PC_B200 *)mg->data)->type, PCGAMGMC) == 0) {
    int l = 0;
    PMGCall(pmg_pc_gamgmc_get_levels(((PC_B200 *)mg->data)->pc, &l));
    if (l > 0) PMGCall(pmg_pc_gamgmc_set_levels(d->pc, l));
  }
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* include/parmgmc/pc/pc_chols.h:16 (src/pc_chols.c:398-406): the coarse sampler of a GAMG hierarchy lives on rank 0 only; the
 * device hierarchy replicates / agglomerates its coarse levels itself, so the flag is recorded and forwarded as an option */
PetscErrorCode PCCholSamplerSetIsCoarseGAMG(PC pc, PetscBool flag)
{
  PC_B200 *d = pc->data;

  PetscFunctionBeginUser;
  d->is_gamg_coarse = flag;
  PMGCall(pmg_pc_set_option(d->pc, "-pc_cholsampler_coarse_gamg", flag ? "1" : "0"));
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* include/parmgmc/pc/woodbury.h:16-17 (src/woodbury.c:188-214): the inner PCs are PCs of this shim; their TYPE is what the
 * device-side woodbury PC needs (it builds its own inner samplers on the base matrix), their options arrive through the
 * prefixed keys in PCSetFromOptions_B200 */
static PetscErrorCode PCWoodburyKeep(PC pc, PC inner, PC *slot, const char *key)
{
  PC_B200 *d = pc->data;

  PetscFunctionBeginUser;
  PetscCheck(inner->ops->applyrichardson == PCApplyRichardson_B200, PetscObjectComm((PetscObject)pc), PETSC_ERR_SUP, "parmgmc_b200 shim: the inner PC of a woodbury PC must be one of mcgibbs | sorgibbs | gamgmc | cholsampler");
  PetscCall(PetscObjectReference((PetscObject)inner));
  PetscCall(PCDestroy(slot));
  *slot = inner;
  PMGCall(pmg_pc_set_option(d->pc, key, ((PC_B200 *)inner->data)->type));
  PetscFunctionReturn(PETSC_SUCCESS);
}
PetscErrorCode PCWoodburySetSolver(PC pc, PC solver) { return PCWoodburyKeep(pc, solver, &((PC_B200 *)pc->data)->wb_solver, "-pc_woodbury_solver"); }
PetscErrorCode PCWoodburySetSampler(PC pc, PC sampler) { return PCWoodburyKeep(pc, sampler, &((PC_B200 *)pc->data)->wb_sampler, "-pc_woodbury_sampler"); }

PetscErrorCode PCSetSampleCallback(PC pc, PetscErrorCode (*cb)(PetscInt, Vec, void *), void *ctx, PetscErrorCode (*deleter)(void *))
{
  PetscFunctionBeginUser;
  PetscUseMethod((PetscObject)pc, "PCSetSampleCallback_C", (PC, PetscErrorCode(*)(PetscInt, Vec, void *), void *, PetscErrorCode (*)(void *)), (pc, cb, ctx, deleter));
  PetscFunctionReturn(PETSC_SUCCESS);
}

PetscErrorCode ParMGMCInitialize(void)
{
  PetscMPIInt rank;
  int         ndev = 0;

  PetscFunctionBeginUser;
  PetscCallMPI(MPI_Comm_rank(PETSC_COMM_WORLD, &rank));
  PMGCall(pmg_device_count(&ndev));
  PetscCheck(ndev > 0, PETSC_COMM_SELF, PETSC_ERR_LIB, "parmgmc_b200: no CUDA device (there is no CPU fallback)");
  PMGCall(pmg_ctx_create(rank % ndev, &g_ctx)); /* one rank per GPU */
  {
    PetscMPIInt   size;
    unsigned char id[128];
    PetscCallMPI(MPI_Comm_size(PETSC_COMM_WORLD, &size));
    if (size > 1) { /* NCCL communicator next to PETSC_COMM_WORLD: MPIAIJ operators exchange their ghosts over it (src/mc_sor.c:318-319) */
      if (rank == 0) PMGCall(pmg_comm_unique_id(id));
      PetscCallMPI(MPI_Bcast(id, 128, MPI_BYTE, 0, PETSC_COMM_WORLD));
      PMGCall(pmg_ctx_comm_init(g_ctx, rank, size, id));
    }
  }
  PetscCall(PCRegister(PCSORGIBBS, PCCreate_SORGibbs));
  PetscCall(PCRegister(PCMCGIBBS, PCCreate_MulticolorGibbs));
  PetscCall(PCRegister(PCGAMGMC, PCCreate_GAMGMC));
  PetscCall(PCRegister(PCCHOLSAMPLER, PCCreate_CholSampler));
  PetscCall(PCRegister(PCWOODBURY, PCCreate_Woodbury));
  PetscCall(PetscClassIdRegister("ParMGMC", &PARMGMC_CLASSID));
  PetscCall(PetscLogEventRegister("MulticolSOR", PARMGMC_CLASSID, &MULTICOL_SOR));
  PetscCall(PetscLogEventRegister("VecSetRandN", PARMGMC_CLASSID, &VEC_SET_RANDOM_NORMAL));
  PetscFunctionReturn(PETSC_SUCCESS);
}

PetscErrorCode ParMGMCFinalize(void)
{
  PetscFunctionBeginUser;
  if (g_rand) PetscCall(PetscRandomDestroy(&g_rand));
  if (g_ctx) PMGCall(pmg_ctx_destroy(g_ctx));
  g_ctx = NULL;
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* ---- include/parmgmc/parmgmc.h:43-44 -------------------------------------------------------------------------------------
 * ParMGMCGetPetscRandom (src/parmgmc.c:56-68): one reference-counted PetscRandom per process.  The samplers draw on the device
 * (Philox keyed by seed / draw block / global row); the PetscRandom's SEED is what a host program controls
 * (examples/benchmark/main.cc:228-236) and it is forwarded to the device generator here and in VecSetRandomStandardNormal. */
PetscErrorCode ParMGMCGetPetscRandom(PetscRandom *pr)
{
  PetscFunctionBeginUser;
  if (!g_rand) {
    PetscCall(PetscRandomCreate(MPI_COMM_WORLD, &g_rand));
    PetscCall(PetscRandomSetFromOptions(g_rand));
  }
  PetscCall(PetscObjectReference((PetscObject)g_rand));
  *pr = g_rand;
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* VecSetRandomStandardNormal (src/parmgmc.c:70-116) on the device: the local part of v is block `g_vec_draws` of the
 * generator keyed by the PetscRandom's seed and the global row (pmg_normal_fill), so the result does not depend on the
 * partition -- the reference's own MKL branch keys on seed ^ rank in the same spirit (:84-90). */
PetscErrorCode VecSetRandomStandardNormal(Vec v, PetscRandom r)
{
  PetscInt     n, lo = 0;
  PetscInt64   seed = 0;
  PetscScalar *a;

  PetscFunctionBeginUser;
  PetscCheck(g_ctx, PETSC_COMM_SELF, PETSC_ERR_ORDER, "ParMGMCInitialize has not been called");
  if (r) PetscCall(PetscRandomGetSeed(r, &seed));
  PetscCall(VecGetLocalSize(v, &n));
  PetscCall(VecGetOwnershipRange(v, &lo, NULL));
  PetscCall(VecGetArray(v, &a));
  PMGCall(pmg_normal_fill(g_ctx, (uint64_t)seed, g_vec_draws++, lo, n, a));
  PetscCall(VecRestoreArray(v, &a));
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* ---- include/parmgmc/mc_sor.h:17-30: the PETSc-typed MCSOR object (used directly by examples/ex3.c, ex5.c) -------------------- */
typedef struct {
  Mat       A;
  pmg_mat   mat, base;
  pmg_mcsor mc;
  PetscReal omega;
  PetscBool omega_set;
  int       sweep;
} MCSOR_B200;

PetscErrorCode MCSORCreate(Mat A, MCSOR *m)
{
  MCSOR       mc;
  MCSOR_B200 *c;

  PetscFunctionBeginUser;
  PetscCheck(g_ctx, PETSC_COMM_SELF, PETSC_ERR_ORDER, "ParMGMCInitialize has not been called");
  PetscCall(PetscNew(&mc));
  PetscCall(PetscNew(&c));
  c->A     = A; /* borrowed, like src/mc_sor.c:630 */
  c->sweep = PMG_SOR_FORWARD_SWEEP;
  PetscCall(PetscOptionsGetReal(NULL, NULL, "-mc_sor_omega", &c->omega, &c->omega_set)); /* src/mc_sor.c:638 */
  mc->ctx = c;
  *m      = mc;
  PetscFunctionReturn(PETSC_SUCCESS);
}

PetscErrorCode MCSORSetUp(MCSOR mc) /* src/mc_sor.c:553-605: diagonal pointers, idiag, colouring, ghost plan -- all inside pmg_mcsor_create */
{
  MCSOR_B200 *c = mc->ctx;

  PetscFunctionBeginUser;
  if (c->mc) PMGCall(pmg_mcsor_destroy(c->mc));
  if (c->mat) PMGCall(pmg_mat_destroy(c->mat));
  if (c->base) PMGCall(pmg_mat_destroy(c->base));
  c->mc = NULL;
  c->mat = c->base = NULL;
  PetscCall(B200MatFromPetsc(c->A, &c->mat, &c->base));
  PMGCall(pmg_mcsor_create(c->mat, &c->mc));
  if (c->omega_set) PMGCall(pmg_mcsor_set_omega(c->mc, c->omega));
  PMGCall(pmg_mcsor_set_sweep_type(c->mc, c->sweep));
  PetscFunctionReturn(PETSC_SUCCESS);
}

PetscErrorCode MCSORDestroy(MCSOR *m)
{
  PetscFunctionBeginUser;
  if (m && *m) {
    MCSOR_B200 *c = (*m)->ctx;
    if (c->mc) PMGCall(pmg_mcsor_destroy(c->mc));
    if (c->mat) PMGCall(pmg_mat_destroy(c->mat));
    if (c->base) PMGCall(pmg_mat_destroy(c->base));
    PetscCall(PetscFree(c));
    PetscCall(PetscFree(*m));
  }
  PetscFunctionReturn(PETSC_SUCCESS);
}

PetscErrorCode MCSORApply(MCSOR mc, Vec b, Vec y) /* src/mc_sor.c:216-239 */
{
  MCSOR_B200        *c = mc->ctx;
  const PetscScalar *barr;
  PetscScalar       *yarr;

  PetscFunctionBeginUser;
  PetscCheck(c->mc, PETSC_COMM_SELF, PETSC_ERR_ORDER, "MCSORSetUp has not been called");
  PetscCall(PetscLogEventBegin(MULTICOL_SOR, b, y, 0, 0));
  PetscCall(VecGetArrayRead(b, &barr));
  PetscCall(VecGetArray(y, &yarr));
  PMGCall(pmg_mcsor_apply(c->mc, barr, yarr));
  PetscCall(VecRestoreArray(y, &yarr));
  PetscCall(VecRestoreArrayRead(b, &barr));
  PetscCall(PetscLogEventEnd(MULTICOL_SOR, b, y, 0, 0));
  PetscFunctionReturn(PETSC_SUCCESS);
}

PetscErrorCode MCSORSetOmega(MCSOR mc, PetscReal omega) /* src/mc_sor.c:412-420 */
{
  MCSOR_B200 *c = mc->ctx;

  PetscFunctionBeginUser;
  c->omega     = omega;
  c->omega_set = PETSC_TRUE;
  if (c->mc) PMGCall(pmg_mcsor_set_omega(c->mc, omega));
  PetscFunctionReturn(PETSC_SUCCESS);
}

PetscErrorCode MCSORSetSweepType(MCSOR mc, MatSORType type) /* src/mc_sor.c:422-430 */
{
  MCSOR_B200 *c = mc->ctx;

  PetscFunctionBeginUser;
  PetscCheck(type == SOR_FORWARD_SWEEP || type == SOR_BACKWARD_SWEEP || type == SOR_SYMMETRIC_SWEEP, PETSC_COMM_SELF, PETSC_ERR_SUP, "Only SOR_FORWARD_SWEEP, SOR_BACKWARD_SWEEP and SOR_SYMMETRIC_SWEEP are supported");
  c->sweep = (int)type;
  if (c->mc) PMGCall(pmg_mcsor_set_sweep_type(c->mc, (int)type));
  PetscFunctionReturn(PETSC_SUCCESS);
}

PetscErrorCode MCSORGetSweepType(MCSOR mc, MatSORType *type) /* src/mc_sor.c:432-439 */
{
  PetscFunctionBeginUser;
  *type = (MatSORType)((MCSOR_B200 *)mc->ctx)->sweep;
  PetscFunctionReturn(PETSC_SUCCESS);
}

PetscErrorCode MCSORGetNumColors(MCSOR mc, PetscInt *ncolors) /* src/mc_sor.c:607-616 */
{
  MCSOR_B200 *c = mc->ctx;
  int         nc = 0;

  PetscFunctionBeginUser;
  PetscCheck(c->mc, PETSC_COMM_SELF, PETSC_ERR_ORDER, "MCSORSetUp has not been called");
  PMGCall(pmg_mcsor_get_num_colors(c->mc, &nc));
  *ncolors = nc;
  PetscFunctionReturn(PETSC_SUCCESS);
}

PetscErrorCode MCSORGetISColoring(MCSOR mc, ISColoring *isc) /* src/mc_sor.c:92-99; the caller destroys the result */
{
  MCSOR_B200      *c = mc->ctx;
  int              nc = 0;
  int64_t          n = 0;
  int32_t         *col;
  ISColoringValue *cv;

  PetscFunctionBeginUser;
  PetscCheck(c->mc, PETSC_COMM_SELF, PETSC_ERR_ORDER, "MCSORSetUp has not been called");
  PMGCall(pmg_mat_get_size(c->base ? c->base : c->mat, &n, NULL, NULL));
  PetscCall(PetscMalloc1(n, &col));
  PetscCall(PetscMalloc1(n, &cv));
  PMGCall(pmg_mat_get_coloring(c->base ? c->base : c->mat, &nc, col));
  for (int64_t r = 0; r < n; ++r) cv[r] = (ISColoringValue)col[r];
  PetscCall(ISColoringCreate(PetscObjectComm((PetscObject)c->A), nc, (PetscInt)n, cv, PETSC_OWN_POINTER, isc));
  PetscCall(ISColoringSetType(*isc, IS_COLORING_LOCAL));
  PetscCall(PetscFree(col));
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* MCSORBuildLRCCorrection (src/mc_sor.c:456-544): Bb = C (S^-1 + B^T C)^-1 with C = M^-1 B, one det_sor call per column of B.
 * det_sor is the caller's deterministic sweep (for the shim's own objects: MCSORApply -> device); the k x k algebra is PETSc's. */
PetscErrorCode MCSORBuildLRCCorrection(PetscErrorCode (*det_sor)(void *, Vec, Vec), void *ctx, Mat Asor, Mat B, Vec S, Mat *Bb)
{
  Mat                C, K, Kinv, Id;
  KSP                ksp;
  Vec                x, Sinv;
  PetscInt           k;
  const PetscScalar *s;
  PetscScalar       *si;
  MPI_Comm           comm;

  PetscFunctionBeginUser;
  PetscCall(PetscObjectGetComm((PetscObject)Asor, &comm));
  PetscCall(MatGetSize(B, NULL, &k));
  PetscCall(MatDuplicate(B, MAT_DO_NOT_COPY_VALUES, &C));
  PetscCall(MatCreateVecs(Asor, &x, NULL));
  for (PetscInt j = 0; j < k; ++j) {
    Vec bj, cj;
    PetscCall(VecZeroEntries(x));
    PetscCall(MatDenseGetColumnVecRead(B, j, &bj));
    PetscCall(det_sor(ctx, bj, x));
    PetscCall(MatDenseRestoreColumnVecRead(B, j, &bj));
    PetscCall(MatDenseGetColumnVecWrite(C, j, &cj));
    PetscCall(VecCopy(x, cj));
    PetscCall(MatDenseRestoreColumnVecWrite(C, j, &cj));
  }
  PetscCall(VecDestroy(&x));
  PetscCall(MatTransposeMatMult(B, C, MAT_INITIAL_MATRIX, 1, &K)); /* K = B^T M^-1 B, replicated k x k */
  PetscCall(MatCreateVecs(K, &Sinv, NULL));
  PetscCall(VecGetArrayRead(S, &s)); /* S is the replicated length-k diagonal the reference scatters into K's layout (:507-516) */
  PetscCall(VecGetArray(Sinv, &si));
  for (PetscInt j = 0; j < k; ++j) si[j] = 1.0 / s[j];
  PetscCall(VecRestoreArray(Sinv, &si));
  PetscCall(VecRestoreArrayRead(S, &s));
  PetscCall(MatDiagonalSet(K, Sinv, ADD_VALUES));
  PetscCall(KSPCreate(comm, &ksp));
  PetscCall(KSPSetOperators(ksp, K, K));
  PetscCall(MatDuplicate(K, MAT_DO_NOT_COPY_VALUES, &Id));
  PetscCall(MatShift(Id, 1));
  PetscCall(MatDuplicate(K, MAT_DO_NOT_COPY_VALUES, &Kinv));
  PetscCall(KSPMatSolve(ksp, Id, Kinv));
  PetscCall(MatMatMult(C, Kinv, MAT_INITIAL_MATRIX, 1, Bb));
  PetscCall(KSPDestroy(&ksp));
  PetscCall(VecDestroy(&Sinv));
  PetscCall(MatDestroy(&Id));
  PetscCall(MatDestroy(&Kinv));
  PetscCall(MatDestroy(&K));
  PetscCall(MatDestroy(&C));
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* ---- include/parmgmc/iact.h:14-15 (src/iact.c:17-92): cuFFT in place of FFTW; the caller frees *acf with PetscFree ---------- */
PetscErrorCode Autocorrelation(PetscInt n, const PetscScalar *x, PetscScalar **acf)
{
  PetscFunctionBeginUser;
  PetscCheck(g_ctx, PETSC_COMM_SELF, PETSC_ERR_ORDER, "ParMGMCInitialize has not been called");
  PetscCall(PetscMalloc1(n, acf));
  PMGCall(pmg_autocorrelation(g_ctx, n, x, *acf));
  PetscFunctionReturn(PETSC_SUCCESS);
}
PetscErrorCode IACT(PetscInt n, const PetscScalar *x, PetscScalar *tau, PetscScalar **acf, PetscBool *valid)
{
  PetscScalar *a = NULL;
  int          ok = 0;

  PetscFunctionBeginUser;
  PetscCheck(g_ctx, PETSC_COMM_SELF, PETSC_ERR_ORDER, "ParMGMCInitialize has not been called");
  if (acf) PetscCall(PetscMalloc1(n, &a));
  PMGCall(pmg_iact(g_ctx, n, x, tau, a, &ok));
  if (acf) *acf = a;
  if (valid) *valid = ok ? PETSC_TRUE : PETSC_FALSE;
  PetscFunctionReturn(PETSC_SUCCESS);
}
