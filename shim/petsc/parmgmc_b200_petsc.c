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

PetscClassId  PARMGMC_CLASSID;
PetscLogEvent MULTICOL_SOR, VEC_SET_RANDOM_NORMAL;

static pmg_ctx g_ctx; /* one device context per process = per MPI rank (replaces the parmgmc_rand singleton, src/parmgmc.c:38-42) */

typedef struct {
  const char *type;
  pmg_pc      pc;
  pmg_mat     mat;
  PetscErrorCode (*scb)(PetscInt, Vec, void *);
  void *cbctx;
  PetscErrorCode (*del_scb)(void *);
  Vec cbvec; /* Vec handed to the user's callback */
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
                                               "-gamgmc_mg_coarse_pc_type", "-pc_cholsampler_dense_threshold", "-pc_b200_coloring", "-pc_b200_noise", "-pc_b200_cycle", "-pc_b200_grid", NULL};

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
  }
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* PCSetUp_*: hand the local CSR block to the device (MatSeqAIJGetCSRAndMemType, src/mc_sor.c:142,250) */
static PetscErrorCode PCSetUp_B200(PC pc)
{
  PC_B200        *d = pc->data;
  Mat             P = pc->pmat;
  MatType         type;
  const PetscInt *ia, *ja;
  PetscScalar    *va;
  PetscInt        n;
  int64_t        *rowptr;

  PetscFunctionBeginUser;
  PetscCall(MatGetType(P, &type));
  PetscCheck(strcmp(type, MATSEQAIJ) == 0, PetscObjectComm((PetscObject)pc), PETSC_ERR_SUP, "parmgmc_b200 shim: SEQAIJ operators (one rank per GPU with slab-partitioned structured grids goes through pmg_mat_create_laplace)");
  PetscCall(MatSeqAIJGetCSRAndMemType(P, &ia, &ja, &va, NULL));
  PetscCall(MatGetSize(P, &n, NULL));
  PetscCall(PetscMalloc1(n + 1, &rowptr));
  for (PetscInt r = 0; r <= n; ++r) rowptr[r] = ia[r];
  if (d->mat) PMGCall(pmg_mat_destroy(d->mat));
  PMGCall(pmg_mat_create_csr(g_ctx, n, rowptr, ja, va, &d->mat)); /* PetscInt is 32-bit here; widen for --with-64-bit-indices */
  PetscCall(PetscFree(rowptr));
  PMGCall(pmg_pc_set_operator(d->pc, d->mat));
  PMGCall(pmg_pc_setup(d->pc));
  PetscFunctionReturn(PETSC_SUCCESS);
}

static int B200SampleTrampoline(int64_t it, const double *y_host, int64_t n, void *ctx)
{
  PC_B200     *d = ctx;
  PetscScalar *a;
  if (!d->scb) return 0;
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
  PMGCall(pmg_pc_apply_richardson(d->pc, barr, yarr, its, guesszero ? 1 : 0, &oits, &r));
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
  PMGCall(pmg_pc_apply(d->pc, xarr, yarr));
  PetscCall(VecRestoreArray(y, &yarr));
  PetscCall(VecRestoreArrayRead(x, &xarr));
  PetscFunctionReturn(PETSC_SUCCESS);
}

static PetscErrorCode PCReset_B200(PC pc)
{
  PC_B200 *d = pc->data;

  PetscFunctionBeginUser;
  PMGCall(pmg_pc_reset(d->pc));
  if (d->mat) PMGCall(pmg_mat_destroy(d->mat));
  d->mat = NULL;
  PetscCall(VecDestroy(&d->cbvec));
  PetscFunctionReturn(PETSC_SUCCESS);
}

static PetscErrorCode PCDestroy_B200(PC pc)
{
  PC_B200 *d = pc->data;

  PetscFunctionBeginUser;
  PetscCall(PCReset_B200(pc));
  if (d->del_scb) PetscCall(d->del_scb(d->cbctx));
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
  PetscCall(PetscObjectComposeFunction((PetscObject)pc, "PCSetSampleCallback_C", PCSetSampleCallback_B200));
  PetscFunctionReturn(PETSC_SUCCESS);
}

PetscErrorCode PCCreate_MulticolorGibbs(PC pc) { return PCCreate_B200(pc, PCMCGIBBS); }
PetscErrorCode PCCreate_SORGibbs(PC pc) { return PCCreate_B200(pc, PCSORGIBBS); }
PetscErrorCode PCCreate_GAMGMC(PC pc) { return PCCreate_B200(pc, PCGAMGMC); }
PetscErrorCode PCCreate_CholSampler(PC pc) { return PCCreate_B200(pc, PCCHOLSAMPLER); }

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
  PetscCall(PCRegister(PCSORGIBBS, PCCreate_SORGibbs));
  PetscCall(PCRegister(PCMCGIBBS, PCCreate_MulticolorGibbs));
  PetscCall(PCRegister(PCGAMGMC, PCCreate_GAMGMC));
  PetscCall(PCRegister(PCCHOLSAMPLER, PCCreate_CholSampler));
  PetscCall(PetscClassIdRegister("ParMGMC", &PARMGMC_CLASSID));
  PetscCall(PetscLogEventRegister("MulticolSOR", PARMGMC_CLASSID, &MULTICOL_SOR));
  PetscCall(PetscLogEventRegister("VecSetRandN", PARMGMC_CLASSID, &VEC_SET_RANDOM_NORMAL));
  PetscFunctionReturn(PETSC_SUCCESS);
}

PetscErrorCode ParMGMCFinalize(void)
{
  PetscFunctionBeginUser;
  if (g_ctx) PMGCall(pmg_ctx_destroy(g_ctx));
  g_ctx = NULL;
  PetscFunctionReturn(PETSC_SUCCESS);
}
