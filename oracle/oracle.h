/* oracle.h -- CPU restatement of ParMGMC's sampling hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This is the checker, not the product.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The shipped path
 * (parmgmc_b200/) never links, imports or calls anything in this directory.
 *
 * Every function cites the reference file:line (relative to /root/reference) whose
 * arithmetic it restates.  PETSc arithmetic that is not in /root/reference (MatSOR,
 * PCMG cycle, DMDA Q1 interpolation, MatPtAP, potrf/trsv) follows SURVEY.md
 * Appendix A (PETSc v3.25.1 semantics).
 *
 * PARITY STATUS: the reference cannot be linked here (needs PETSc + MPI), and it
 * ships no golden vectors.  The restatement is pinned by
 *   (1) oracle/_ref: the reference's OWN mc_sor.c / pc_mcgibbs.c / parmgmc.c compiled
 *       unmodified against a container-only PETSc API stub (oracle/petsc_stub), see
 *       oracle/Makefile target `ref`, and compared bit-for-bit in tests/test_oracle_ref.py;
 *   (2) the reference's own test identities (ex5: sym == fwd;bwd, ex1/ex4: mean
 *       convergence) and the exact stationarity identity of SURVEY.md section 8(c).
 * The PETSc-internal pieces (PCMG cycle, MatPtAP, JP colouring) remain "parity
 * unpinned" against PETSc itself and are judged statistically.
 *
 * Floating-point contract (shared with the CUDA path so that injected-noise sweeps
 * are bit-identical, not merely 1e-12 close):
 *   - row accumulation  sum = fma(-a_k, y[c_k], sum), k ascending, diagonal skipped
 *   - row update        y_r = fma(idiag_r, sum, (1-omega)*y_r)
 *   - rhs preparation   w_r = (z_r*sqrtdiag_r) + b_r         (two roundings, no fma:
 *                        the reference does it as VecPointwiseMult then VecAXPY)
 * Compile with -ffp-contract=off so that gcc fuses nothing on its own.
 */
#ifndef PARMGMC_ORACLE_H
#define PARMGMC_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* PETSc MatSORType values used by the reference (petscmat.h) */
#define ORC_SOR_FORWARD 1
#define ORC_SOR_BACKWARD 2
#define ORC_SOR_SYMMETRIC 3

/* ---- problems.c ---------------------------------------------------------------- */
int64_t orc_laplace_nnz(int dim, int64_t nx, int64_t ny, int64_t nz);
void    orc_laplace_csr(int dim, int64_t nx, int64_t ny, int64_t nz, double kappa, int64_t *rowptr, int32_t *col, double *val);

/* ---- mc_sor.c -------------------------------------------------------------------- */
int  orc_diag_ptrs(int64_t n, const int64_t *rowptr, const int32_t *col, int64_t *diagptr);
void orc_idiag(int64_t n, const double *val, const int64_t *diagptr, double omega, double *idiag);
void orc_sweep_seq(int64_t n, const int64_t *rowptr, const int32_t *col, const double *val, const int64_t *diagptr, const double *idiag, double omega, int ncolors, const int64_t *colorptr, const int32_t *colorrows, int dir, const double *b, double *y);
void orc_mcsor_apply(int64_t n, const int64_t *rowptr, const int32_t *col, const double *val, const int64_t *diagptr, const double *idiag, double omega, int ncolors, const int64_t *colorptr, const int32_t *colorrows, int type, const double *b, double *y);

/* colourings (the reference's 1-rank colouring is "all rows colour 0", mc_sor.c:397-410) */
int  orc_coloring_lists(int64_t n, const int32_t *color, int ncolors, int64_t *colorptr, int32_t *colorrows);
int  orc_coloring_greedy(int64_t n, const int64_t *rowptr, const int32_t *col, int32_t *color);
int  orc_coloring_levelset(int64_t n, const int64_t *rowptr, const int32_t *col, int32_t *color);
int  orc_coloring_valid(int64_t n, const int64_t *rowptr, const int32_t *col, const double *val, const int32_t *color);

/* partitioned ("MPIAIJ") sweep, mc_sor.c:152-214 + :298-381, ranks emulated by threads */
typedef struct orc_part_s orc_part;
orc_part *orc_part_create(int64_t n, const int64_t *rowptr, const int32_t *col, const double *val, int nranks, const int64_t *rowstart, int ncolors, const int32_t *color, double omega);
void      orc_part_destroy(orc_part *p);
void      orc_part_sweep(orc_part *p, int dir, const double *b, double *y, int nthreads);
int64_t   orc_part_ghost_count(const orc_part *p, int rank, int color);
void      orc_part_ghost_index(const orc_part *p, int rank, int color, int64_t *out);

/* ---- parmgmc.c RNG + noise sources ------------------------------------------------ */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void orc_normal_philox(uint64_t seed, uint64_t call, int64_t row0, int64_t n, double *out);
void orc_normal_philox_grid(uint64_t seed, uint64_t call, int64_t row0, int64_t n, int64_t nx, int64_t pad, double *out);

typedef struct {
  int           mode; /* 0 = tape (injected), 1 = philox, 2 = rander48 Box-Muller (parmgmc.c:100-110) */
  const double *tape;
  int64_t       tape_len, tape_pos;
  uint64_t      seed, call;
  uint64_t      x48;
  /* philox mode only: blocks of exactly grid_n rows belong to a matrix-free grid operator of row length grid_nx and are
   * keyed on the PADDED index id = g + (g / grid_nx) * grid_pad (parmgmc_b200/csrc/philox.cuh); 0 = plain global rows */
  int64_t grid_n, grid_nx, grid_pad;
} orc_noise;
void orc_noise_init_tape(orc_noise *ns, const double *tape, int64_t len);
void orc_noise_init_philox(orc_noise *ns, uint64_t seed);
void orc_noise_init_rander48(orc_noise *ns, uint64_t seed);
int  orc_noise_fill(orc_noise *ns, int64_t row0, int64_t n, double *out);

/* ---- pc_mcgibbs.c / pc_sorgibbs.c ---------------------------------------------------- */
void orc_sqrtdiag(int64_t n, const double *val, const int64_t *diagptr, double omega, double *sqrtdiag);
void orc_prepare_rhs(int64_t n, const double *b, const double *sqrtdiag, const double *z, double *w);
typedef int (*orc_sample_cb)(int64_t it, const double *y, void *ctx);
int orc_gibbs_richardson(int64_t n, const int64_t *rowptr, const int32_t *col, const double *val, double omega, int ncolors, const int64_t *colorptr, const int32_t *colorrows, int type, orc_noise *ns, const double *b, double *y, int64_t its, orc_sample_cb cb, void *cbctx);

/* ---- pc_chols.c ------------------------------------------------------------------------ */
int  orc_potrf_lower(int64_t n, double *a);
void orc_trsv_lower(int64_t n, const double *l, int trans, double *x);
int  orc_chol_sample(int64_t n, const double *l, orc_noise *ns, const double *b, double *y);

/* ---- sparse helpers + PCMG (Appendix A) + pc_gamgmc.c ------------------------------------ */
void    orc_spmv(int64_t n, const int64_t *rowptr, const int32_t *col, const double *val, const double *x, double *y);
int64_t orc_q1_nnz(int dim, const int64_t nf[3], const int64_t nc[3]);
void    orc_q1_coarse_dims(int dim, const int64_t nf[3], int64_t nc[3]);
void    orc_q1_interp(int dim, const int64_t nf[3], const int64_t nc[3], int64_t *rowptr, int32_t *col, double *val);

typedef struct orc_mg_s orc_mg;
/* level numbering follows PCMG: 0 = coarsest, nlevels-1 = finest */
orc_mg *orc_mg_create(int nlevels);
void    orc_mg_destroy(orc_mg *mg);
int     orc_mg_set_fine(orc_mg *mg, int64_t n, const int64_t *rowptr, const int32_t *col, const double *val);
int     orc_mg_set_interp(orc_mg *mg, int level, int64_t nf, int64_t nc, const int64_t *rowptr, const int32_t *col, const double *val);
int     orc_mg_galerkin(orc_mg *mg);
int     orc_mg_build_geometric(orc_mg *mg, int dim, int64_t nx, int64_t ny, int64_t nz, double kappa);
void    orc_mg_level_dims(const orc_mg *mg, int level, int64_t dims[3]);
int64_t orc_mg_level_n(const orc_mg *mg, int level);
int64_t orc_mg_level_nnz(const orc_mg *mg, int level);
void    orc_mg_level_csr(const orc_mg *mg, int level, int64_t *rowptr, int32_t *col, double *val);
/* smoother: kind 0 = sorgibbs (omega fixed 1, pc_sorgibbs.c), 1 = mcgibbs (pc_mcgibbs.c); colouring NULL => one colour (1-rank reference) */
int orc_mg_set_smoother(orc_mg *mg, int level, int kind, double omega, int type, int its, int ncolors, const int32_t *color);
/* coarse: kind 2 = cholsampler (pc_chols.c dense path), else as above */
int orc_mg_set_coarse(orc_mg *mg, int kind, double omega, int type, int its, int ncolors, const int32_t *color);
int orc_mg_setup(orc_mg *mg);
int orc_mg_apply(orc_mg *mg, orc_noise *ns, const double *b, double *x); /* PCApply_MG: x = 0; V-cycle */
int orc_gamgmc_richardson(orc_mg *mg, orc_noise *ns, const double *b, double *y, int64_t its, int guesszero, orc_sample_cb cb, void *cbctx);

/* ---- iact.c / stats.c / ex7.c ---------------------------------------------------------------- */
void   orc_autocorrelation(int64_t n, const double *x, double *acf);
int    orc_iact(int64_t n, const double *x, double *tau, double *acf_or_null, int *valid);
int    orc_cov_errors(int64_t n, const double *adense_rowmajor, int64_t chains, int64_t samples_per_chain, const double *samples, double *errs);
double orc_gelman_rubin(int64_t n, int64_t chains, int64_t len, const double *samples);

#ifdef __cplusplus
}
#endif
#endif
