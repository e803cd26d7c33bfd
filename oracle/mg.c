/* mg.c -- oracle restatement of the MGMC V-cycle.  TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Follows src/pc_gamgmc.c:227-264 (outer Richardson), :296-350 (defaults) of /root/reference and,
 * for the PETSc-internal parts that are not in /root/reference, SURVEY.md Appendix A:
 *   A.3 PCApply_MG / PCMGMCycle_Private (multiplicative V-cycle, R = P^T, Galerkin A_c = P^T A P)
 *   A.4 DMDA Q1 interpolation (vertex centred, ratio 2)
 * The PETSc pieces are "parity unpinned" against PETSc itself; they are validated by the exact
 * invariance identity in tests/test_oracle.py.
 *
 * FP contract for the transfer/residual ops (shared with the CUDA path):
 *   residual      ax = 0; ax = fma(a_k, x[c_k], ax) k ascending;  r = b - ax     (MatMult; VecAYPX)
 *   restriction   bc_J = 0; bc_J = fma(P_iJ, r_i, bc_J), fine rows i ascending    (MatMultTranspose)
 *   prolongation  s = x_i; s = fma(P_iJ, xc_J, s), coarse cols J ascending        (MatMultAdd)
 */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int64_t  n, m; /* rows, cols */
  int64_t *rowptr;
  int32_t *col;
  double  *val;
} csr;

static void csr_free(csr *a)
{
  free(a->rowptr); free(a->col); free(a->val);
  memset(a, 0, sizeof(*a));
}

static csr csr_copy(int64_t n, int64_t m, const int64_t *rowptr, const int32_t *col, const double *val)
{
  csr           a   = {n, m, NULL, NULL, NULL};
  const int64_t nnz = rowptr[n];
  a.rowptr          = malloc(sizeof(int64_t) * (size_t)(n + 1));
  a.col             = malloc(sizeof(int32_t) * (size_t)(nnz ? nnz : 1));
  a.val             = malloc(sizeof(double) * (size_t)(nnz ? nnz : 1));
  memcpy(a.rowptr, rowptr, sizeof(int64_t) * (size_t)(n + 1));
  memcpy(a.col, col, sizeof(int32_t) * (size_t)nnz);
  memcpy(a.val, val, sizeof(double) * (size_t)nnz);
  return a;
}

static csr csr_transpose(const csr *a)
{
  csr           t   = {a->m, a->n, NULL, NULL, NULL};
  const int64_t nnz = a->rowptr[a->n];
  t.rowptr          = calloc((size_t)(a->m + 1), sizeof(int64_t));
  t.col             = malloc(sizeof(int32_t) * (size_t)(nnz ? nnz : 1));
  t.val             = malloc(sizeof(double) * (size_t)(nnz ? nnz : 1));
  for (int64_t k = 0; k < nnz; ++k) t.rowptr[a->col[k] + 1]++;
  for (int64_t c = 0; c < a->m; ++c) t.rowptr[c + 1] += t.rowptr[c];
  int64_t *pos = malloc(sizeof(int64_t) * (size_t)(a->m ? a->m : 1));
  memcpy(pos, t.rowptr, sizeof(int64_t) * (size_t)a->m);
  for (int64_t r = 0; r < a->n; ++r)
    for (int64_t k = a->rowptr[r]; k < a->rowptr[r + 1]; ++k) {
      const int64_t q = pos[a->col[k]]++;
      t.col[q]        = (int32_t)r;
      t.val[q]        = a->val[k];
    }
  free(pos);
  return t;
}

/* C = A B, Gustavson, rows sorted by column */
static csr csr_matmul(const csr *a, const csr *b)
{
  csr      c     = {a->n, b->m, NULL, NULL, NULL};
  int64_t *mark  = malloc(sizeof(int64_t) * (size_t)(b->m ? b->m : 1));
  double  *acc   = malloc(sizeof(double) * (size_t)(b->m ? b->m : 1));
  int32_t *list  = malloc(sizeof(int32_t) * (size_t)(b->m ? b->m : 1));
  int64_t  cap   = a->rowptr[a->n] * 3 + 16, nnz = 0;
  c.rowptr       = malloc(sizeof(int64_t) * (size_t)(a->n + 1));
  c.col          = malloc(sizeof(int32_t) * (size_t)cap);
  c.val          = malloc(sizeof(double) * (size_t)cap);
  for (int64_t j = 0; j < b->m; ++j) mark[j] = -1;
  c.rowptr[0] = 0;
  for (int64_t i = 0; i < a->n; ++i) {
    int64_t cnt = 0;
    for (int64_t k = a->rowptr[i]; k < a->rowptr[i + 1]; ++k) {
      const int64_t kk = a->col[k];
      const double  av = a->val[k];
      for (int64_t l = b->rowptr[kk]; l < b->rowptr[kk + 1]; ++l) {
        const int32_t j = b->col[l];
        if (mark[j] != i) {
          mark[j]     = i;
          acc[j]      = 0;
          list[cnt++] = j;
        }
        acc[j] = fma(av, b->val[l], acc[j]);
      }
    }
    for (int64_t x = 1; x < cnt; ++x) { /* insertion sort: rows are short */
      const int32_t v = list[x];
      int64_t       y = x - 1;
      while (y >= 0 && list[y] > v) { list[y + 1] = list[y]; --y; }
      list[y + 1] = v;
    }
    if (nnz + cnt > cap) {
      cap   = (nnz + cnt) * 2;
      c.col = realloc(c.col, sizeof(int32_t) * (size_t)cap);
      c.val = realloc(c.val, sizeof(double) * (size_t)cap);
    }
    for (int64_t x = 0; x < cnt; ++x) {
      c.col[nnz]   = list[x];
      c.val[nnz++] = acc[list[x]];
    }
    c.rowptr[i + 1] = nnz;
  }
  free(mark); free(acc); free(list);
  return c;
}

void orc_spmv(int64_t n, const int64_t *rowptr, const int32_t *col, const double *val, const double *x, double *y)
{
  for (int64_t r = 0; r < n; ++r) {
    double s = 0;
    for (int64_t k = rowptr[r]; k < rowptr[r + 1]; ++k) s = fma(val[k], x[col[k]], s);
    y[r] = s;
  }
}

/* ---- Appendix A.4: Q1 interpolation.  nc = (nf+1)/2 per direction: for nf = 2^k+1 this is
 * PETSc's M_c = 1 + (M_f-1)/2; for even nf (512^3) the last fine node keeps only its
 * existing left coarse neighbour (any full-rank P with Galerkin A_c leaves the sampler exact). */
void orc_q1_coarse_dims(int dim, const int64_t nf[3], int64_t nc[3])
{
  for (int d = 0; d < 3; ++d) nc[d] = (d < dim && nf[d] > 1) ? (nf[d] + 1) / 2 : 1;
}

static int q1_1d(int64_t i, int64_t nc, int64_t idx[2], double w[2])
{
  if (nc == 1 && i == 0) { idx[0] = 0; w[0] = 1; return 1; }
  if ((i & 1) == 0) { idx[0] = i / 2; w[0] = 1; return 1; }
  int n = 0;
  idx[n] = (i - 1) / 2; w[n++] = 0.5;
  if ((i + 1) / 2 < nc) { idx[n] = (i + 1) / 2; w[n++] = 0.5; }
  return n;
}

int64_t orc_q1_nnz(int dim, const int64_t nf[3], const int64_t nc[3])
{
  int64_t tot = 1;
  for (int d = 0; d < 3; ++d) {
    int64_t s = 0, idx[2];
    double  w[2];
    if (d >= dim || nf[d] == 1) continue;
    for (int64_t i = 0; i < nf[d]; ++i) s += q1_1d(i, nc[d], idx, w);
    tot *= s;
  }
  return tot;
}

void orc_q1_interp(int dim, const int64_t nf[3], const int64_t nc[3], int64_t *rowptr, int32_t *col, double *val)
{
  int64_t k = 0;
  rowptr[0] = 0;
  for (int64_t z = 0; z < nf[2]; ++z)
    for (int64_t j = 0; j < nf[1]; ++j)
      for (int64_t i = 0; i < nf[0]; ++i) {
        int64_t ix[2], iy[2] = {0, 0}, iz[2] = {0, 0};
        double  wx[2], wy[2] = {1, 0}, wz[2] = {1, 0};
        const int cx = q1_1d(i, nc[0], ix, wx);
        const int cy = (dim >= 2 && nf[1] > 1) ? q1_1d(j, nc[1], iy, wy) : 1;
        const int cz = (dim >= 3 && nf[2] > 1) ? q1_1d(z, nc[2], iz, wz) : 1;
        for (int c = 0; c < cz; ++c)
          for (int b = 0; b < cy; ++b)
            for (int a = 0; a < cx; ++a) {
              col[k]   = (int32_t)(ix[a] + nc[0] * (iy[b] + nc[1] * iz[c]));
              val[k++] = wx[a] * wy[b] * wz[c];
            }
        rowptr[(i + nf[0] * (j + nf[1] * z)) + 1] = k;
      }
}

/* ---- the hierarchy ------------------------------------------------------------------------- */
typedef struct {
  csr      A, P, R; /* P: n_l x n_{l-1} (level >= 1), R = P^T */
  int64_t  dims[3];
  int      kind;  /* 0 sorgibbs, 1 mcgibbs, 2 cholsampler */
  double   omega;
  int      type, its, ncolors;
  int64_t *colorptr;
  int32_t *colorrows;
  int64_t *diagptr;
  double  *idiag, *sqrtdiag, *L;
  double  *b, *x, *r, *w, *z;
} level;

struct orc_mg_s {
  int    nlevels;
  level *lv;
  double *work, *wfine;
};

orc_mg *orc_mg_create(int nlevels)
{
  orc_mg *mg  = calloc(1, sizeof(*mg));
  mg->nlevels = nlevels;
  mg->lv      = calloc((size_t)nlevels, sizeof(level));
  for (int l = 0; l < nlevels; ++l) { /* pc_gamgmc.c:305-349 defaults */
    mg->lv[l].kind  = l == 0 ? 2 : 0;
    mg->lv[l].omega = 1;
    mg->lv[l].type  = ORC_SOR_FORWARD;
    mg->lv[l].its   = 1;
  }
  return mg;
}

void orc_mg_destroy(orc_mg *mg)
{
  if (!mg) return;
  for (int l = 0; l < mg->nlevels; ++l) {
    level *v = &mg->lv[l];
    csr_free(&v->A); csr_free(&v->P); csr_free(&v->R);
    free(v->colorptr); free(v->colorrows); free(v->diagptr); free(v->idiag); free(v->sqrtdiag); free(v->L);
    free(v->b); free(v->x); free(v->r); free(v->w); free(v->z);
  }
  free(mg->lv); free(mg->work); free(mg->wfine);
  free(mg);
}

int orc_mg_set_fine(orc_mg *mg, int64_t n, const int64_t *rowptr, const int32_t *col, const double *val)
{
  level *v = &mg->lv[mg->nlevels - 1];
  csr_free(&v->A);
  v->A = csr_copy(n, n, rowptr, col, val);
  return 0;
}

int orc_mg_set_interp(orc_mg *mg, int level_, int64_t nf, int64_t nc, const int64_t *rowptr, const int32_t *col, const double *val)
{
  if (level_ < 1 || level_ >= mg->nlevels) return 1;
  level *v = &mg->lv[level_];
  csr_free(&v->P); csr_free(&v->R);
  v->P = csr_copy(nf, nc, rowptr, col, val);
  v->R = csr_transpose(&v->P);
  return 0;
}

/* -pc_mg_galerkin both: A_{l-1} = P_l^T A_l P_l (Appendix A.3) */
int orc_mg_galerkin(orc_mg *mg)
{
  for (int l = mg->nlevels - 1; l >= 1; --l) {
    level *f = &mg->lv[l], *c = &mg->lv[l - 1];
    if (!f->A.rowptr || !f->P.rowptr) return 1;
    csr ap = csr_matmul(&f->A, &f->P);
    csr_free(&c->A);
    c->A = csr_matmul(&f->R, &ap);
    csr_free(&ap);
  }
  return 0;
}

int orc_mg_build_geometric(orc_mg *mg, int dim, int64_t nx, int64_t ny, int64_t nz, double kappa)
{
  const int L = mg->nlevels;
  level    *f = &mg->lv[L - 1];
  f->dims[0] = nx; f->dims[1] = ny; f->dims[2] = dim == 3 ? nz : 1;
  const int64_t n = f->dims[0] * f->dims[1] * f->dims[2], nnz = orc_laplace_nnz(dim, nx, ny, nz);
  csr_free(&f->A);
  f->A.n = f->A.m = n;
  f->A.rowptr     = malloc(sizeof(int64_t) * (size_t)(n + 1));
  f->A.col        = malloc(sizeof(int32_t) * (size_t)nnz);
  f->A.val        = malloc(sizeof(double) * (size_t)nnz);
  orc_laplace_csr(dim, nx, ny, nz, kappa, f->A.rowptr, f->A.col, f->A.val);
  for (int l = L - 1; l >= 1; --l) {
    level *v = &mg->lv[l], *c = &mg->lv[l - 1];
    orc_q1_coarse_dims(dim, v->dims, c->dims);
    const int64_t nf = v->dims[0] * v->dims[1] * v->dims[2], nc = c->dims[0] * c->dims[1] * c->dims[2];
    if (nc == nf) return 2; /* cannot coarsen further */
    const int64_t pn = orc_q1_nnz(dim, v->dims, c->dims);
    csr_free(&v->P); csr_free(&v->R);
    v->P.n = nf; v->P.m = nc;
    v->P.rowptr = malloc(sizeof(int64_t) * (size_t)(nf + 1));
    v->P.col    = malloc(sizeof(int32_t) * (size_t)pn);
    v->P.val    = malloc(sizeof(double) * (size_t)pn);
    orc_q1_interp(dim, v->dims, c->dims, v->P.rowptr, v->P.col, v->P.val);
    v->R = csr_transpose(&v->P);
  }
  return orc_mg_galerkin(mg);
}

void    orc_mg_level_dims(const orc_mg *mg, int l, int64_t dims[3]) { memcpy(dims, mg->lv[l].dims, sizeof(int64_t) * 3); }
int64_t orc_mg_level_n(const orc_mg *mg, int l) { return mg->lv[l].A.n; }
int64_t orc_mg_level_nnz(const orc_mg *mg, int l) { return mg->lv[l].A.rowptr[mg->lv[l].A.n]; }
void    orc_mg_level_csr(const orc_mg *mg, int l, int64_t *rowptr, int32_t *col, double *val)
{
  const csr *a = &mg->lv[l].A;
  memcpy(rowptr, a->rowptr, sizeof(int64_t) * (size_t)(a->n + 1));
  memcpy(col, a->col, sizeof(int32_t) * (size_t)a->rowptr[a->n]);
  memcpy(val, a->val, sizeof(double) * (size_t)a->rowptr[a->n]);
}

static int set_sampler(level *v, int kind, double omega, int type, int its, int ncolors, const int32_t *color)
{
  v->kind  = kind;
  v->omega = kind == 0 ? 1. : omega; /* pc_sorgibbs.c: no omega option, fixed 1 */
  v->type  = type;
  v->its   = its;
  free(v->colorptr); free(v->colorrows);
  v->colorptr = NULL; v->colorrows = NULL;
  v->ncolors  = 0;
  if (kind != 2) {
    const int64_t n = v->A.n;
    if (n <= 0) return 1;
    v->ncolors   = color ? ncolors : 1; /* mc_sor.c:397-410: one colour on one rank */
    v->colorptr  = malloc(sizeof(int64_t) * (size_t)(v->ncolors + 1));
    v->colorrows = malloc(sizeof(int32_t) * (size_t)n);
    if (color) return orc_coloring_lists(n, color, ncolors, v->colorptr, v->colorrows);
    v->colorptr[0] = 0; v->colorptr[1] = n;
    for (int64_t r = 0; r < n; ++r) v->colorrows[r] = (int32_t)r;
  }
  return 0;
}

int orc_mg_set_smoother(orc_mg *mg, int l, int kind, double omega, int type, int its, int ncolors, const int32_t *color)
{
  if (l < 1 || l >= mg->nlevels || kind == 2) return 1;
  return set_sampler(&mg->lv[l], kind, omega, type, its, ncolors, color);
}
int orc_mg_set_coarse(orc_mg *mg, int kind, double omega, int type, int its, int ncolors, const int32_t *color) { return set_sampler(&mg->lv[0], kind, omega, type, its, ncolors, color); }

int orc_mg_setup(orc_mg *mg)
{
  for (int l = 0; l < mg->nlevels; ++l) {
    level        *v = &mg->lv[l];
    const int64_t n = v->A.n;
    if (!v->A.rowptr) return 1;
    free(v->diagptr); free(v->idiag); free(v->sqrtdiag); free(v->L);
    free(v->b); free(v->x); free(v->r); free(v->w); free(v->z);
    v->L = NULL;
    v->diagptr  = malloc(sizeof(int64_t) * (size_t)n);
    v->idiag    = malloc(sizeof(double) * (size_t)n);
    v->sqrtdiag = malloc(sizeof(double) * (size_t)n);
    v->b = calloc((size_t)n, sizeof(double)); v->x = calloc((size_t)n, sizeof(double));
    v->r = calloc((size_t)n, sizeof(double)); v->w = calloc((size_t)n, sizeof(double)); v->z = calloc((size_t)n, sizeof(double));
    if (orc_diag_ptrs(n, v->A.rowptr, v->A.col, v->diagptr)) return 2;
    if (v->kind == 2) { /* pc_chols.c:173-195: dense column-major copy, potrf "L" */
      v->L = calloc((size_t)(n * n), sizeof(double));
      for (int64_t r = 0; r < n; ++r)
        for (int64_t k = v->A.rowptr[r]; k < v->A.rowptr[r + 1]; ++k) v->L[r + (int64_t)v->A.col[k] * n] = v->A.val[k];
      if (orc_potrf_lower(n, v->L)) return 3;
    } else {
      if (!v->colorptr && set_sampler(v, v->kind, v->omega, v->type, v->its, 0, NULL)) return 4;
      orc_idiag(n, v->A.val, v->diagptr, v->omega, v->idiag);
      orc_sqrtdiag(n, v->A.val, v->diagptr, v->omega, v->sqrtdiag);
    }
  }
  const int64_t nf = mg->lv[mg->nlevels - 1].A.n;
  free(mg->work); free(mg->wfine);
  mg->work  = calloc((size_t)nf, sizeof(double));
  mg->wfine = calloc((size_t)nf, sizeof(double));
  return 0;
}

/* one level KSP(richardson, max_it = its) around the level sampler: pc_sorgibbs.c:115-134 /
 * pc_mcgibbs.c:155-188 / pc_chols.c:293-342 (its == 1 -> PCApply_CholSampler :262-291) */
static int level_sample(level *v, orc_noise *ns)
{
  const int64_t n = v->A.n;
  if (v->kind == 2) {
    if (v->its == 1) return orc_chol_sample(n, v->L, ns, v->b, v->x);
    /* its > 1: forward solve cached once (pc_chols.c:306-336) */
    double *vc = malloc(sizeof(double) * (size_t)n);
    int     err = 0;
    memcpy(vc, v->b, sizeof(double) * (size_t)n);
    orc_trsv_lower(n, v->L, 0, vc);
    for (int it = 0; it < v->its && !err; ++it) {
      err = orc_noise_fill(ns, 0, n, v->z);
      for (int64_t i = 0; i < n; ++i) v->x[i] = vc[i] + v->z[i];
      orc_trsv_lower(n, v->L, 1, v->x);
    }
    free(vc);
    return err;
  }
  for (int it = 0; it < v->its; ++it) {
    const int nsw = v->type == ORC_SOR_SYMMETRIC ? 2 : 1;
    for (int s = 0; s < nsw; ++s) {
      const int dir = v->type == ORC_SOR_SYMMETRIC ? (s == 0 ? ORC_SOR_FORWARD : ORC_SOR_BACKWARD) : v->type;
      if (orc_noise_fill(ns, 0, n, v->z)) return 1;
      orc_prepare_rhs(n, v->b, v->sqrtdiag, v->z, v->w);
      orc_sweep_seq(n, v->A.rowptr, v->A.col, v->A.val, v->diagptr, v->idiag, v->omega, v->ncolors, v->colorptr, v->colorrows, dir, v->w, v->x);
    }
  }
  return 0;
}

/* PCMGMCycle_Private, V-cycle (Appendix A.3) */
static int mcycle(orc_mg *mg, int l, orc_noise *ns)
{
  level *v = &mg->lv[l];
  int    err;
  if ((err = level_sample(v, ns))) return err;
  if (l == 0) return 0;
  level        *c  = &mg->lv[l - 1];
  const int64_t nf = v->A.n, nc = c->A.n;
  for (int64_t r = 0; r < nf; ++r) { /* r = b - A x */
    double s = 0;
    for (int64_t k = v->A.rowptr[r]; k < v->A.rowptr[r + 1]; ++k) s = fma(v->A.val[k], v->x[v->A.col[k]], s);
    v->r[r] = v->b[r] - s;
  }
  for (int64_t J = 0; J < nc; ++J) { /* b_c = P^T r, fine rows ascending */
    double s = 0;
    for (int64_t k = v->R.rowptr[J]; k < v->R.rowptr[J + 1]; ++k) s = fma(v->R.val[k], v->r[v->R.col[k]], s);
    c->b[J] = s;
  }
  memset(c->x, 0, sizeof(double) * (size_t)nc);
  if ((err = mcycle(mg, l - 1, ns))) return err;
  for (int64_t i = 0; i < nf; ++i) { /* x += P x_c */
    double s = v->x[i];
    for (int64_t k = v->P.rowptr[i]; k < v->P.rowptr[i + 1]; ++k) s = fma(v->P.val[k], c->x[v->P.col[k]], s);
    v->x[i] = s;
  }
  return level_sample(v, ns);
}

/* PCApply_MG: x = 0, one V-cycle */
int orc_mg_apply(orc_mg *mg, orc_noise *ns, const double *b, double *x)
{
  level        *f = &mg->lv[mg->nlevels - 1];
  const int64_t n = f->A.n;
  memcpy(f->b, b, sizeof(double) * (size_t)n);
  memset(f->x, 0, sizeof(double) * (size_t)n);
  const int err = mcycle(mg, mg->nlevels - 1, ns);
  memcpy(x, f->x, sizeof(double) * (size_t)n);
  return err;
}

/* pc_gamgmc.c:227-264 PCApplyRichardson_GAMGMC */
int orc_gamgmc_richardson(orc_mg *mg, orc_noise *ns, const double *b, double *y, int64_t its, int guesszero, orc_sample_cb cb, void *cbctx)
{
  level        *f   = &mg->lv[mg->nlevels - 1];
  const int64_t n   = f->A.n;
  int           err = 0;
  for (int64_t it = 0; it < its && !err; ++it) {
    if (it == 0 && guesszero) {
      err = orc_mg_apply(mg, ns, b, y); /* :246 */
    } else {
      orc_spmv(n, f->A.rowptr, f->A.col, f->A.val, y, mg->wfine);        /* :253 MatMult */
      for (int64_t i = 0; i < n; ++i) mg->wfine[i] = b[i] - mg->wfine[i]; /* :254 VecAYPX(w,-1,b) */
      err = orc_mg_apply(mg, ns, mg->wfine, mg->work);                   /* :255 */
      for (int64_t i = 0; i < n; ++i) y[i] = y[i] + mg->work[i];         /* :256 */
    }
    if (!err && cb) err = cb(it, y, cbctx);
  }
  return err;
}
