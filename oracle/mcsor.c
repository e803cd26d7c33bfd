/* mcsor.c -- oracle restatement of the multicolour SOR engine and the synthetic problem
 * generator.  TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Follows: src/problems.c:14-75, src/mc_sor.c:114-150, :152-214, :216-239, :241-296,
 *          :298-381, :397-410 of /root/reference.
 */
#include "oracle.h"
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------------------------
 * problems.c:14-75  MatAssembleShiftedLaplaceFD.  hinv2 = 1/((mx-1)*(mx-1)) (it is h^2,
 * SURVEY F8, and uses mx for every direction); off-diagonals -hinv2 to existing
 * neighbours; diag = kappa^2 accumulated with += hinv2 once per existing neighbour in the
 * order south, west, north, east (:31-58) -- all increments are equal so only the count
 * matters.  PETSc stores the row with ascending columns.  dim==3 is the 7-point extension
 * SURVEY F8 defines (neighbour order down, south, west, [diag], east, north, up).
 * --------------------------------------------------------------------------------- */
int64_t orc_laplace_nnz(int dim, int64_t nx, int64_t ny, int64_t nz)
{
  if (dim == 2) return nx * ny + 2 * ((nx - 1) * ny + nx * (ny - 1));
  return nx * ny * nz + 2 * ((nx - 1) * ny * nz + nx * (ny - 1) * nz + nx * ny * (nz - 1));
}

void orc_laplace_csr(int dim, int64_t nx, int64_t ny, int64_t nz, double kappa, int64_t *rowptr, int32_t *col, double *val)
{
  const double hinv2 = 1. / (double)((nx - 1) * (nx - 1));
  int64_t      k     = 0;
  if (dim == 2) nz = 1;
  rowptr[0] = 0;
  for (int64_t z = 0; z < nz; ++z)
    for (int64_t j = 0; j < ny; ++j)
      for (int64_t i = 0; i < nx; ++i) {
        const int64_t r    = i + nx * (j + ny * z);
        double        diag = kappa * kappa;
        /* diagonal: one += per existing neighbour (problems.c:31-58) */
        if (dim == 3 && z > 0) diag += hinv2;
        if (j > 0) diag += hinv2;
        if (i > 0) diag += hinv2;
        if (j < ny - 1) diag += hinv2;
        if (i < nx - 1) diag += hinv2;
        if (dim == 3 && z < nz - 1) diag += hinv2;
        /* ascending columns */
        if (dim == 3 && z > 0) { col[k] = (int32_t)(r - nx * ny); val[k++] = -hinv2; }
        if (j > 0) { col[k] = (int32_t)(r - nx); val[k++] = -hinv2; }
        if (i > 0) { col[k] = (int32_t)(r - 1); val[k++] = -hinv2; }
        col[k] = (int32_t)r; val[k++] = diag;
        if (i < nx - 1) { col[k] = (int32_t)(r + 1); val[k++] = -hinv2; }
        if (j < ny - 1) { col[k] = (int32_t)(r + nx); val[k++] = -hinv2; }
        if (dim == 3 && z < nz - 1) { col[k] = (int32_t)(r + nx * ny); val[k++] = -hinv2; }
        rowptr[r + 1] = k;
      }
}

/* mc_sor.c:126-150  MatGetDiagonalPointers: position of the diagonal entry in each row.
 * Returns the number of rows without a diagonal (the reference leaves those uninitialised;
 * we mark them -1 and report). */
int orc_diag_ptrs(int64_t n, const int64_t *rowptr, const int32_t *col, int64_t *diagptr)
{
  int missing = 0;
  for (int64_t r = 0; r < n; ++r) {
    diagptr[r] = -1;
    for (int64_t k = rowptr[r]; k < rowptr[r + 1]; ++k)
      if (col[k] == r) diagptr[r] = k;
    if (diagptr[r] < 0) ++missing;
  }
  return missing;
}

/* mc_sor.c:114-124  MCSORUpdateIDiag: idiag = omega * (1/diag)  (VecReciprocal then VecScale:
 * two roundings). */
void orc_idiag(int64_t n, const double *val, const int64_t *diagptr, double omega, double *idiag)
{
  for (int64_t r = 0; r < n; ++r) {
    double d = val[diagptr[r]];
    d        = 1. / d;
    idiag[r] = d * omega;
  }
}

/* mc_sor.c:260-268 (forward) / :277-285 (backward): one row update.  FP contract in oracle.h. */
static inline void row_update(const int64_t *rowptr, const int32_t *col, const double *val, const int64_t *diagptr, const double *idiag, double omega, const double *b, double *y, int64_t r)
{
  double sum = b[r];
  for (int64_t k = rowptr[r]; k < diagptr[r]; ++k) sum = fma(-val[k], y[col[k]], sum);
  for (int64_t k = diagptr[r] + 1; k < rowptr[r + 1]; ++k) sum = fma(-val[k], y[col[k]], sum);
  const double t = (1. - omega) * y[r];
  y[r]           = fma(idiag[r], sum, t);
}

/* mc_sor.c:241-296  MCSORApply_SEQAIJ, one directional sweep.  Forward: colours ascending,
 * rows in list order; backward: colours descending, rows reversed. */
void orc_sweep_seq(int64_t n, const int64_t *rowptr, const int32_t *col, const double *val, const int64_t *diagptr, const double *idiag, double omega, int ncolors, const int64_t *colorptr, const int32_t *colorrows, int dir, const double *b, double *y)
{
  (void)n;
  if (dir == ORC_SOR_FORWARD) {
    for (int c = 0; c < ncolors; ++c)
      for (int64_t i = colorptr[c]; i < colorptr[c + 1]; ++i) row_update(rowptr, col, val, diagptr, idiag, omega, b, y, colorrows[i]);
  } else {
    for (int c = ncolors - 1; c >= 0; --c)
      for (int64_t i = colorptr[c + 1] - 1; i >= colorptr[c]; --i) row_update(rowptr, col, val, diagptr, idiag, omega, b, y, colorrows[i]);
  }
}

/* mc_sor.c:216-239  MCSORApply: symmetric = forward then backward with the same b. */
void orc_mcsor_apply(int64_t n, const int64_t *rowptr, const int32_t *col, const double *val, const int64_t *diagptr, const double *idiag, double omega, int ncolors, const int64_t *colorptr, const int32_t *colorrows, int type, const double *b, double *y)
{
  if (type == ORC_SOR_SYMMETRIC) {
    orc_sweep_seq(n, rowptr, col, val, diagptr, idiag, omega, ncolors, colorptr, colorrows, ORC_SOR_FORWARD, b, y);
    orc_sweep_seq(n, rowptr, col, val, diagptr, idiag, omega, ncolors, colorptr, colorrows, ORC_SOR_BACKWARD, b, y);
  } else orc_sweep_seq(n, rowptr, col, val, diagptr, idiag, omega, ncolors, colorptr, colorrows, type, b, y);
}

/* ISColoringGetIS semantics (mc_sor.c:251, :392): per-colour lists of local rows, ascending. */
int orc_coloring_lists(int64_t n, const int32_t *color, int ncolors, int64_t *colorptr, int32_t *colorrows)
{
  memset(colorptr, 0, sizeof(int64_t) * (size_t)(ncolors + 1));
  for (int64_t r = 0; r < n; ++r) {
    if (color[r] < 0 || color[r] >= ncolors) return 1;
    colorptr[color[r] + 1]++;
  }
  for (int c = 0; c < ncolors; ++c) colorptr[c + 1] += colorptr[c];
  int64_t *pos = malloc(sizeof(int64_t) * (size_t)ncolors);
  memcpy(pos, colorptr, sizeof(int64_t) * (size_t)ncolors);
  for (int64_t r = 0; r < n; ++r) colorrows[pos[color[r]]++] = (int32_t)r;
  free(pos);
  return 0;
}

/* Deterministic first-fit greedy distance-1 colouring in natural order.  Stands in for
 * MATCOLORINGJP (mc_sor.c:383-395), whose output is PETSc-internal and unpinned. */
int orc_coloring_greedy(int64_t n, const int64_t *rowptr, const int32_t *col, int32_t *color)
{
  int      ncolors = 0, cap = 64;
  int64_t *mark = malloc(sizeof(int64_t) * (size_t)cap);
  for (int i = 0; i < cap; ++i) mark[i] = -1;
  for (int64_t r = 0; r < n; ++r) color[r] = -1;
  for (int64_t r = 0; r < n; ++r) {
    for (int64_t k = rowptr[r]; k < rowptr[r + 1]; ++k) {
      const int32_t c = col[k];
      if (c != r && color[c] >= 0) mark[color[c]] = r;
    }
    int c = 0;
    while (c < ncolors && mark[c] == r) ++c;
    if (c == ncolors) {
      ++ncolors;
      if (ncolors >= cap) {
        mark = realloc(mark, sizeof(int64_t) * (size_t)(2 * cap));
        for (int i = cap; i < 2 * cap; ++i) mark[i] = -1;
        cap *= 2;
      }
    }
    color[r] = c;
  }
  free(mark);
  return ncolors;
}

/* Level-set ("wavefront") colouring: level(r) = 1 + max level of lower-numbered neighbours.
 * Sweeping its colours in ascending order reproduces the 1-colour lexicographic sweep of the
 * 1-rank reference (mc_sor.c:397-410 + :257-271) exactly, row for row. */
int orc_coloring_levelset(int64_t n, const int64_t *rowptr, const int32_t *col, int32_t *color)
{
  int ncolors = 0;
  for (int64_t r = 0; r < n; ++r) {
    int lvl = 0;
    for (int64_t k = rowptr[r]; k < rowptr[r + 1]; ++k)
      if (col[k] < r && color[col[k]] + 1 > lvl) lvl = color[col[k]] + 1;
    color[r] = lvl;
    if (lvl + 1 > ncolors) ncolors = lvl + 1;
  }
  return ncolors;
}

/* distance-1 validity: no stored off-diagonal entry joins two rows of one colour.
 * returns the number of violations */
int orc_coloring_valid(int64_t n, const int64_t *rowptr, const int32_t *col, const double *val, const int32_t *color)
{
  int bad = 0;
  (void)val;
  for (int64_t r = 0; r < n; ++r)
    for (int64_t k = rowptr[r]; k < rowptr[r + 1]; ++k)
      if (col[k] != r && color[col[k]] == color[r]) ++bad;
  return bad;
}

/* ---------------------------------------------------------------------------------
 * Partitioned sweep.  Mirrors MatMPIAIJGetSeqAIJ's split (diag block with local column ids,
 * off-diag block with compressed column ids + colmap), the per-colour ghost lists of
 * MatCreateScatters (mc_sor.c:152-214: for every colour, in colour-row order, one slot per
 * off-diag nonzero, duplicates kept) and the sweep of MCSORApply_MPIAIJ (mc_sor.c:298-381:
 * sum starts at 0, diag-block part, ghost part through a running counter, then
 * y = (1-omega) y + idiag (sum + b)).  Ranks are emulated by threads with a barrier where the
 * reference has VecScatterBegin/End.
 * --------------------------------------------------------------------------------- */
typedef struct {
  int64_t  n, row0;
  int64_t *drowptr, *orowptr, *ddiag;
  int32_t *dcol, *ocol; /* ocol indexes colmap */
  double  *dval, *oval, *idiag;
  int64_t  ncolmap;
  int64_t *colmap; /* global column of each compressed off-diag column */
  int64_t *colorptr;
  int32_t *colorrows;
  int64_t *ghostptr; /* per colour offset into ghostidx */
  int64_t *ghostidx; /* global row to gather, one per off-diag nnz */
  double  *ghostbuf;
} orc_rank;

struct orc_part_s {
  int       nranks, ncolors;
  int64_t   n;
  double    omega;
  orc_rank *r;
};

static int cmp_i64(const void *a, const void *b)
{
  const int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
  return (x > y) - (x < y);
}

orc_part *orc_part_create(int64_t n, const int64_t *rowptr, const int32_t *col, const double *val, int nranks, const int64_t *rowstart, int ncolors, const int32_t *color, double omega)
{
  orc_part *p = calloc(1, sizeof(*p));
  p->nranks   = nranks;
  p->ncolors  = ncolors;
  p->n        = n;
  p->omega    = omega;
  p->r        = calloc((size_t)nranks, sizeof(orc_rank));
  for (int q = 0; q < nranks; ++q) {
    orc_rank     *R  = &p->r[q];
    const int64_t r0 = rowstart[q], r1 = rowstart[q + 1], nl = r1 - r0;
    R->n    = nl;
    R->row0 = r0;
    int64_t dn = 0, on = 0;
    for (int64_t r = r0; r < r1; ++r)
      for (int64_t k = rowptr[r]; k < rowptr[r + 1]; ++k) {
        if (col[k] >= r0 && col[k] < r1) ++dn;
        else ++on;
      }
    R->drowptr = malloc(sizeof(int64_t) * (size_t)(nl + 1));
    R->orowptr = malloc(sizeof(int64_t) * (size_t)(nl + 1));
    R->ddiag   = malloc(sizeof(int64_t) * (size_t)(nl ? nl : 1));
    R->dcol    = malloc(sizeof(int32_t) * (size_t)(dn ? dn : 1));
    R->dval    = malloc(sizeof(double) * (size_t)(dn ? dn : 1));
    R->ocol    = malloc(sizeof(int32_t) * (size_t)(on ? on : 1));
    R->oval    = malloc(sizeof(double) * (size_t)(on ? on : 1));
    R->idiag   = malloc(sizeof(double) * (size_t)(nl ? nl : 1));
    /* colmap: sorted unique global off-process columns (PETSc's garray) */
    int64_t *tmp = malloc(sizeof(int64_t) * (size_t)(on ? on : 1));
    int64_t  t   = 0;
    for (int64_t r = r0; r < r1; ++r)
      for (int64_t k = rowptr[r]; k < rowptr[r + 1]; ++k)
        if (!(col[k] >= r0 && col[k] < r1)) tmp[t++] = col[k];
    qsort(tmp, (size_t)t, sizeof(int64_t), cmp_i64);
    int64_t u = 0;
    for (int64_t i = 0; i < t; ++i)
      if (i == 0 || tmp[i] != tmp[i - 1]) tmp[u++] = tmp[i];
    R->ncolmap = u;
    R->colmap  = tmp;
    int64_t dk = 0, ok = 0;
    R->drowptr[0] = R->orowptr[0] = 0;
    for (int64_t r = r0; r < r1; ++r) {
      R->ddiag[r - r0] = -1;
      for (int64_t k = rowptr[r]; k < rowptr[r + 1]; ++k) {
        if (col[k] >= r0 && col[k] < r1) {
          if (col[k] == r) R->ddiag[r - r0] = dk;
          R->dcol[dk] = (int32_t)(col[k] - r0);
          R->dval[dk] = val[k];
          ++dk;
        } else {
          const int64_t  g  = col[k];
          const int64_t *f  = bsearch(&g, R->colmap, (size_t)u, sizeof(int64_t), cmp_i64);
          R->ocol[ok]       = (int32_t)(f - R->colmap);
          R->oval[ok]       = val[k];
          ++ok;
        }
      }
      R->drowptr[r - r0 + 1] = dk;
      R->orowptr[r - r0 + 1] = ok;
    }
    orc_idiag(nl, R->dval, R->ddiag, omega, R->idiag);
    /* IS_COLORING_LOCAL lists (mc_sor.c:392) */
    R->colorptr  = malloc(sizeof(int64_t) * (size_t)(ncolors + 1));
    R->colorrows = malloc(sizeof(int32_t) * (size_t)(nl ? nl : 1));
    orc_coloring_lists(nl, color + r0, ncolors, R->colorptr, R->colorrows);
    /* MatCreateScatters mc_sor.c:171-205 */
    R->ghostptr = malloc(sizeof(int64_t) * (size_t)(ncolors + 1));
    R->ghostidx = malloc(sizeof(int64_t) * (size_t)(on ? on : 1));
    R->ghostbuf = malloc(sizeof(double) * (size_t)(on ? on : 1));
    int64_t cnt = 0;
    for (int c = 0; c < ncolors; ++c) {
      R->ghostptr[c] = cnt;
      for (int64_t i = R->colorptr[c]; i < R->colorptr[c + 1]; ++i) {
        const int64_t lr = R->colorrows[i];
        for (int64_t k = R->orowptr[lr]; k < R->orowptr[lr + 1]; ++k) R->ghostidx[cnt++] = R->colmap[R->ocol[k]];
      }
    }
    R->ghostptr[ncolors] = cnt;
  }
  return p;
}

void orc_part_destroy(orc_part *p)
{
  if (!p) return;
  for (int q = 0; q < p->nranks; ++q) {
    orc_rank *R = &p->r[q];
    free(R->drowptr); free(R->orowptr); free(R->ddiag); free(R->dcol); free(R->dval); free(R->ocol); free(R->oval);
    free(R->idiag); free(R->colmap); free(R->colorptr); free(R->colorrows); free(R->ghostptr); free(R->ghostidx); free(R->ghostbuf);
  }
  free(p->r);
  free(p);
}

int64_t orc_part_ghost_count(const orc_part *p, int rank, int color) { return p->r[rank].ghostptr[color + 1] - p->r[rank].ghostptr[color]; }
void    orc_part_ghost_index(const orc_part *p, int rank, int color, int64_t *out)
{
  const orc_rank *R = &p->r[rank];
  memcpy(out, R->ghostidx + R->ghostptr[color], sizeof(int64_t) * (size_t)(R->ghostptr[color + 1] - R->ghostptr[color]));
}

/* one rank, one colour: gather (VecScatter, mc_sor.c:318-319) */
static void rank_gather(const orc_part *p, int q, int c, const double *y)
{
  (void)p;
  orc_rank *R = &p->r[q];
  for (int64_t g = R->ghostptr[c]; g < R->ghostptr[c + 1]; ++g) R->ghostbuf[g] = y[R->ghostidx[g]];
}

/* one rank, one colour: rows (mc_sor.c:326-335 forward, :358-369 backward) */
static void rank_rows(const orc_part *p, int q, int c, int dir, const double *b, double *y)
{
  orc_rank     *R     = &p->r[q];
  const double  omega = p->omega;
  double       *yl    = y + R->row0;
  const double *bl    = b + R->row0;
  const double *ghost = R->ghostbuf + R->ghostptr[c];
  if (dir == ORC_SOR_FORWARD) {
    int64_t gcnt = 0;
    for (int64_t i = R->colorptr[c]; i < R->colorptr[c + 1]; ++i) {
      const int64_t r   = R->colorrows[i];
      double        sum = 0;
      for (int64_t k = R->drowptr[r]; k < R->ddiag[r]; ++k) sum = fma(-R->dval[k], yl[R->dcol[k]], sum);
      for (int64_t k = R->ddiag[r] + 1; k < R->drowptr[r + 1]; ++k) sum = fma(-R->dval[k], yl[R->dcol[k]], sum);
      for (int64_t k = R->orowptr[r]; k < R->orowptr[r + 1]; ++k) sum = fma(-R->oval[k], ghost[gcnt++], sum);
      const double t = (1 - omega) * yl[r];
      yl[r]          = fma(R->idiag[r], sum + bl[r], t);
    }
  } else {
    int64_t gcnt = R->ghostptr[c + 1] - R->ghostptr[c];
    for (int64_t i = R->colorptr[c + 1] - 1; i >= R->colorptr[c]; --i) {
      const int64_t r   = R->colorrows[i];
      double        sum = 0;
      gcnt -= R->orowptr[r + 1] - R->orowptr[r];
      for (int64_t k = R->drowptr[r]; k < R->ddiag[r]; ++k) sum = fma(-R->dval[k], yl[R->dcol[k]], sum);
      for (int64_t k = R->ddiag[r] + 1; k < R->drowptr[r + 1]; ++k) sum = fma(-R->dval[k], yl[R->dcol[k]], sum);
      int64_t go = gcnt;
      for (int64_t k = R->orowptr[r]; k < R->orowptr[r + 1]; ++k) sum = fma(-R->oval[k], ghost[go++], sum);
      const double t = (1 - omega) * yl[r];
      yl[r]          = fma(R->idiag[r], sum + bl[r], t);
    }
  }
}

typedef struct {
  const orc_part    *p;
  int                tid, nthreads, dir;
  const double      *b;
  double            *y;
  pthread_barrier_t *bar;
} part_job;

static void *part_worker(void *arg)
{
  part_job       *j = arg;
  const orc_part *p = j->p;
  for (int s = 0; s < p->ncolors; ++s) {
    const int c = j->dir == ORC_SOR_FORWARD ? s : p->ncolors - 1 - s;
    for (int q = j->tid; q < p->nranks; q += j->nthreads) rank_gather(p, q, c, j->y);
    if (j->bar) pthread_barrier_wait(j->bar);
    for (int q = j->tid; q < p->nranks; q += j->nthreads) rank_rows(p, q, c, j->dir, j->b, j->y);
    if (j->bar) pthread_barrier_wait(j->bar);
  }
  return NULL;
}

void orc_part_sweep(orc_part *p, int dir, const double *b, double *y, int nthreads)
{
  if (dir == ORC_SOR_SYMMETRIC) {
    orc_part_sweep(p, ORC_SOR_FORWARD, b, y, nthreads);
    orc_part_sweep(p, ORC_SOR_BACKWARD, b, y, nthreads);
    return;
  }
  if (nthreads > p->nranks) nthreads = p->nranks;
  if (nthreads <= 1) {
    part_job j = {p, 0, 1, dir, b, y, NULL};
    part_worker(&j);
    return;
  }
  pthread_barrier_t bar;
  pthread_barrier_init(&bar, NULL, (unsigned)nthreads);
  pthread_t *th = malloc(sizeof(pthread_t) * (size_t)nthreads);
  part_job  *jb = malloc(sizeof(part_job) * (size_t)nthreads);
  for (int t = 0; t < nthreads; ++t) {
    jb[t] = (part_job){p, t, nthreads, dir, b, y, &bar};
    pthread_create(&th[t], NULL, part_worker, &jb[t]);
  }
  for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
  pthread_barrier_destroy(&bar);
  free(th);
  free(jb);
}
