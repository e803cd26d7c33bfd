// csr_op.cu -- general sparse operator: the fused multicolour Gibbs/SOR sweep on a sliced-ELL
// (SELL-32) copy of the CSR matrix, plus residual / SpMV / grid-transfer kernels.
//
// Reference loops being replaced (one kernel each):
//   src/mc_sor.c:260-268, :277-285   row update of MCSORApply_SEQAIJ           -> sell_sweep_kernel
//   src/pc_mcgibbs.c:124-126 + src/parmgmc.c:100-110  noise-perturbed rhs      -> fused into it
//   src/pc_gamgmc.c:253-254, PCMG residual                        -> sell_apply_kernel<RESIDUAL>
//   MatRestrict / MatInterpolateAdd                               -> sell_apply_kernel<SPMV/ADD>
//
// Layout: rows of one colour are stored together in slices of 32 rows; inside a slice the k-th
// off-diagonal entries of the 32 rows are adjacent (col/val loads are fully coalesced 128/256-byte
// transactions, one row per lane, no warp reduction needed).  Slices never straddle colours.
// The accumulation order inside a row is the CSR order, so results are bit-identical to the
// sequential row update (FP contract: oracle/oracle.h).
#include <algorithm>
#include <numeric>

#include "common.hpp"
#include "philox.cuh"

namespace {

struct Sell {
  int64_t         nslices = 0;
  DevBuf<int64_t> slice_off; // nslices + 1, in elements
  DevBuf<int32_t> col;
  DevBuf<double>  val;
  DevBuf<int32_t> rows; // nslices*32 output row of each lane, -1 = padding lane
};

// rows: list of rows to store in order (already padded with -1 to a multiple of 32)
int build_sell(pmg_ctx ctx, const HostCsr &a, const std::vector<int32_t> &rows, bool skip_diag, Sell &s)
{
  const int64_t        nsl = (int64_t)rows.size() / 32;
  std::vector<int64_t> off((size_t)nsl + 1, 0);
  for (int64_t sl = 0; sl < nsl; ++sl) {
    int64_t w = 0;
    for (int l = 0; l < 32; ++l) {
      const int32_t r = rows[(size_t)sl * 32 + l];
      if (r < 0) continue;
      int64_t len = a.rowptr[r + 1] - a.rowptr[r];
      if (skip_diag) {
        for (int64_t k = a.rowptr[r]; k < a.rowptr[r + 1]; ++k)
          if (a.col[k] == r) { --len; break; }
      }
      w = std::max(w, len);
    }
    off[sl + 1] = off[sl] + w * 32;
  }
  std::vector<int32_t> col((size_t)off[nsl]);
  std::vector<double>  val((size_t)off[nsl], 0.0);
  for (int64_t sl = 0; sl < nsl; ++sl) {
    const int64_t w = (off[sl + 1] - off[sl]) / 32;
    for (int l = 0; l < 32; ++l) {
      const int32_t r   = rows[(size_t)sl * 32 + l];
      int64_t       cnt = 0;
      int32_t       pad = 0;
      if (r >= 0) {
        bool skipped = false;
        for (int64_t k = a.rowptr[r]; k < a.rowptr[r + 1]; ++k) {
          if (skip_diag && !skipped && a.col[k] == r) { skipped = true; continue; }
          col[(size_t)(off[sl] + cnt * 32 + l)] = a.col[k];
          val[(size_t)(off[sl] + cnt * 32 + l)] = a.val[k];
          ++cnt;
        }
        pad = skip_diag ? r : (a.rowptr[r + 1] > a.rowptr[r] ? a.col[a.rowptr[r]] : 0);
      }
      for (; cnt < w; ++cnt) col[(size_t)(off[sl] + cnt * 32 + l)] = pad; // value 0: fma(-0, y, s) == s
    }
  }
  s.nslices = nsl;
  PMG_TRY(s.slice_off.upload(off, ctx->stream));
  PMG_TRY(s.col.upload(col, ctx->stream));
  PMG_TRY(s.val.upload(val, ctx->stream));
  PMG_TRY(s.rows.upload(rows, ctx->stream));
  PMG_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// ---- the fused colour sweep ---------------------------------------------------------------------
// one warp per slice, one row per lane
__global__ void __launch_bounds__(256) sell_sweep_kernel(const int64_t *__restrict__ slice_off, const int32_t *__restrict__ col, const double *__restrict__ val, const int32_t *__restrict__ rows, const double *__restrict__ idiag, const double *__restrict__ sqrtdiag, const double *__restrict__ b, double *__restrict__ y, double one_minus_omega, NoiseArgs na, int64_t slice0, int64_t nslices)
{
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int     lane = threadIdx.x & 31;
  if (warp >= nslices) return;
  const int64_t s   = slice0 + warp;
  const int64_t p   = s * 32 + lane;
  const int32_t r   = rows[p];
  const int64_t off = slice_off[s];
  const int     w   = (int)((slice_off[s + 1] - off) >> 5);
  if (r < 0) return;
  double sum = noisy_rhs(na, r, sqrtdiag[p], b ? b[r] : 0.0);
  const int32_t *cp = col + off + lane;
  const double  *vp = val + off + lane;
#pragma unroll 4
  for (int k = 0; k < w; ++k) sum = fma(-vp[(int64_t)k * 32], y[cp[(int64_t)k * 32]], sum);
  const double t = __dmul_rn(one_minus_omega, y[r]);
  y[r]           = fma(idiag[p], sum, t);
}

enum ApplyMode { MODE_SPMV = 0, MODE_RESIDUAL = 1, MODE_ADD = 2 };

// out_r = sum_k a_k x[c_k]          (SPMV: MatMult / MatMultTranspose with the transposed matrix)
// out_r = b_r - sum_k a_k x[c_k]    (RESIDUAL: MatMult then VecAYPX(-1, b))
// out_r = out_r + sum ... started from out_r (ADD: MatMultAdd, MatInterpolateAdd)
template <int MODE> __global__ void __launch_bounds__(256) sell_apply_kernel(const int64_t *__restrict__ slice_off, const int32_t *__restrict__ col, const double *__restrict__ val, const int32_t *__restrict__ rows, const double *__restrict__ x, const double *__restrict__ b, double *__restrict__ out, int64_t nslices)
{
  const int64_t s    = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int     lane = threadIdx.x & 31;
  if (s >= nslices) return;
  const int32_t r = rows[s * 32 + lane];
  if (r < 0) return;
  const int64_t  off = slice_off[s];
  const int      w   = (int)((slice_off[s + 1] - off) >> 5);
  const int32_t *cp  = col + off + lane;
  const double  *vp  = val + off + lane;
  double         sum = MODE == MODE_ADD ? out[r] : 0.0;
#pragma unroll 4
  for (int k = 0; k < w; ++k) sum = fma(vp[(int64_t)k * 32], x[cp[(int64_t)k * 32]], sum);
  out[r] = MODE == MODE_RESIDUAL ? __dsub_rn(b[r], sum) : sum;
}

int launch_apply(pmg_ctx ctx, int mode, const Sell &s, const double *x, const double *b, double *out)
{
  if (s.nslices == 0) return 0;
  const int  wpb  = 8;
  const dim3 grid((unsigned)((s.nslices + wpb - 1) / wpb)), block(wpb * 32);
  if (mode == MODE_SPMV) sell_apply_kernel<MODE_SPMV><<<grid, block, 0, ctx->stream>>>(s.slice_off.p, s.col.p, s.val.p, s.rows.p, x, b, out, s.nslices);
  else if (mode == MODE_RESIDUAL) sell_apply_kernel<MODE_RESIDUAL><<<grid, block, 0, ctx->stream>>>(s.slice_off.p, s.col.p, s.val.p, s.rows.p, x, b, out, s.nslices);
  else sell_apply_kernel<MODE_ADD><<<grid, block, 0, ctx->stream>>>(s.slice_off.p, s.col.p, s.val.p, s.rows.p, x, b, out, s.nslices);
  PMG_CUDA(cudaGetLastError());
  ctx->launches++;
  return 0;
}

std::vector<int32_t> natural_rows(int64_t n)
{
  std::vector<int32_t> rows((size_t)((n + 31) / 32 * 32), -1);
  for (int64_t r = 0; r < n; ++r) rows[(size_t)r] = (int32_t)r;
  return rows;
}

struct CsrOp final : LevelOp {
  HostCsr              A;
  std::vector<int64_t> diagptr;
  std::vector<int32_t> color;
  int                  ncol = 0;
  std::vector<int64_t> color_slice; // first slice of each colour, ncol+1
  std::vector<int32_t> sweep_rows;  // padded position -> row
  Sell                 sw_sell, full;
  bool                 sweep_ready = false;
  int                  gdim        = 0; // > 0: rows are the nodes of a gdims[0] x gdims[1] x gdims[2] grid, natural order
  int64_t              gdims[3]    = {0, 0, 0};

  bool structured(int &dim, int64_t dims[3]) const override
  {
    if (!gdim) return false;
    dim = gdim;
    std::memcpy(dims, gdims, sizeof gdims);
    return true;
  }

  int64_t        n() const override { return A.n; }
  int            ncolors() const override { return ncol; }
  const HostCsr *host_csr() override { return &A; }

  int init()
  {
    diagptr.assign((size_t)A.n, -1);
    for (int64_t r = 0; r < A.n; ++r) { // MatGetDiagonalPointers, src/mc_sor.c:126-150
      for (int64_t k = A.rowptr[r]; k < A.rowptr[r + 1]; ++k)
        if (A.col[k] == r) diagptr[r] = k;
      if (diagptr[r] < 0) PMG_FAIL(PMG_ERR_ARG, "row %lld has no diagonal entry", (long long)r);
    }
    PMG_TRY(build_sell(ctx, A, natural_rows(A.n), false, full));
    return set_coloring_auto(PMG_COLORING_GREEDY);
  }

  int rebuild_sweep()
  {
    color_slice.assign((size_t)ncol + 1, 0);
    std::vector<int64_t> cnt((size_t)ncol, 0);
    for (int64_t r = 0; r < A.n; ++r) cnt[color[r]]++;
    for (int c = 0; c < ncol; ++c) color_slice[c + 1] = color_slice[c] + (cnt[c] + 31) / 32;
    sweep_rows.assign((size_t)color_slice[ncol] * 32, -1);
    std::vector<int64_t> pos((size_t)ncol);
    for (int c = 0; c < ncol; ++c) pos[c] = color_slice[c] * 32;
    for (int64_t r = 0; r < A.n; ++r) sweep_rows[(size_t)pos[color[r]]++] = (int32_t)r; // ascending rows per colour (ISColoringGetIS)
    PMG_TRY(build_sell(ctx, A, sweep_rows, true, sw_sell));
    sweep_ready = true;
    return 0;
  }

  int get_coloring(std::vector<int32_t> &c) override
  {
    c = color;
    return 0;
  }
  int set_coloring(int nc, const int32_t *c) override
  {
    if (nc < 1) PMG_FAIL(PMG_ERR_ARG, "need at least one colour");
    std::vector<int32_t> cand(c, c + A.n);
    for (int64_t r = 0; r < A.n; ++r)
      if (cand[r] < 0 || cand[r] >= nc) PMG_FAIL(PMG_ERR_ARG, "colour of row %lld out of range", (long long)r);
    const int64_t bad = host_coloring_violations(A, cand);
    if (bad) PMG_FAIL(PMG_ERR_COLORING, "not a distance-1 colouring: %lld adjacent same-colour pairs (a parallel sweep would race)", (long long)bad);
    color = std::move(cand);
    ncol  = nc;
    return rebuild_sweep();
  }
  int set_coloring_auto(int policy) override
  {
    if (policy == PMG_COLORING_GREEDY) ncol = host_coloring_greedy(A, color);
    else if (policy == PMG_COLORING_LEXICOGRAPHIC) ncol = host_coloring_levelset(A, color);
    else if (policy == PMG_COLORING_PARITY) {
      if (!gdim) PMG_FAIL(PMG_ERR_SUP, "parity colouring needs a structured operator");
      int64_t maxlen = 0;
      for (int64_t r = 0; r < A.n; ++r) maxlen = std::max(maxlen, A.rowptr[r + 1] - A.rowptr[r]);
      const bool star = maxlen <= 2 * gdim + 1; // star stencil: red-black; box stencil: 2^d colours
      color.resize((size_t)A.n);
      for (int64_t r = 0; r < A.n; ++r) {
        const int64_t i = r % gdims[0], j = (r / gdims[0]) % gdims[1], k = r / (gdims[0] * gdims[1]);
        color[r] = star ? (int32_t)((i + j + k) & 1) : (int32_t)((i & 1) + 2 * (j & 1) + (gdim == 3 ? 4 * (k & 1) : 0));
      }
      ncol = star ? 2 : (gdim == 3 ? 8 : 4);
      if (host_coloring_violations(A, color)) PMG_FAIL(PMG_ERR_COLORING, "parity colouring is not valid for this operator");
    } else PMG_FAIL(PMG_ERR_ARG, "unknown colouring policy %d", policy);
    return rebuild_sweep();
  }

  int make_coeffs(double omega, SweepCoeffs &c) override
  {
    if (!(omega > 0 && omega < 2)) PMG_FAIL(PMG_ERR_ARG, "omega must be in (0,2), got %g", omega); // PetscOptionsRangeReal, src/pc_mcgibbs.c:197
    const size_t        np = sweep_rows.size();
    std::vector<double> idiag(np, 0.0), sq(np, 0.0);
    const double        f = std::sqrt((2 - omega) / omega);
    for (size_t p = 0; p < np; ++p) {
      const int32_t r = sweep_rows[p];
      if (r < 0) continue;
      const double d = A.val[(size_t)diagptr[r]];
      double       i = 1.0 / d; // VecReciprocal; VecScale (src/mc_sor.c:119-121)
      idiag[p]       = i * omega;
      sq[p]          = std::sqrt(std::fabs(d)) * f; // VecSqrtAbs; VecScale (src/pc_mcgibbs.c:148-150)
    }
    PMG_TRY(c.idiag.upload(idiag, ctx->stream));
    PMG_TRY(c.sqrtdiag.upload(sq, ctx->stream));
    PMG_CUDA(cudaStreamSynchronize(ctx->stream));
    c.omega = omega;
    return 0;
  }

  int sweep_colour(int c, const SweepCoeffs &co, const double *b, double *y, const NoiseArgs &na)
  {
    const int64_t s0 = color_slice[c], ns = color_slice[c + 1] - s0;
    if (ns == 0) return 0;
    const int  wpb = 8;
    const dim3 grid((unsigned)((ns + wpb - 1) / wpb)), block(wpb * 32);
    sell_sweep_kernel<<<grid, block, 0, ctx->stream>>>(sw_sell.slice_off.p, sw_sell.col.p, sw_sell.val.p, sw_sell.rows.p, co.idiag.p, co.sqrtdiag.p, b, y, 1.0 - co.omega, na, s0, ns);
    PMG_CUDA(cudaGetLastError());
    ctx->launches++;
    return 0;
  }

  int sweep(int dir, const SweepCoeffs &co, const double *b, double *y, const NoiseArgs &na) override
  {
    if (!sweep_ready) PMG_FAIL(PMG_ERR_ORDER, "operator has no colouring");
    if (dir == PMG_SOR_FORWARD_SWEEP)
      for (int c = 0; c < ncol; ++c) PMG_TRY(sweep_colour(c, co, b, y, na));
    else
      for (int c = ncol - 1; c >= 0; --c) PMG_TRY(sweep_colour(c, co, b, y, na));
    ctx->dof_updates += A.n;
    return 0;
  }
  int residual(const double *b, const double *x, double *r) override { return launch_apply(ctx, MODE_RESIDUAL, full, x, b, r); }
  int mult(const double *x, double *y) override { return launch_apply(ctx, MODE_SPMV, full, x, nullptr, y); }
  void describe(std::string &out) override
  {
    char buf[256];
    snprintf(buf, sizeof buf, "CSR operator (SELL-32 on device): %lld rows, %lld nonzeros, %d colours", (long long)A.n, (long long)A.nnz(), ncol);
    out = buf;
  }
};

struct CsrTransfer final : Transfer {
  pmg_ctx ctx;
  Sell    P, R;
  int restrict_to(const double *r_fine, double *b_coarse) override { return launch_apply(ctx, MODE_SPMV, R, r_fine, nullptr, b_coarse); }
  int prolong_add(const double *x_coarse, double *x_fine) override { return launch_apply(ctx, MODE_ADD, P, x_coarse, nullptr, x_fine); }
};

} // namespace

int make_csr_op(pmg_ctx ctx, HostCsr &&a, std::unique_ptr<LevelOp> &op)
{
  if (a.n <= 0 || a.n >= INT32_MAX) PMG_FAIL(PMG_ERR_ARG, "unsupported row count %lld", (long long)a.n);
  for (int64_t r = 0; r < a.n; ++r)
    for (int64_t k = a.rowptr[r]; k < a.rowptr[r + 1]; ++k) {
      if (a.col[k] < 0 || a.col[k] >= a.m) PMG_FAIL(PMG_ERR_ARG, "column index out of range in row %lld", (long long)r);
      if (k > a.rowptr[r] && a.col[k] <= a.col[k - 1]) PMG_FAIL(PMG_ERR_ARG, "row %lld: columns must be strictly ascending", (long long)r);
    }
  auto o = std::make_unique<CsrOp>();
  o->ctx = ctx;
  o->A   = std::move(a);
  PMG_TRY(o->init());
  op = std::move(o);
  return 0;
}

int make_csr_grid_op(pmg_ctx ctx, HostCsr &&a, int dim, const int64_t dims[3], std::unique_ptr<LevelOp> &op)
{
  PMG_TRY(make_csr_op(ctx, std::move(a), op));
  auto *o = static_cast<CsrOp *>(op.get());
  o->gdim = dim;
  std::memcpy(o->gdims, dims, sizeof o->gdims);
  return 0;
}

int make_csr_transfer(pmg_ctx ctx, const HostCsr &p, std::unique_ptr<Transfer> &t)
{
  auto    x = std::make_unique<CsrTransfer>();
  HostCsr r;
  x->ctx = ctx;
  host_transpose(p, r); // R = P^T; a coarse row lists its fine rows ascending = MatMultTranspose's accumulation order
  PMG_TRY(build_sell(ctx, p, natural_rows(p.n), false, x->P));
  PMG_TRY(build_sell(ctx, r, natural_rows(r.n), false, x->R));
  t = std::move(x);
  return 0;
}

// ---- small vector kernels ------------------------------------------------------------------------
__global__ void normal_fill_kernel(NoiseArgs na, int64_t n, double *z)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) z[i] = noise_value(na, i);
}
int launch_normal_fill(pmg_ctx ctx, const NoiseArgs &na, int64_t n, double *z_dev)
{
  if (n == 0) return 0;
  normal_fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(na, n, z_dev);
  PMG_CUDA(cudaGetLastError());
  ctx->launches++;
  return 0;
}

__global__ void axpy_kernel(int64_t n, double a, const double *__restrict__ x, double *__restrict__ y)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = __dadd_rn(y[i], __dmul_rn(a, x[i]));
}
int launch_axpy(pmg_ctx ctx, int64_t n, double a, const double *x, double *y)
{
  if (n == 0) return 0;
  axpy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, a, x, y);
  PMG_CUDA(cudaGetLastError());
  ctx->launches++;
  return 0;
}
