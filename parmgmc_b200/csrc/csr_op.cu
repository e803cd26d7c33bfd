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
// Row-partitioned operators (MCSORApply_MPIAIJ, src/mc_sor.c:298-381): column indices >= nl address the ghost values
// gathered from the other ranks (`ghost`, one slot per distinct off-rank column); nl = INT32_MAX on one device.
__device__ __forceinline__ double ext_load(const double *__restrict__ y, const double *__restrict__ ghost, int32_t nl, int32_t c) { return c < nl ? y[c] : ghost[c - nl]; }

__global__ void __launch_bounds__(256) sell_sweep_kernel(const int64_t *__restrict__ slice_off, const int32_t *__restrict__ col, const double *__restrict__ val, const int32_t *__restrict__ rows, const double *__restrict__ idiag, const double *__restrict__ sqrtdiag, const double *__restrict__ b, double *__restrict__ y, double one_minus_omega, NoiseArgs na, int64_t slice0, int64_t nslices, int32_t nl, const double *__restrict__ ghost)
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
  for (int k = 0; k < w; ++k) sum = fma(-vp[(int64_t)k * 32], ext_load(y, ghost, nl, cp[(int64_t)k * 32]), sum);
  const double t = __dmul_rn(one_minus_omega, y[r]);
  y[r]           = fma(idiag[p], sum, t);
}

// pack the locally owned values the other ranks need (send list grouped by destination rank)
__global__ void gather_kernel(int64_t n, const int32_t *__restrict__ idx, const double *__restrict__ y, double *__restrict__ out)
{
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n) out[q] = y[idx[q]];
}

// unpack a per-colour receive buffer into the ghost array
__global__ void scatter_kernel(int64_t n, const int32_t *__restrict__ idx, const double *__restrict__ in, double *__restrict__ ghost)
{
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n) ghost[idx[q]] = in[q];
}

enum ApplyMode { MODE_SPMV = 0, MODE_RESIDUAL = 1, MODE_ADD = 2 };

// out_r = sum_k a_k x[c_k]          (SPMV: MatMult / MatMultTranspose with the transposed matrix)
// out_r = b_r - sum_k a_k x[c_k]    (RESIDUAL: MatMult then VecAYPX(-1, b))
// out_r = out_r + sum ... started from out_r (ADD: MatMultAdd, MatInterpolateAdd)
template <int MODE> __global__ void __launch_bounds__(256) sell_apply_kernel(const int64_t *__restrict__ slice_off, const int32_t *__restrict__ col, const double *__restrict__ val, const int32_t *__restrict__ rows, const double *__restrict__ x, const double *__restrict__ b, double *__restrict__ out, int64_t nslices, int32_t nl, const double *__restrict__ ghost)
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
  for (int k = 0; k < w; ++k) sum = fma(vp[(int64_t)k * 32], ext_load(x, ghost, nl, cp[(int64_t)k * 32]), sum);
  out[r] = MODE == MODE_RESIDUAL ? __dsub_rn(b[r], sum) : sum;
}

int launch_apply(pmg_ctx ctx, int mode, const Sell &s, const double *x, const double *b, double *out, int32_t nl = INT32_MAX, const double *ghost = nullptr)
{
  if (s.nslices == 0) return 0;
  const int  wpb  = 8;
  const dim3 grid((unsigned)((s.nslices + wpb - 1) / wpb)), block(wpb * 32);
  if (mode == MODE_SPMV) sell_apply_kernel<MODE_SPMV><<<grid, block, 0, ctx->stream>>>(s.slice_off.p, s.col.p, s.val.p, s.rows.p, x, b, out, s.nslices, nl, ghost);
  else if (mode == MODE_RESIDUAL) sell_apply_kernel<MODE_RESIDUAL><<<grid, block, 0, ctx->stream>>>(s.slice_off.p, s.col.p, s.val.p, s.rows.p, x, b, out, s.nslices, nl, ghost);
  else sell_apply_kernel<MODE_ADD><<<grid, block, 0, ctx->stream>>>(s.slice_off.p, s.col.p, s.val.p, s.rows.p, x, b, out, s.nslices, nl, ghost);
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

  // ---- row-partitioned operator (MatMPIAIJGetSeqAIJ's diagonal block, off-diagonal block and column map,
  //      src/mc_sor.c:308-310): A holds this rank's rows with LOCAL column indices, columns >= n_local are ghosts ----
  bool                 dist = false;
  int64_t              row_start = 0, n_global = 0, n_local = 0, n_ghost = 0;
  HostCsr              Aloc; // the diagonal block alone (local graph: colouring)
  std::vector<int64_t> ghost_gid, send_off, recv_off; // sorted global ids of the ghost columns; per-rank offsets
  DevBuf<int32_t>      send_idx;
  DevBuf<double>       send_buf, ghost;
  int32_t              nl() const { return dist ? (int32_t)n_local : INT32_MAX; }
  int64_t              nglobal() const override { return dist ? n_global : A.n; }
  int64_t              row0() const override { return row_start; }
  // gather the current values of the ghost columns (the reference does this per colour with one VecScatter per colour
  // holding only that colour's columns, src/mc_sor.c:152-214; one plan for all ghost columns moves a superset)
  int halo(const double *y)
  {
    if (!dist || ctx->nranks == 1) return 0;
    const int64_t ns = send_off.back();
    if (ns) {
      gather_kernel<<<(unsigned)((ns + 255) / 256), 256, 0, ctx->stream>>>(ns, send_idx.p, y, send_buf.p);
      PMG_CUDA(cudaGetLastError());
      ctx->launches++;
    }
    return comm_exchange_v(ctx, send_buf.p, send_off.data(), ghost.p, recv_off.data(), ctx->stream);
  }
  // ---- per-colour ghost plans: the reference builds one VecScatter per colour holding only the ghost columns that rows of
  //      that colour reference (src/mc_sor.c:152-214, gathered before the colour at :318-319).  Same here: colour c moves
  //      only the values its rows read. ----
  struct ColourPlan {
    std::vector<int64_t> send_off, recv_off; // per-rank offsets into this colour's packed buffers
    int64_t              send_base = 0, recv_base = 0; // offsets of this colour in csend_idx / crecv_idx
  };
  std::vector<ColourPlan> cplan;
  DevBuf<int32_t>         csend_idx, crecv_idx; // all colours back to back
  DevBuf<double>          crecv_buf;
  bool                    cplan_ok = false;
  int build_colour_plans()
  {
    cplan_ok = false;
    cplan.clear();
    if (!dist || ctx->nranks == 1) return 0;
    const int P = ctx->nranks, me = ctx->rank;
    // (colour, ghost column) pairs: colour c of MY rows reads ghost column q; sorted by colour, then by ghost index (= owner, global id)
    std::vector<std::pair<int32_t, int32_t>> pairs;
    for (int64_t r = 0; r < n_local; ++r)
      for (int64_t k = A.rowptr[r]; k < A.rowptr[r + 1]; ++k)
        if (A.col[k] >= n_local) pairs.emplace_back(color[(size_t)r], (int32_t)(A.col[k] - n_local));
    std::sort(pairs.begin(), pairs.end());
    pairs.erase(std::unique(pairs.begin(), pairs.end()), pairs.end());
    // everybody learns everybody's pairs as (colour, GLOBAL id) (set-up only)
    std::vector<int64_t> cnts((size_t)P), mine1{(int64_t)pairs.size()};
    PMG_TRY(comm_allgather_i64(ctx, mine1.data(), 1, cnts.data()));
    int64_t mx = 1;
    for (int64_t v : cnts) mx = std::max(mx, v);
    std::vector<int64_t> pad((size_t)2 * mx, -1), all((size_t)2 * mx * P);
    for (size_t q = 0; q < pairs.size(); ++q) {
      pad[2 * q]     = pairs[q].first;
      pad[2 * q + 1] = ghost_gid[(size_t)pairs[q].second];
    }
    PMG_TRY(comm_allgather_i64(ctx, pad.data(), (int)(2 * mx), all.data()));
    // per rank: first pair of each colour (the lists are sorted by colour)
    auto colour_range = [&](int r, int c, int64_t &lo, int64_t &hi) {
      const int64_t *base = &all[(size_t)2 * mx * r];
      int64_t        a0 = 0, a1 = cnts[(size_t)r];
      while (a0 < a1) { const int64_t m = (a0 + a1) / 2; if (base[2 * m] < c) a0 = m + 1; else a1 = m; }
      lo = a0;
      a1 = cnts[(size_t)r];
      while (a0 < a1) { const int64_t m = (a0 + a1) / 2; if (base[2 * m] <= c) a0 = m + 1; else a1 = m; }
      hi = a0;
    };
    std::vector<int64_t> starts((size_t)P + 1, 0); // row ownership, from the all-ghost plan's set-up
    {
      std::vector<int64_t> me2{row_start, n_local}, a2((size_t)2 * P);
      PMG_TRY(comm_allgather_i64(ctx, me2.data(), 2, a2.data()));
      for (int r = 0; r < P; ++r) starts[(size_t)r] = a2[(size_t)2 * r];
      starts[(size_t)P] = a2[(size_t)2 * (P - 1)] + a2[(size_t)2 * (P - 1) + 1];
    }
    std::vector<int32_t> sidx, ridx;
    cplan.resize((size_t)ncol);
    int64_t max_recv = 1, max_send = 1;
    for (int c = 0; c < ncol; ++c) {
      ColourPlan &pl = cplan[(size_t)c];
      pl.send_base = (int64_t)sidx.size();
      pl.recv_base = (int64_t)ridx.size();
      pl.send_off.assign((size_t)P + 1, 0);
      pl.recv_off.assign((size_t)P + 1, 0);
      int64_t mlo, mhi;
      colour_range(me, c, mlo, mhi);
      int64_t mq = mlo; // my pairs of this colour are sorted by ghost index, i.e. grouped by owner in rank order
      for (int r = 0; r < P; ++r) {
        if (r != me) {
          int64_t lo, hi;
          colour_range(r, c, lo, hi);
          for (int64_t q = lo; q < hi; ++q) { // rank r's colour-c ghosts that I own, in r's ghost order
            const int64_t g = all[(size_t)2 * mx * r + 2 * (size_t)q + 1];
            if (g >= row_start && g < row_start + n_local) sidx.push_back((int32_t)(g - row_start));
          }
        }
        for (; mq < mhi && all[(size_t)2 * mx * me + 2 * (size_t)mq + 1] < starts[(size_t)r + 1]; ++mq) // my colour-c ghosts owned by r
          ridx.push_back(pairs[(size_t)mq].second);
        pl.send_off[(size_t)r + 1] = (int64_t)sidx.size() - pl.send_base;
        pl.recv_off[(size_t)r + 1] = (int64_t)ridx.size() - pl.recv_base;
      }
      max_recv = std::max(max_recv, pl.recv_off[(size_t)P]);
      max_send = std::max(max_send, pl.send_off[(size_t)P]);
    }
    (void)max_send; // a colour sends a subset of the all-ghost send list: send_buf is large enough
    if (sidx.empty()) sidx.push_back(0);
    if (ridx.empty()) ridx.push_back(0);
    PMG_TRY(csend_idx.upload(sidx, ctx->stream));
    PMG_TRY(crecv_idx.upload(ridx, ctx->stream));
    PMG_TRY(crecv_buf.alloc((size_t)max_recv));
    PMG_CUDA(cudaStreamSynchronize(ctx->stream));
    cplan_ok = true;
    return 0;
  }
  // bytes one sweep moves to this rank: per-colour plans against the all-ghost plan before every colour
  void halo_volume(int64_t &per_colour, int64_t &all_ghosts) const
  {
    per_colour = 0;
    for (const ColourPlan &pl : cplan) per_colour += pl.recv_off.back();
    all_ghosts = (int64_t)ncol * n_ghost;
  }
  int halo_colour(int c, const double *y)
  {
    if (!dist || ctx->nranks == 1) return 0;
    if (!cplan_ok || std::getenv("PMG_CSR_HALO_ALL")) return halo(y);
    const ColourPlan &pl = cplan[(size_t)c];
    const int64_t     ns = pl.send_off.back(), nr = pl.recv_off.back();
    if (ns) {
      gather_kernel<<<(unsigned)((ns + 255) / 256), 256, 0, ctx->stream>>>(ns, csend_idx.p + pl.send_base, y, send_buf.p);
      PMG_CUDA(cudaGetLastError());
      ctx->launches++;
    }
    PMG_TRY(comm_exchange_v(ctx, send_buf.p, pl.send_off.data(), crecv_buf.p, pl.recv_off.data(), ctx->stream));
    if (nr) {
      scatter_kernel<<<(unsigned)((nr + 255) / 256), 256, 0, ctx->stream>>>(nr, crecv_idx.p + pl.recv_base, crecv_buf.p, ghost.p);
      PMG_CUDA(cudaGetLastError());
      ctx->launches++;
    }
    return 0;
  }
  // colours of the ghost columns (for validating / building a GLOBAL distance-1 colouring)
  int ghost_colours(const std::vector<int32_t> &mine, std::vector<int32_t> &gc)
  {
    gc.assign((size_t)n_ghost, -1);
    if (!dist || ctx->nranks == 1) return 0;
    std::vector<double> h((size_t)n_local);
    for (int64_t r = 0; r < n_local; ++r) h[(size_t)r] = (double)mine[(size_t)r];
    DevBuf<double> d;
    PMG_TRY(d.upload(h, ctx->stream));
    PMG_TRY(halo(d.p));
    std::vector<double> g((size_t)n_ghost);
    if (n_ghost) PMG_CUDA(cudaMemcpyAsync(g.data(), ghost.p, sizeof(double) * (size_t)n_ghost, cudaMemcpyDeviceToHost, ctx->stream));
    PMG_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int64_t q = 0; q < n_ghost; ++q) gc[(size_t)q] = (int32_t)g[(size_t)q];
    return 0;
  }
  // adjacent same-colour pairs, local and across ranks, summed over all ranks (every rank gets the same number)
  int global_violations(const std::vector<int32_t> &cand, int64_t &bad)
  {
    bad = host_coloring_violations(Aloc, cand);
    std::vector<int32_t> gc;
    PMG_TRY(ghost_colours(cand, gc));
    for (int64_t r = 0; r < n_local; ++r)
      for (int64_t k = A.rowptr[r]; k < A.rowptr[r + 1]; ++k)
        if (A.col[k] >= n_local && gc[(size_t)(A.col[k] - n_local)] == cand[(size_t)r]) ++bad;
    std::vector<int64_t> all((size_t)ctx->nranks);
    PMG_TRY(comm_allgather_i64(ctx, &bad, 1, all.data()));
    bad = std::accumulate(all.begin(), all.end(), (int64_t)0);
    return 0;
  }

  bool structured(int &dim, int64_t dims[3]) const override
  {
    if (!gdim) return false;
    dim = gdim;
    std::memcpy(dims, gdims, sizeof gdims);
    return true;
  }

  int64_t        n() const override { return A.n; }
  int            ncolors() const override { return ncol; }
  const HostCsr *host_csr() override { return dist ? nullptr : &A; } // a slab of rows is not an operator on its own

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

  // set-up of the ghost plan: who owns my ghost columns, and which of my rows the others need
  int init_dist(const std::vector<int64_t> &gids)
  {
    const int P = ctx->nranks, me = ctx->rank;
    ghost_gid = gids;
    n_ghost   = (int64_t)gids.size();
    std::vector<int64_t> mine{row_start, n_local, n_ghost}, all((size_t)3 * P);
    PMG_TRY(comm_allgather_i64(ctx, mine.data(), 3, all.data()));
    std::vector<int64_t> starts((size_t)P + 1);
    int64_t              max_ghost = 0;
    for (int r = 0; r < P; ++r) {
      starts[(size_t)r] = all[(size_t)3 * r];
      max_ghost         = std::max(max_ghost, all[(size_t)3 * r + 2]);
      if (r > 0 && all[(size_t)3 * r] != all[(size_t)3 * (r - 1)] + all[(size_t)3 * (r - 1) + 1]) PMG_FAIL(PMG_ERR_ARG, "row-partitioned CSR: the row ranges of ranks %d and %d are not contiguous", r - 1, r);
    }
    starts[(size_t)P] = all[(size_t)3 * (P - 1)] + all[(size_t)3 * (P - 1) + 1];
    if (starts[0] != 0 || starts[(size_t)P] != n_global) PMG_FAIL(PMG_ERR_ARG, "row-partitioned CSR: the row ranges do not tile 0 .. %lld", (long long)n_global);
    // the ghost ids are sorted, the ownership ranges ascending: the ghost array is grouped by owner in rank order
    recv_off.assign((size_t)P + 1, 0);
    for (int64_t g : ghost_gid) {
      if (g < 0 || g >= n_global || (g >= row_start && g < row_start + n_local)) PMG_FAIL(PMG_ERR_ARG, "row-partitioned CSR: bad ghost column %lld", (long long)g);
      const int owner = (int)(std::upper_bound(starts.begin(), starts.end(), g) - starts.begin()) - 1;
      recv_off[(size_t)owner + 1]++;
    }
    for (int r = 0; r < P; ++r) recv_off[(size_t)r + 1] += recv_off[(size_t)r];
    // everybody learns everybody's ghost list (set-up only), and keeps the ids it owns as its send list for that rank
    std::vector<int64_t> padded((size_t)std::max<int64_t>(1, max_ghost), -1), lists((size_t)std::max<int64_t>(1, max_ghost) * P);
    std::copy(ghost_gid.begin(), ghost_gid.end(), padded.begin());
    PMG_TRY(comm_allgather_i64(ctx, padded.data(), (int)padded.size(), lists.data()));
    std::vector<int32_t> sidx;
    send_off.assign((size_t)P + 1, 0);
    for (int r = 0; r < P; ++r) {
      if (r != me)
        for (int64_t q = 0; q < all[(size_t)3 * r + 2]; ++q) {
          const int64_t g = lists[(size_t)r * padded.size() + (size_t)q];
          if (g >= row_start && g < row_start + n_local) sidx.push_back((int32_t)(g - row_start));
        }
      send_off[(size_t)r + 1] = (int64_t)sidx.size();
    }
    if (sidx.empty()) sidx.push_back(0);
    PMG_TRY(send_idx.upload(sidx, ctx->stream));
    PMG_TRY(send_buf.alloc(sidx.size()));
    PMG_TRY(ghost.alloc((size_t)std::max<int64_t>(1, n_ghost)));
    PMG_TRY(ghost.zero(ctx->stream));
    PMG_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
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
    PMG_TRY(build_colour_plans()); // collective on a row-partitioned operator (set_coloring* are collective there)
    sweep_ready    = true;
    layout_version = pmg_next_layout_version(); // cached per-row coefficients of the old layout are stale now
    return 0;
  }

  int get_coloring(std::vector<int32_t> &c) override
  {
    c = color;
    return 0;
  }
  int set_coloring(int nc, const int32_t *c) override
  {
    // every rank must take the same decision before the first collective (a rank that fails alone would leave the others
    // blocked in NCCL), and the sweep issues one collective halo per colour, so the colour count must be the same everywhere
    int64_t local_ok = nc >= 1 ? 1 : 0, first_bad = -1;
    std::vector<int32_t> cand(c, c + A.n);
    for (int64_t r = 0; r < A.n && local_ok; ++r)
      if (cand[r] < 0 || cand[r] >= nc) { local_ok = 0; first_bad = r; }
    if (dist) {
      std::vector<int64_t> mine{local_ok, (int64_t)nc}, all((size_t)2 * ctx->nranks);
      PMG_TRY(comm_allgather_i64(ctx, mine.data(), 2, all.data()));
      for (int r = 0; r < ctx->nranks; ++r) {
        if (!all[(size_t)2 * r]) PMG_FAIL(PMG_ERR_ARG, "rank %d passed an invalid colouring (colour count < 1 or a colour out of range)", r);
        if (all[(size_t)2 * r + 1] != nc) PMG_FAIL(PMG_ERR_ARG, "the colour count must be the GLOBAL one on every rank: rank %d passed %lld, this rank %d", r, (long long)all[(size_t)2 * r + 1], nc);
      }
    } else if (!local_ok) {
      if (nc < 1) PMG_FAIL(PMG_ERR_ARG, "need at least one colour");
      PMG_FAIL(PMG_ERR_ARG, "colour of row %lld out of range", (long long)first_bad);
    }
    int64_t bad = 0;
    if (dist) PMG_TRY(global_violations(cand, bad)); // collective: the colouring must be valid ACROSS ranks (src/mc_sor.c:383-395)
    else bad = host_coloring_violations(A, cand);
    if (bad) PMG_FAIL(PMG_ERR_COLORING, "not a distance-1 colouring: %lld adjacent same-colour pairs (a parallel sweep would race)", (long long)bad);
    color = std::move(cand);
    ncol  = nc;
    return rebuild_sweep();
  }
  // A global distance-1 colouring without PETSc's Jones-Plassmann (which is PETSc-internal and unpinned):
  // greedy on the local graph with K = the largest local colour count, then colour + K (rank mod 2) when every rank's
  // ghost owners have the other parity (contiguous row blocks of banded matrices), else colour + K rank.
  // PMG_COLORING_LEXICOGRAPHIC on a row-partitioned operator: the level sets of the GLOBAL natural order, level(r) = 1 + max level
  // of the coupled rows with a smaller global index.  Sweeping the levels in ascending order IS the lexicographic Gauss-Seidel
  // sweep over all ranks' rows (what the reference's PCPARSOR computes with its pipelined exact parallel SOR,
  // src/pc_parsor.c:703-878, and its 1-rank MatSOR path), so a multi-GPU run reproduces the one-rank natural-order sampler.
  // Rows of rank p only depend on ranks < p through their ghost columns: P rounds, in round p rank p finishes its levels.
  int set_coloring_lexicographic_dist()
  {
    const int            P = ctx->nranks, me = ctx->rank;
    std::vector<int32_t> lev((size_t)n_local, 0), gl;
    for (int p = 0; p < P; ++p) {
      PMG_TRY(ghost_colours(lev, gl)); // collective: current levels of my ghost columns (final for the ranks below p)
      if (p != me) continue;
      for (int64_t r = 0; r < n_local; ++r) {
        int32_t L = 0;
        for (int64_t k = A.rowptr[r]; k < A.rowptr[r + 1]; ++k) {
          const int64_t c = A.col[k];
          if (c < n_local) {
            if (c < r) L = std::max(L, lev[(size_t)c] + 1);
          } else if (ghost_gid[(size_t)(c - n_local)] < row_start + r) L = std::max(L, gl[(size_t)(c - n_local)] + 1);
        }
        lev[(size_t)r] = L;
      }
    }
    int64_t mymax = 0;
    for (int32_t v : lev) mymax = std::max<int64_t>(mymax, v);
    std::vector<int64_t> all((size_t)P);
    PMG_TRY(comm_allgather_i64(ctx, &mymax, 1, all.data()));
    for (int64_t v : all) mymax = std::max(mymax, v);
    return set_coloring((int)(mymax + 1), lev.data());
  }
  int set_coloring_auto_dist(int policy)
  {
    if (policy == PMG_COLORING_LEXICOGRAPHIC) return set_coloring_lexicographic_dist();
    if (policy != PMG_COLORING_GREEDY) PMG_FAIL(PMG_ERR_SUP, "row-partitioned operators: greedy or lexicographic (global level-set) colouring");
    std::vector<int32_t> loc;
    const int            P = ctx->nranks, me = ctx->rank;
    int64_t              kloc = host_coloring_greedy(Aloc, loc), parity_ok = 1;
    for (int r = 0; r < P; ++r)
      if (recv_off[(size_t)r + 1] > recv_off[(size_t)r] && ((r ^ me) & 1) == 0) parity_ok = 0;
    std::vector<int64_t> mine{kloc, parity_ok}, all((size_t)2 * P);
    PMG_TRY(comm_allgather_i64(ctx, mine.data(), 2, all.data()));
    int64_t K = 1;
    bool    par = true;
    for (int r = 0; r < P; ++r) {
      K   = std::max(K, all[(size_t)2 * r]);
      par = par && all[(size_t)2 * r + 1] != 0;
    }
    const int64_t shift = par ? K * (me & 1) : K * me;
    for (auto &c : loc) c = (int32_t)(c + shift);
    return set_coloring((int)(par ? 2 * K : K * P), loc.data());
  }
  int set_coloring_auto(int policy) override
  {
    if (dist) return set_coloring_auto_dist(policy);
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
    // the greedy / level-set policies look at the entries of row r only: on a structurally non-symmetric pattern two coupled rows
    // could share a colour and the sweep kernel would race
    if (const int64_t bad = host_coloring_violations(A, color)) PMG_FAIL(PMG_ERR_COLORING, "automatic colouring is not a distance-1 colouring of this (structurally non-symmetric?) operator: %lld violations", (long long)bad);
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
    PMG_TRY(halo_colour(c, y)); // collective: every rank takes part for every colour, also for colours it has no rows of
    const int64_t s0 = color_slice[c], ns = color_slice[c + 1] - s0;
    if (ns == 0) return 0;
    const int  wpb = 8;
    const dim3 grid((unsigned)((ns + wpb - 1) / wpb)), block(wpb * 32);
    sell_sweep_kernel<<<grid, block, 0, ctx->stream>>>(sw_sell.slice_off.p, sw_sell.col.p, sw_sell.val.p, sw_sell.rows.p, co.idiag.p, co.sqrtdiag.p, b, y, 1.0 - co.omega, na, s0, ns, nl(), ghost.p);
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
  int residual(const double *b, const double *x, double *r) override
  {
    PMG_TRY(halo(x));
    return launch_apply(ctx, MODE_RESIDUAL, full, x, b, r, nl(), ghost.p);
  }
  int mult(const double *x, double *y) override
  {
    PMG_TRY(halo(x));
    return launch_apply(ctx, MODE_SPMV, full, x, nullptr, y, nl(), ghost.p);
  }
  void describe(std::string &out) override
  {
    char buf[256];
    snprintf(buf, sizeof buf, "CSR operator (SELL-32 on device): %lld rows, %lld nonzeros, %d colours", (long long)A.n, (long long)A.nnz(), ncol);
    out = buf;
    if (dist) {
      snprintf(buf, sizeof buf, "; rows %lld..%lld of %lld, %lld ghost columns", (long long)row_start, (long long)(row_start + n_local), (long long)n_global, (long long)n_ghost);
      out += buf;
      if (cplan_ok) {
        int64_t pc = 0, ag = 0;
        halo_volume(pc, ag);
        snprintf(buf, sizeof buf, "; per-colour ghost plans: %lld values per sweep (all ghosts before every colour: %lld)", (long long)pc, (long long)ag);
        out += buf;
      }
    }
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

int make_csr_dist_op(pmg_ctx ctx, int64_t n_global, int64_t row_start, int64_t n_local, const int64_t *rowptr, const int64_t *col_global, const double *val, std::unique_ptr<LevelOp> &op)
{
  if (n_local <= 0 || n_global >= INT32_MAX || row_start < 0 || row_start + n_local > n_global || rowptr[0] != 0) PMG_FAIL(PMG_ERR_ARG, "row-partitioned CSR: bad sizes");
  const int64_t nnz = rowptr[n_local];
  // distinct off-rank columns, sorted = the column map of the off-diagonal block (src/mc_sor.c:308-310, `colmap`)
  std::vector<int64_t> gids;
  for (int64_t k = 0; k < nnz; ++k) {
    const int64_t c = col_global[k];
    if (c < 0 || c >= n_global) PMG_FAIL(PMG_ERR_ARG, "row-partitioned CSR: column index %lld out of range", (long long)c);
    if (c < row_start || c >= row_start + n_local) gids.push_back(c);
  }
  std::sort(gids.begin(), gids.end());
  gids.erase(std::unique(gids.begin(), gids.end()), gids.end());
  if (n_local + (int64_t)gids.size() >= INT32_MAX) PMG_FAIL(PMG_ERR_ARG, "row-partitioned CSR: too many local + ghost columns");
  // local numbering: own columns first (ascending), ghost columns after them (ascending global id); per row this is the
  // reference's accumulation order: diagonal block, then off-diagonal block (src/mc_sor.c:323-334)
  HostCsr a, aloc;
  a.n = n_local;
  a.m = n_local + (int64_t)gids.size();
  a.rowptr.assign(rowptr, rowptr + n_local + 1);
  a.col.resize((size_t)nnz);
  a.val.resize((size_t)nnz);
  aloc.n = aloc.m = n_local;
  aloc.rowptr.assign((size_t)n_local + 1, 0);
  std::vector<std::pair<int32_t, double>> row;
  for (int64_t r = 0; r < n_local; ++r) {
    row.clear();
    for (int64_t k = rowptr[r]; k < rowptr[r + 1]; ++k) {
      const int64_t c = col_global[k];
      const int32_t lc = (c >= row_start && c < row_start + n_local) ? (int32_t)(c - row_start) : (int32_t)(n_local + (std::lower_bound(gids.begin(), gids.end(), c) - gids.begin()));
      row.emplace_back(lc, val[k]);
    }
    std::sort(row.begin(), row.end(), [](const std::pair<int32_t, double> &x, const std::pair<int32_t, double> &y) { return x.first < y.first; });
    for (size_t q = 0; q < row.size(); ++q) {
      if (q > 0 && row[q].first == row[q - 1].first) PMG_FAIL(PMG_ERR_ARG, "row-partitioned CSR: duplicate column in local row %lld", (long long)r);
      a.col[(size_t)rowptr[r] + q] = row[q].first;
      a.val[(size_t)rowptr[r] + q] = row[q].second;
      if (row[q].first < n_local) {
        aloc.col.push_back(row[q].first);
        aloc.val.push_back(row[q].second);
      }
    }
    aloc.rowptr[(size_t)r + 1] = (int64_t)aloc.col.size();
  }
  auto o       = std::make_unique<CsrOp>();
  o->ctx       = ctx;
  o->dist      = true;
  o->row_start = row_start;
  o->n_global  = n_global;
  o->n_local   = n_local;
  o->A         = std::move(a);
  o->Aloc      = std::move(aloc);
  PMG_TRY(o->init_dist(gids));
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

// one thread = one generator call = four consecutive padded columns of one grid row of one segment
__global__ void __launch_bounds__(256) noise_prefill_kernel(const __grid_constant__ PrefillArgs a)
{
  pdl_launch_dependents();
  pdl_wait(); // the buffers were read by the sweeps of the sample before
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < a.total; q += (int64_t)gridDim.x * blockDim.x) {
    int s = 0;
    while (s + 1 < a.nseg && q >= a.seg[s + 1].q0) ++s;
    const int64_t ql   = q - a.seg[s].q0; // = (j * pitch + c) >> 2: the quad index of box2d.cuh's noise_row
    const int     qrow = a.seg[s].pitch >> 2, nx = a.seg[s].nx;
    const int     j = (int)(ql / qrow), c = 4 * (int)(ql - (int64_t)j * qrow);
    double        z[4];
    philox_normal_quad(a.seed, a.seg[s].call, (uint64_t)ql, z);
    double *d = a.seg[s].dst + (int64_t)j * nx + c;
#pragma unroll
    for (int m = 0; m < 4; ++m)
      if (c + m < nx) d[m] = z[m];
  }
}
int launch_noise_prefill(pmg_ctx ctx, const PrefillArgs &a)
{
  if (a.total == 0) return 0;
  const unsigned grid = (unsigned)std::min<int64_t>((a.total + 255) / 256, (int64_t)ctx->sm_count * 8);
  PMG_CUDA(launch_pdl(ctx->stream, noise_prefill_kernel, dim3(grid), dim3(256), 0, a));
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
