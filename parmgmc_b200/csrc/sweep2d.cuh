// sweep2d.cuh -- TMA-fed, warp-streaming red-black Gibbs sweep for the matrix-free 2D 5-point operator (K1 of SURVEY 8(d)).
//
// One pass over memory does what the reference does in five (VecSetRandomStandardNormal, VecPointwiseMult, VecAXPY and
// the two colour phases of MCSORApply: src/pc_mcgibbs.c:119-128, src/mc_sor.c:257-271): 8 B of x and 8 B of b read and
// 8 B written per DOF-update.
//
// Work decomposition (as stream2d.cuh): a warp owns a strip of 128 columns (lane l owns columns c0+4l .. c0+4l+3; lanes
// 1..30 are written back, lanes 0 and 31 recompute the neighbouring strips' edge columns) and a band of rows, walks down the
// band and keeps three rows in registers.  At row step jj it updates the first colour of row jj (all neighbours old), then
// the second colour of row jj-1 (all neighbours new) and stores row jj-1 out of place.  East / west neighbours come from
// warp shuffles, north / south neighbours are the thread's own registers; no block barrier in the loop.
//
// What is new here:
//  * Rows arrive through the TMA: lane 0 of each warp issues cp.async.bulk.tensor.2d copies of 128 x 2 boxes of x and b
//    into the warp's private ring of shared-memory stages and every lane waits on the stage's mbarrier.  In-flight data
//    costs no registers, the prefetch distance is a ring depth, and out-of-grid coordinates are zero-filled by the
//    hardware, so grid edges need no load predicates (the pad columns nx .. pitch-1 of the vectors are kept at zero).
//  * The loop is unrolled over the two row parities: no colour selects, no divergent code.
//  * Everything that does not change in the loop is a compile-time choice (noise mode, interior / edge warp).
//  * The generator is keyed on the padded index (philox.cuh), so a thread's four columns are one Philox call.
// Arithmetic per node is exactly that of the per-colour kernel (stencil_op.cu lap_sweep_kernel), fma for fma: edge nodes
// add h * 0 for a missing neighbour, which is exact, so the result is bit-identical (up to the sign of a zero).
#pragma once
#include <cuda.h>

#include "common.hpp"
#include "fastnormal.cuh"
#include "philox.cuh"

namespace sweep2d {

constexpr int STRIP_OUT = 120; // columns written per warp
constexpr int STAGE_ROWS = 2;  // rows per TMA box
constexpr int ROW_BYTES = 128 * 8;
constexpr int STAGE_BYTES = 2 * STAGE_ROWS * ROW_BYTES; // x rows + b rows

enum { NOISE_NONE = 0, NOISE_TAPE = 1, NOISE_PHILOX = 2 };

// one warp's work: output columns of strip `strip`, output rows [ja, jb)
struct Item {
  int strip, ja, jb;
};

struct Coef { // per-node coefficients of the edge warps, indexed by the number of existing neighbours; 5 = no such node
  double idiag, sd, omo, diag;
};

struct Args {
  CUtensorMap   tm_x, tm_b; // {4, pitch/4, local rows} FP64 tensors (SWIZZLE_32B), box 4 x 32 x 2 = 128 columns x 2 rows
  int           nx, ny;     // global grid
  int           slo, shi;   // owned rows: the rows that are written, and the rows the injected tape covers
  int           tlo, thi;   // rows held by the tensors and by xout: the owned rows plus two ghost rows per side on a slab
  const Item   *items;
  int           nitems;
  int           pitch; // row stride of xout (and of the tensors)
  int           flip;  // 0: forward sweep (colour (i+j) even first); 1: backward
  int           has_b;
  int           swizzle; // 1: tensors are {4, pitch/4, rows} with SWIZZLE_32B; 0: {pitch, rows, 1}, plain rows
  double       *xout;
  const double *xc;   // non-null: the sweep starts from xin + P xc (MatInterpolateAdd fused into the post-smoother); xc is the
  int           cnx, cny, cpitch; // coarse level's iterate on its cnx x cny grid, row stride cpitch (Q1)
  int           ctlo;             // first coarse row held by xc / bc (a slab of the coarse level carries ghost rows)
  double       *bc;   // non-null (RESTRICT kernels): b_c = P^T (b - A xout), MatResidual + MatRestrict fused into the pre-smoother
  const double *tape; // injected noise of this block: natural layout (row stride nx), local rows
  double        h, idiag, sd, omo, diag; // interior coefficients
  Coef          coef[6];
  PhiloxKeys    pk;
  uint32_t      call_lo, call_hi;
};

// ---- PTX helpers -------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void     mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void     mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void     mbar_wait(uint32_t bar, uint32_t parity)
{
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// The tensors are described to the TMA as {4, pitch/4, rows...} with SWIZZLE_32B: a row of 128 doubles lands in shared memory
// as 32 lane segments of 32 bytes, with the two 16-byte halves of a segment exchanged in every other 128-byte line.  A
// lane's LDS.128 of "its first half" then hits banks that the lanes 4 further do not (conflict-free row reads).
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *tm, int c0, int c1, int c2, uint32_t bar)
{
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t lane_seg(int lane) { return (uint32_t)(lane * 32); }           // byte offset of the lane's segment in a row
__device__ __forceinline__ uint32_t lane_swz(int lane) { return (uint32_t)(((lane >> 2) & 1) * 16); } // swizzle of that segment
// the lane's four doubles of the row at `row` (shared-memory byte address of the row start, 1024-byte aligned)
__device__ __forceinline__ void lds256(uint32_t row, int lane, double (&v)[4], bool swizzled = true)
{
  const uint32_t a = row + lane_seg(lane), s = swizzled ? lane_swz(lane) : 0u;
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "r"(a + s));
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v[2]), "=d"(v[3]) : "r"(a + (16u - s)));
}
__device__ __forceinline__ void st256(double *p, const double (&v)[4]) { asm volatile("st.global.v4.f64 [%4], {%0,%1,%2,%3};" ::"d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]), "l"(p) : "memory"); }
__device__ __forceinline__ double shfl_up1(double v) { return __shfl_up_sync(0xffffffffu, v, 1); }
__device__ __forceinline__ double shfl_dn1(double v) { return __shfl_down_sync(0xffffffffu, v, 1); }

// ---- one warp ----------------------------------------------------------------------------------------------------
template <int NOISE, bool INTERIOR> struct Warp {
  const Args              &a;
  const fastnormal::Tables ft;
  const Coef              *coef; // shared-memory copy of a.coef (edge warps)
  int                      lane, c, ja, jb;
  bool                     out_lane;
  // edge warps: column part of the node classification
  int  colmiss[4];
  bool colok[4];

  __device__ __forceinline__ Warp(const Args &a_, const fastnormal::Tables &ft_, const Coef *coef_, int lane_, int c_, int ja_, int jb_) : a(a_), ft(ft_), coef(coef_), lane(lane_), c(c_), ja(ja_), jb(jb_)
  {
    out_lane = lane >= 1 && lane <= 30 && (INTERIOR || c < a.nx);
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      colok[m]   = c + m >= 0 && c + m < a.nx;
      colmiss[m] = (c + m == 0 ? 1 : 0) + (c + m == a.nx - 1 ? 1 : 0);
    }
  }

  // w = b + sqrtdiag z of the four nodes of row jj (src/pc_mcgibbs.c:124-126: two roundings)
  __device__ __forceinline__ void noisy_rhs4(int jj, const double (&b)[4], const int (&ci)[4], double (&w)[4])
  {
    if (NOISE == NOISE_NONE) {
#pragma unroll
      for (int m = 0; m < 4; ++m) w[m] = b[m];
      return;
    }
    double z[4];
    if (NOISE == NOISE_TAPE) {
      const bool rowok = INTERIOR || (jj >= a.slo && jj < a.shi);
      const double *p  = a.tape + (long long)(jj - a.slo) * a.nx + c;
#pragma unroll
      for (int m = 0; m < 4; ++m) z[m] = (rowok && (INTERIOR || colok[m])) ? p[m] : 0.0;
    } else {
      const long long quad = ((long long)jj * a.pitch + c) >> 2;
      uint32_t        w0, w1, w2, w3;
      philox4x32_10_keys((uint32_t)quad, (uint32_t)((unsigned long long)quad >> 32), a.call_lo, a.call_hi, a.pk, w0, w1, w2, w3);
      fastnormal::box_muller(ft, w0, w1, z[0], z[1]);
      fastnormal::box_muller(ft, w2, w3, z[2], z[3]);
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const double sd = INTERIOR ? a.sd : coef[ci[m]].sd;
      w[m]            = __dadd_rn(__dmul_rn(z[m], sd), b[m]);
    }
  }

  // src/mc_sor.c:260-268 for column M of `row`: accumulation order of the assembled row (south, west, east, north)
  template <int M> __device__ __forceinline__ void update(double (&row)[4], const double (&south)[4], const double (&north)[4], double west, double east, double w, int ci)
  {
    const double xw = M == 0 ? west : row[M == 0 ? 0 : M - 1];
    const double xe = M == 3 ? east : row[M == 3 ? 3 : M + 1];
    double       sum = w;
    sum = fma(a.h, south[M], sum);
    sum = fma(a.h, xw, sum);
    sum = fma(a.h, xe, sum);
    sum = fma(a.h, north[M], sum);
    if (INTERIOR) {
      const double t0 = __dmul_rn(a.omo, row[M]);
      row[M]          = fma(a.idiag, sum, t0);
    } else {
      const Coef   k  = coef[ci];
      const double t0 = __dmul_rn(k.omo, row[M]);
      row[M]          = fma(k.idiag, sum, t0);
    }
  }

  // Row step jj for rows of parity P = (jj + flip) & 1: first-colour columns of row jj (and second-colour columns of row
  // jj-1) are M = P, P+2.  xss / xs / x0 = rows jj-2 / jj-1 / jj, xn / bb = row jj+1 of x (old) and row jj of b; wk carries
  // the noisy right-hand side of the two nodes of row jj-1 that phase B updates, and returns that of row jj.
  template <int P> __device__ __forceinline__ void step(int jj, const double (&xss)[4], double (&xs)[4], double (&x0)[4], const double (&xn)[4], const double (&bb)[4], double (&wk)[2], const int (&cis)[4])
  {
    int ci[4] = {0, 0, 0, 0}; // coefficient classes of row jj (edge warps)
    if (!INTERIOR) {
      const bool rowok   = jj >= 0 && jj < a.ny;
      const int  rowmiss = (jj == 0 ? 1 : 0) + (jj == a.ny - 1 ? 1 : 0);
#pragma unroll
      for (int m = 0; m < 4; ++m) ci[m] = (rowok && colok[m]) ? 4 - colmiss[m] - rowmiss : 5;
    }
    double w[4];
    noisy_rhs4(jj, bb, ci, w);
    { // phase A: first colour of row jj, every neighbour still old
      const double west = P == 0 ? shfl_up1(x0[3]) : 0.0, east = P == 1 ? shfl_dn1(x0[0]) : 0.0;
      update<P>(x0, xs, xn, west, east, w[P], ci[P]);
      update<P + 2>(x0, xs, xn, west, east, w[P + 2], ci[P + 2]);
    }
    { // phase B: second colour of row jj-1 (the same columns), every neighbour new
      const double west = P == 0 ? shfl_up1(xs[3]) : 0.0, east = P == 1 ? shfl_dn1(xs[0]) : 0.0;
      update<P>(xs, xss, x0, west, east, wk[0], cis[P]);
      update<P + 2>(xs, xss, x0, west, east, wk[1], cis[P + 2]);
    }
    wk[0] = w[1 - P];
    wk[1] = w[3 - P];
    const int jo = jj - 1; // row jj-1 is final
    if (out_lane && jo >= ja && jo < jb) st256(a.xout + (long long)(jo - a.tlo) * a.pitch + c, xs);
  }

  // r = b - A x of the lane's four columns of row jr (stream2d::resid: the assembled row's order S, W, C, E, N); nodes that
  // do not exist give 0
  __device__ __forceinline__ void resid_row(int jr, const double (&row)[4], const double (&south)[4], const double (&north)[4], const double (&bv)[4], double (&r)[4]) const
  {
    const double west = shfl_up1(row[3]), east = shfl_dn1(row[0]), mh = -a.h;
    int          ci[4] = {0, 0, 0, 0};
    if (!INTERIOR) {
      const bool rowok   = jr >= 0 && jr < a.ny;
      const int  rowmiss = (jr == 0 ? 1 : 0) + (jr == a.ny - 1 ? 1 : 0);
#pragma unroll
      for (int m = 0; m < 4; ++m) ci[m] = (rowok && colok[m]) ? 4 - colmiss[m] - rowmiss : 5;
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const double xw = m == 0 ? west : row[m == 0 ? 0 : m - 1];
      const double xe = m == 3 ? east : row[m == 3 ? 3 : m + 1];
      double       ax = 0.0;
      ax = fma(mh, south[m], ax);
      ax = fma(mh, xw, ax);
      ax = fma(INTERIOR ? a.diag : coef[ci[m]].diag, row[m], ax);
      ax = fma(mh, xe, ax);
      ax = fma(mh, north[m], ax);
      r[m] = __dsub_rn(bv[m], ax);
      if (!INTERIOR && ci[m] == 5) r[m] = 0.0;
    }
  }
  // b_c row J (fine centre row 2J) from the residual rows 2J-1, 2J, 2J+1 (restrict_kernel: ascending fine index)
  __device__ __forceinline__ void emit_coarse(int J, const double (&rs)[4], const double (&rm)[4], const double (&rn)[4]) const
  {
    const double rsw = shfl_up1(rs[3]), rmw = shfl_up1(rm[3]), rnw = shfl_up1(rn[3]);
    if (!out_lane) return;
    const int I0 = c >> 1;
    double   *p  = a.bc + (long long)(J - a.ctlo) * a.cpitch + I0;
#pragma unroll
    for (int q = 0; q < 2; ++q) { // coarse columns c/2 (fine c) and c/2 + 1 (fine c + 2)
      const double sW = q == 0 ? rsw : rs[1], sC = q == 0 ? rs[0] : rs[2], sE = q == 0 ? rs[1] : rs[3];
      const double cW = q == 0 ? rmw : rm[1], cC = q == 0 ? rm[0] : rm[2], cE = q == 0 ? rm[1] : rm[3];
      const double nW = q == 0 ? rnw : rn[1], nC = q == 0 ? rn[0] : rn[2], nE = q == 0 ? rn[1] : rn[3];
      double       s  = 0.0;
      s = fma(0.25, sW, s);
      s = fma(0.5, sC, s);
      s = fma(0.25, sE, s);
      s = fma(0.5, cW, s);
      s = fma(1.0, cC, s);
      s = fma(0.5, cE, s);
      s = fma(0.25, nW, s);
      s = fma(0.5, nC, s);
      s = fma(0.25, nE, s);
      if (INTERIOR || I0 + q < a.cnx) p[q] = s;
    }
  }
};

// smem layout per CTA: [WARPS][STAGES] stages of STAGE_BYTES | tables | coef | mbarriers
template <int WARPS, int STAGES> constexpr size_t smem_bytes() { return (size_t)WARPS * STAGES * STAGE_BYTES + sizeof(fastnormal::SharedTables) + 6 * sizeof(Coef) + (size_t)WARPS * (STAGES + 1) * 8 + 1024; }

// RESTRICT: the pass also forms b_c = P^T (b - A xout) (stream2d.cuh's fused residual + restriction on the TMA structure):
// after row step jj rows <= jj-1 are final, so the residual of row jj-2 follows, and a coarse row is emitted when the three
// residual rows around its centre row 2J are there.  The band runs three rows ahead of and two rows past its outputs.  The
// right-hand side rows of the residual are re-read from the PREVIOUS ring stage, which is therefore refilled one iteration
// later than in the plain sweep (prefetch distance STAGES-1).
template <int NOISE, bool INTERIOR, int STAGES, bool RESTRICT = false>
__device__ __forceinline__ void run_warp(const Args &a, const fastnormal::Tables &ft, const Coef *coef, uint32_t ring, uint32_t bars, int lane, const Item it)
{
  const int c0 = it.strip * STRIP_OUT - 4, c = c0 + 4 * lane;
  Warp<NOISE, INTERIOR> W(a, ft, coef, lane, c, it.ja, it.jb);
  // first step row: ja-lead or ja-lead-1, whichever makes (J0 + flip) even, so that the unrolled pair is always (P=0, P=1)
  constexpr int lead = RESTRICT ? 3 : 1;
  const int J0 = (((it.ja - lead + a.flip) & 1) == 0) ? it.ja - lead : it.ja - lead - 1;
  const int N  = it.jb + (RESTRICT ? 2 : 0) - J0 + 1; // row steps J0 .. jb (+2)
  const int T  = (N + 1) >> 1;    // stages: stage t feeds steps J0+2t, J0+2t+1 with x rows J0+2t+1, J0+2t+2 and b rows J0+2t, J0+2t+1
  const uint32_t bytes = a.has_b ? STAGE_BYTES : STAGE_BYTES / 2;
  const uint32_t bar_pro = bars + STAGES * 8;
  const bool     swz     = a.swizzle != 0;

  auto issue = [&](int t) {
    const int      s   = t % STAGES;
    const uint32_t dst = ring + s * STAGE_BYTES, bar = bars + s * 8;
    mbar_expect_tx(bar, bytes);
    if (a.swizzle) {
      tma_load_3d(dst, &a.tm_x, 0, c0 >> 2, J0 + 2 * t + 1 - a.tlo, bar);
      if (a.has_b) tma_load_3d(dst + STAGE_ROWS * ROW_BYTES, &a.tm_b, 0, c0 >> 2, J0 + 2 * t - a.tlo, bar);
    } else {
      tma_load_3d(dst, &a.tm_x, c0, J0 + 2 * t + 1 - a.tlo, 0, bar);
      if (a.has_b) tma_load_3d(dst + STAGE_ROWS * ROW_BYTES, &a.tm_b, c0, J0 + 2 * t - a.tlo, 0, bar);
    }
  };
  // prologue rows J0-1, J0 travel through the x half of the LAST ring slot, whose first real stage is issued afterwards
  if (lane == 0) {
    mbar_expect_tx(bar_pro, STAGE_ROWS * ROW_BYTES);
    if (a.swizzle) tma_load_3d(ring + (STAGES - 1) * STAGE_BYTES, &a.tm_x, 0, c0 >> 2, J0 - 1 - a.tlo, bar_pro);
    else tma_load_3d(ring + (STAGES - 1) * STAGE_BYTES, &a.tm_x, c0, J0 - 1 - a.tlo, 0, bar_pro);
    for (int t = 0; t < STAGES - 1 && t < T; ++t) issue(t);
  }
  double xss[4] = {0, 0, 0, 0}, xs[4], x0[4], wk[2] = {0, 0};
  int    cis[4] = {5, 5, 5, 5}; // coefficient classes of row jj-1
  mbar_wait(bar_pro, 0);
  {
    const uint32_t p = ring + (STAGES - 1) * STAGE_BYTES;
    lds256(p, lane, xs, swz);
    lds256(p + ROW_BYTES, lane, x0, swz);
  }
  __syncwarp();
  if (lane == 0 && STAGES - 1 < T) issue(STAGES - 1);

  // x_old row j = xin row j + (P xc) row j, in MatInterpolateAdd's order: s = x; s = fma(w, xc_J, s) over ascending coarse index
  auto prolong = [&](int j, double (&out)[4]) {
    if (a.xc == nullptr) return; // warp-uniform
    const int    Jlo = j >> 1, nJ = (j & 1) ? 2 : 1;
    const double wj = (j & 1) ? 0.5 : 1.0, wh = 0.5 * wj;
    const int    I0 = c >> 1; // c = 0 mod 4: fine columns c .. c+3 see coarse columns I0, I0+1, I0+2
    if (INTERIOR) {
      const double *p = a.xc + (long long)(Jlo - a.ctlo) * a.cpitch + I0;
      for (int q = 0; q < nJ; ++q, p += a.cpitch) {
        const double c0v = p[0], c1v = p[1], c2v = p[2];
        out[0] = fma(wj, c0v, out[0]);
        out[1] = fma(wh, c1v, fma(wh, c0v, out[1]));
        out[2] = fma(wj, c1v, out[2]);
        out[3] = fma(wh, c2v, fma(wh, c1v, out[3]));
      }
      return;
    }
    if (j < 0 || j >= a.ny) return;
    for (int q = 0; q < nJ; ++q) {
      const int J = Jlo + q;
      if (J >= a.cny) continue;
      double        cv[3];
      const double *p = a.xc + (long long)(J - a.ctlo) * a.cpitch + I0;
#pragma unroll
      for (int m = 0; m < 3; ++m) cv[m] = (I0 + m >= 0 && I0 + m < a.cnx) ? p[m] : 0.0;
      if (c >= 0 && c < a.nx) out[0] = fma(wj, cv[0], out[0]);
      if (c + 1 >= 0 && c + 1 < a.nx) {
        out[1] = fma(wh, cv[0], out[1]);
        if (I0 + 1 < a.cnx) out[1] = fma(wh, cv[1], out[1]);
      }
      if (c + 2 >= 0 && c + 2 < a.nx) out[2] = fma(wj, cv[1], out[2]);
      if (c + 3 >= 0 && c + 3 < a.nx) {
        out[3] = fma(wh, cv[1], out[3]);
        if (I0 + 2 < a.cnx) out[3] = fma(wh, cv[2], out[3]);
      }
    }
  };
  auto prefetch_coarse = [&](int j) { // the coarse rows that fine rows j, j+1 will read
    if (!INTERIOR || a.xc == nullptr) return;
    const double *p = a.xc + (long long)((j >> 1) - a.ctlo) * a.cpitch + (c >> 1);
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p + 2));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p + a.cpitch));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p + a.cpitch + 2));
  };
  prolong(J0 - 1, xs);
  prolong(J0, x0);
  prefetch_coarse(J0 + 1);

  auto row_classes = [&](int jj, int (&ci)[4]) {
    if (INTERIOR) return;
    const bool rowok   = jj >= 0 && jj < a.ny;
    const int  rowmiss = (jj == 0 ? 1 : 0) + (jj == a.ny - 1 ? 1 : 0);
#pragma unroll
    for (int m = 0; m < 4; ++m) ci[m] = (rowok && W.colok[m]) ? 4 - W.colmiss[m] - rowmiss : 5;
  };

  double xsss[4] = {0, 0, 0, 0};                         // RESTRICT: row jj-3
  double r1[4] = {0, 0, 0, 0}, r2[4] = {0, 0, 0, 0};      // RESTRICT: residual rows jr-1, jr-2 of the residual row jr being formed
  // residual of row jr (rows jr-1 .. jr+1 final) with the right-hand side row at shared-memory address bsrc, then the coarse row
  // whose centre is row jr-1
  auto residual_step = [&](int jr, const double (&south)[4], const double (&row)[4], const double (&north)[4], uint32_t bsrc, bool have_b) {
    double bv[4] = {0, 0, 0, 0}, r0[4];
    if (have_b) lds256(bsrc, lane, bv, swz);
    W.resid_row(jr, row, south, north, bv, r0);
    const int jc = jr - 1;
    if ((jc & 1) == 0 && jc >= it.ja && jc < it.jb) W.emit_coarse(jc >> 1, r2, r1, r0);
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      r2[m] = r1[m];
      r1[m] = r0[m];
    }
  };

  int jj = J0;
  for (int t = 0; t < T; ++t, jj += 2) {
    const int      s   = t % STAGES;
    const uint32_t src = ring + s * STAGE_BYTES;
    const uint32_t prv = ring + ((t + STAGES - 1) % STAGES) * STAGE_BYTES + 2 * ROW_BYTES; // b rows jj-2, jj-1 (stage t-1)
    mbar_wait(bars + s * 8, (uint32_t)(t / STAGES) & 1u);
    double xa[4], ba[4] = {0, 0, 0, 0};
    lds256(src, lane, xa, swz);
    if (a.has_b) lds256(src + 2 * ROW_BYTES, lane, ba, swz);
    prolong(jj + 1, xa);
    if (jj + 4 <= it.jb) prefetch_coarse(jj + 3);
    W.template step<0>(jj, xss, xs, x0, xa, ba, wk, cis);
    row_classes(jj, cis);
    if (RESTRICT) residual_step(jj - 2, xsss, xss, xs, prv, a.has_b && t > 0);
    if (2 * t + 1 < N) {
      double xb[4], bb[4] = {0, 0, 0, 0};
      lds256(src + ROW_BYTES, lane, xb, swz);
      if (a.has_b) lds256(src + 3 * ROW_BYTES, lane, bb, swz);
      prolong(jj + 2, xb);
      if (!RESTRICT) {
        __syncwarp();
        if (lane == 0 && t + STAGES < T) issue(t + STAGES);
      }
      W.template step<1>(jj + 1, xs, x0, xa, xb, bb, wk, cis);
      row_classes(jj + 1, cis);
      if (RESTRICT) residual_step(jj - 1, xss, xs, x0, prv + ROW_BYTES, a.has_b && t > 0);
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        if (RESTRICT) xsss[m] = xs[m];
        xss[m] = x0[m];
        xs[m]  = xa[m];
        x0[m]  = xb[m];
      }
    }
    if (RESTRICT) { // stage t-1 has now been read for the last time: refill its slot
      __syncwarp();
      if (lane == 0 && t >= 1 && t - 1 + STAGES < T) issue(t - 1 + STAGES);
    }
  }
}

template <int NOISE, int WARPS, int STAGES, int MINB, bool RESTRICT = false> __global__ void __launch_bounds__(WARPS * 32, MINB) sweep2d_kernel(const __grid_constant__ Args a)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char *base = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  fastnormal::SharedTables *fts  = reinterpret_cast<fastnormal::SharedTables *>(base + (size_t)WARPS * STAGES * STAGE_BYTES);
  Coef                     *coef = reinterpret_cast<Coef *>(fts + 1);
  unsigned long long       *bar  = reinterpret_cast<unsigned long long *>(coef + 6);
  const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
  pdl_launch_dependents();
  const fastnormal::Tables ft = fastnormal::load_tables(*fts);
  if (threadIdx.x < 6) coef[threadIdx.x] = a.coef[threadIdx.x];
  if (lane == 0) {
    for (int s = 0; s <= STAGES; ++s) mbar_init(smem_u32(bar + wl * (STAGES + 1) + s), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  const int w = blockIdx.x * WARPS + wl;
  if (w >= a.nitems) return;
  const Item it = a.items[w];
  const int  c0 = it.strip * STRIP_OUT - 4;
  const int  J0 = it.ja - (RESTRICT ? 4 : 2), jl = it.jb + (RESTRICT ? 2 : 0); // lowest possible first step row, last step row
  // every node the warp updates exists and has all four neighbours, and every row it reads is owned
  const bool interior = c0 >= 1 && c0 + 127 <= a.nx - 2 && J0 >= 1 && jl <= a.ny - 2 && J0 - 1 >= a.tlo && jl + 2 < a.thi;
  const uint32_t ring = smem_u32(base) + wl * STAGES * STAGE_BYTES, bars = smem_u32(bar + wl * (STAGES + 1));
  pdl_wait(); // everything above reads launch constants only (common.hpp: programmatic dependent launch)
  if (interior) run_warp<NOISE, true, STAGES, RESTRICT>(a, ft, coef, ring, bars, lane, it);
  else run_warp<NOISE, false, STAGES, RESTRICT>(a, ft, coef, ring, bars, lane, it);
}

} // namespace sweep2d
