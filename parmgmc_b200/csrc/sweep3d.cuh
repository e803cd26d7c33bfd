// sweep3d.cuh -- TMA-fed fused red-black Gibbs sweep for the matrix-free 3D 7-point operator (K1 of SURVEY 8(d)): both
// colours, the noise and the right-hand-side perturbation in ONE pass over memory (the reference needs
// VecSetRandomStandardNormal, VecPointwiseMult, VecAXPY and the two colour phases of MCSORApply:
// src/pc_mcgibbs.c:119-128, src/mc_sor.c:257-271).
//
// 2.5D blocking.  A CTA of NW warps owns a tile of 120 columns (x) by NW-2 rows (y) and walks along z through a band of
// planes.  A warp is one grid row of the tile (lane l owns columns c0+4l .. c0+4l+3, lanes 0 / 31 are halo columns, as in
// sweep2d.cuh); warps 0 and NW-1 are halo rows.  Along z a thread keeps a rolling window of three planes in registers.
// At plane step k the first colour of plane k is updated (all its neighbours are still old), then the second colour of
// plane k-1 (all its neighbours are new by then), and plane k-1 is written out of place: 8 B of x and 8 B of b read and
// 8 B written per DOF-update instead of the 48 B a colour-by-colour sweep moves.
//
// Data movement: one elected thread feeds two shared-memory rings with the TMA -- 128 x (NW+2) x 1 boxes of x (plane k+1:
// the "up" neighbours, and one step later the OLD north / south neighbours of plane k+1, read straight from the box) and
// 128 x NW x 1 boxes of b; out-of-grid coordinates are zero-filled by the hardware.  The half-updated plane k is
// published to the neighbouring rows through a double-buffered shared-memory array; one __syncthreads per plane.
// East / west neighbours come from warp shuffles.  Arithmetic per node is exactly lap_sweep_kernel<3>'s (stencil_op.cu),
// fma for fma (a missing neighbour adds h * 0, which is exact), so the result is bit-identical to the colour-by-colour path.
#pragma once
#include <cuda.h>

#include <type_traits>

#include "common.hpp"
#include "fastnormal.cuh"
#include "philox.cuh"
#include "sweep2d.cuh"

namespace sweep3d {

using sweep2d::lds256;
using sweep2d::mbar_expect_tx;
using sweep2d::mbar_init;
using sweep2d::mbar_wait;
using sweep2d::shfl_dn1;
using sweep2d::shfl_up1;
using sweep2d::smem_u32;
using sweep2d::st256;
using sweep2d::tma_load_3d;
using sweep2d::Coef;

constexpr int STRIP_OUT = 120;
constexpr int ROW_BYTES = 128 * 8;
enum { NOISE_NONE = 0, NOISE_TAPE = 1, NOISE_PHILOX = 2 };

// one CTA's work: output columns of strip `strip`, output rows [ya, ya + NW - 2), output planes [ka, kb)
struct Item {
  int strip, ya, ka, kb;
  int narrow; // 1: the last, short strip of a row (at most 56 columns): 16 lanes per grid row, two grid rows per warp
};

struct WsQueue { // work queue of the persistent kernel (sweep3d_ws.cuh): device memory, zero between launches
  unsigned next;
};

struct Args {
  CUtensorMap   tm_x, tm_b; // {4, pitch/4, ny, local planes} FP64 tensors (SWIZZLE_32B); boxes 4 x 32 x (NW+2) x 1 and 4 x 32 x NW x 1
  CUtensorMap   tm_x16, tm_b16; // the same tensors with the boxes of a narrow strip: 4 x 16 x (2NW+2) x 1 and 4 x 16 x 2NW x 1
  int           nx, ny, nz;
  int           slo, shi; // owned planes: the planes that are written, and the planes the injected tape covers
  int           tlo, thi; // planes held by the tensors and by xout: the owned planes plus two ghost planes per side on a slab
  const Item   *items;
  int           nitems;   // persistent kernel: tiles of this launch, drawn from `queue`
  WsQueue      *queue;
  int           pitch;    // row stride of xout
  long long     pplane;   // plane stride of xout = pitch * ny
  int           flip;     // 0: forward sweep (colour (i+j+k) even first); 1: backward
  int           has_b;
  int           swizzle;  // 1: tensors {4, pitch/4, ny, planes} with SWIZZLE_32B; 0: {pitch, ny, planes, 1}, plain rows
  double       *xout;
  const double *tape;     // injected noise of this block: natural layout, local planes
  double        h, idiag, sd, omo; // interior coefficients
  Coef          coef[8];  // by number of existing neighbours; 7 = no such node
  PhiloxKeys    pk;
  uint32_t      call_lo, call_hi;
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *tm, int c0, int c1, int c2, int c3, uint32_t bar)
{
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}
__device__ __forceinline__ double lds64(uint32_t addr)
{
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts64(uint32_t addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }
// byte offset of column M of the lane's segment inside a TMA row (SWIZZLE_32B layout, sweep2d.cuh)
__device__ __forceinline__ uint32_t box_off(int lane, int M, bool swz) { return sweep2d::lane_seg(lane) + (((uint32_t)(M >> 1) * 16u) ^ (swz ? sweep2d::lane_swz(lane) : 0u)) + (uint32_t)(M & 1) * 8u; }
// the half-updated planes are published component-major ([4][32] doubles per row): conflict-free 8-byte accesses
template <int LPR> __device__ __forceinline__ uint32_t new_off(int l, int M) { return (uint32_t)(M * (LPR * 8) + l * 8); }

template <int NW, int SX, int SB, bool WS = false> struct Smem {
  static constexpr int    TREP = NW >= 16 ? 8 : 1; // copies of the Box-Muller tables (fastnormal.cuh): conflict-free gathers where the space is there
  static constexpr int    XSTAGE = (NW + 2) * ROW_BYTES, BSTAGE = NW * ROW_BYTES, NEWBUF = NW * ROW_BYTES;
  static constexpr size_t off_x = 0, off_b = off_x + (size_t)SX * XSTAGE, off_new = off_b + (size_t)SB * BSTAGE, off_z = off_new + 2 * (size_t)NEWBUF; // z: WS only
  static constexpr size_t off_tab = off_z + (WS ? 2 * (size_t)NEWBUF : 0);
  static constexpr size_t off_coef = off_tab + sizeof(fastnormal::SharedTablesT<TREP>), off_bar = off_coef + 8 * sizeof(Coef), total = off_bar + (SX + SB) * 8 + 1024;
};

// src/mc_sor.c:260-268 for column M of `row`: accumulation order of the assembled row (down, south, west, east, north, up)
template <int M, bool INTERIOR>
__device__ __forceinline__ void update(const Args &a, const Coef *coef, double (&row)[4], double down, double south, double north, double up, double west, double east, double w, int ci)
{
  const double xw = M == 0 ? west : row[M == 0 ? 0 : M - 1];
  const double xe = M == 3 ? east : row[M == 3 ? 3 : M + 1];
  double       sum = w;
  sum = fma(a.h, down, sum);
  sum = fma(a.h, south, sum);
  sum = fma(a.h, xw, sum);
  sum = fma(a.h, xe, sum);
  sum = fma(a.h, north, sum);
  sum = fma(a.h, up, sum);
  if (INTERIOR) {
    const double t0 = __dmul_rn(a.omo, row[M]);
    row[M]          = fma(a.idiag, sum, t0);
  } else {
    const Coef   k  = coef[ci];
    const double t0 = __dmul_rn(k.omo, row[M]);
    row[M]          = fma(k.idiag, sum, t0);
  }
}

// both colour phases of one plane step for row parity P (first-colour columns M = P, P+2)
template <int P, bool INTERIOR, int LPR>
__device__ __forceinline__ void phases(const Args &a, const Coef *coef, int lane, bool swz, bool inner_row, const double (&xm2)[4], double (&xm1)[4], double (&x0)[4], const double (&xp1)[4], uint32_t old_s, uint32_t old_n, uint32_t new_s, uint32_t new_n,
                                       const double (&w)[4], const double (&wk)[2], const int (&ci)[4], const int (&cis)[4])
{
  { // phase A: first colour of plane kk, every neighbour still old
    const double west = P == 0 ? shfl_up1(x0[3]) : 0.0, east = P == 1 ? shfl_dn1(x0[0]) : 0.0;
    update<P, INTERIOR>(a, coef, x0, xm1[P], lds64(old_s + box_off(lane, P, swz)), lds64(old_n + box_off(lane, P, swz)), xp1[P], west, east, w[P], ci[P]);
    update<P + 2, INTERIOR>(a, coef, x0, xm1[P + 2], lds64(old_s + box_off(lane, P + 2, swz)), lds64(old_n + box_off(lane, P + 2, swz)), xp1[P + 2], west, east, w[P + 2], ci[P + 2]);
  }
  { // phase B: second colour of plane kk-1 (the same columns), every neighbour new; halo rows take part in the shuffles only
    const double west = P == 0 ? shfl_up1(xm1[3]) : 0.0, east = P == 1 ? shfl_dn1(xm1[0]) : 0.0;
    if (inner_row) {
      update<P, INTERIOR>(a, coef, xm1, xm2[P], lds64(new_s + new_off<LPR>(lane, P)), lds64(new_n + new_off<LPR>(lane, P)), x0[P], west, east, wk[0], cis[P]);
      update<P + 2, INTERIOR>(a, coef, xm1, xm2[P + 2], lds64(new_s + new_off<LPR>(lane, P + 2)), lds64(new_n + new_off<LPR>(lane, P + 2)), x0[P + 2], west, east, wk[1], cis[P + 2]);
    }
  }
}

// WS (warp-specialised, Philox only): warps NW .. 2NW-1 are noise producers.  Producer i generates sqrtdiag * z of row i one
// plane ahead into a double-buffered shared-memory array, stencil warp i reads it: the long FP64 transcendental chains and
// the latency-sensitive stencil part run in different warps, and twice as many warps are resident (setmaxnreg gives the
// producers 40 registers and the stencil warps 88).
//
// ZCONST (edge tiles whose planes all have both z neighbours): the node classes do not change from plane to plane and are
// computed once.  LPR = 16 (narrow strips): a grid row is 16 lanes (14 output lanes = 56 columns) and a warp holds two grid
// rows of the same parity (tile rows 4(i/2) + (i%2) and + 2 for warp i), so that the colour pattern stays warp-uniform; the
// shuffles that cross from one row to the other only reach halo lanes.  The shared-memory stages keep their sizes (a box of
// 2NW+2 rows of 512 B fits a stage of NW+2 rows of 1 KB).
template <int NOISE, bool INTERIOR, int NW, int SX, int SB, bool WS, bool ZCONST = false, int LPR = 32, typename FT>
__device__ __forceinline__ void run_cta(const Args &a, const FT &ft, const Coef *coef, uint32_t sm, const Item it)
{
  using L = Smem<NW, SX, SB, WS>;
  static_assert(LPR == 32 || LPR == 16, "lanes per grid row");
  static_assert(!(INTERIOR && (ZCONST || LPR != 32)), "interior tiles are full-width and need no classes");
  constexpr int RPW = 32 / LPR, ROWS = NW * RPW, RB = LPR * 32; // grid rows per warp, rows of the tile (2 of them halo), bytes per row
  const int  wlane = threadIdx.x & 31, lane = wlane % LPR, warp = threadIdx.x >> 5, wi = WS ? warp % NW : warp;
  const int  w = RPW == 1 ? wi : 4 * (wi >> 1) + (wi & 1) + 2 * (wlane / LPR);
  const bool producer = WS && warp >= NW;
  const int  c0 = it.strip * STRIP_OUT - 4, c = c0 + 4 * lane;
  const int  y = it.ya - 1 + w;
  const bool inner_row = w >= 1 && w <= ROWS - 2;
  const bool out_thread = inner_row && lane >= 1 && lane <= LPR - 2 && (INTERIOR || (c < a.nx && y < a.ny));
  const int  K0 = it.ka - 1, K1 = it.kb;         // plane steps; phase B planes K0 .. K1-1, of which ka .. kb-1 are stored
  const int  nsteps = K1 - K0 + 1;
  const uint32_t bar_x = sm + (uint32_t)L::off_bar, bar_b = bar_x + SX * 8;
  const uint32_t xbytes = (ROWS + 2) * RB, bbytes = ROWS * RB;
  const bool     swz = a.swizzle != 0;
  const CUtensorMap *tmx = LPR == 32 ? &a.tm_x : &a.tm_x16, *tmb = LPR == 32 ? &a.tm_b : &a.tm_b16;

  // x sequence q = 0, 1, ...: plane K0 - 1 + q;  b sequence r = 0, 1, ...: plane K0 + r
  auto issue_x = [&](int q) {
    const uint32_t bar = bar_x + (q % SX) * 8;
    mbar_expect_tx(bar, xbytes);
    if (swz) tma_load_4d(sm + (uint32_t)L::off_x + (q % SX) * L::XSTAGE, tmx, 0, c0 >> 2, it.ya - 2, K0 - 1 + q - a.tlo, bar);
    else tma_load_4d(sm + (uint32_t)L::off_x + (q % SX) * L::XSTAGE, tmx, c0, it.ya - 2, K0 - 1 + q - a.tlo, 0, bar);
  };
  auto issue_b = [&](int r) {
    const uint32_t bar = bar_b + (r % SB) * 8;
    mbar_expect_tx(bar, bbytes);
    if (swz) tma_load_4d(sm + (uint32_t)L::off_b + (r % SB) * L::BSTAGE, tmb, 0, c0 >> 2, it.ya - 1, K0 + r - a.tlo, bar);
    else tma_load_4d(sm + (uint32_t)L::off_b + (r % SB) * L::BSTAGE, tmb, c0, it.ya - 1, K0 + r - a.tlo, 0, bar);
  };
  const int nq = nsteps + 2; // x planes K0-1 .. K1+1
  if (threadIdx.x == 0) {
    for (int q = 0; q < SX && q < nq; ++q) issue_x(q);
    if (a.has_b)
      for (int r = 0; r < SB && r < nsteps; ++r) issue_b(r);
  }

  // column / row parts of the node classification (edge tiles)
  int  colmiss[4] = {0, 0, 0, 0};
  bool colok[4]   = {true, true, true, true};
  const bool rowok   = y >= 0 && y < a.ny;
  const int  rowmiss = (y == 0 ? 1 : 0) + (y == a.ny - 1 ? 1 : 0);
  if (!INTERIOR) {
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      colok[m]   = c + m >= 0 && c + m < a.nx;
      colmiss[m] = (c + m == 0 ? 1 : 0) + (c + m == a.nx - 1 ? 1 : 0);
    }
  }
  int ci0[4] = {0, 0, 0, 0}; // ZCONST: the classes of every plane
  if (ZCONST) {
#pragma unroll
    for (int m = 0; m < 4; ++m) ci0[m] = (rowok && colok[m]) ? 6 - colmiss[m] - rowmiss : 7;
  }
  auto classes = [&](int kk, int (&ci)[4]) {
    if (INTERIOR) return;
    if (ZCONST) {
#pragma unroll
      for (int m = 0; m < 4; ++m) ci[m] = ci0[m];
      return;
    }
    const bool kok   = kk >= 0 && kk < a.nz;
    const int  kmiss = (kk == 0 ? 1 : 0) + (kk == a.nz - 1 ? 1 : 0);
#pragma unroll
    for (int m = 0; m < 4; ++m) ci[m] = (kok && rowok && colok[m]) ? 6 - colmiss[m] - rowmiss - kmiss : 7;
  };
  // sqrtdiag * z of row (y, kk), the first rounding of w = b + sqrtdiag z (src/pc_mcgibbs.c:124-126)
  auto scaled_normals = [&](int kk, const int (&ci)[4], double (&zs)[4]) {
    const long long quad = (((long long)kk * a.ny + y) * a.pitch + c) >> 2;
    uint32_t        w0, w1, w2, w3;
    double          z[4];
    philox4x32_10_keys((uint32_t)quad, (uint32_t)((unsigned long long)quad >> 32), a.call_lo, a.call_hi, a.pk, w0, w1, w2, w3);
    fastnormal::box_muller(ft, w0, w1, z[0], z[1]);
    fastnormal::box_muller(ft, w2, w3, z[2], z[3]);
#pragma unroll
    for (int m = 0; m < 4; ++m) zs[m] = __dmul_rn(z[m], INTERIOR ? a.sd : coef[ci[m]].sd);
  };
  if (producer) { // ---- noise producers: z of step s goes to buffer s & 1, one step ahead of its use ----
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    for (int s = 0; s <= nsteps; ++s) { // iteration s produces step s, then joins the barrier that ends step s-1 (s = 0: the prologue's)
      if (s < nsteps) {
        int    ci[4] = {0, 0, 0, 0};
        double zs[4];
        classes(K0 + s, ci);
        scaled_normals(K0 + s, ci, zs);
        const uint32_t p = sm + (uint32_t)L::off_z + (s & 1) * L::NEWBUF + w * RB;
#pragma unroll
        for (int m = 0; m < 4; ++m) sts64(p + new_off<LPR>(lane, m), zs[m]);
      }
      __syncthreads();
    }
    return;
  }
  if (WS) asm volatile("setmaxnreg.inc.sync.aligned.u32 88;");

  double A[4] = {0, 0, 0, 0}, B[4], C[4], D[4], wk[2] = {0, 0}; // rolling window: planes kk-2, kk-1, kk and the incoming kk+1
  int    cis[4] = {ZCONST ? ci0[0] : 7, ZCONST ? ci0[1] : 7, ZCONST ? ci0[2] : 7, ZCONST ? ci0[3] : 7};
  const uint32_t own = (uint32_t)((w + 1) * RB); // this warp's row inside an x box
  auto publish = [&](int plane, const double (&v)[4]) {
    const uint32_t p = sm + (uint32_t)L::off_new + (plane & 1) * L::NEWBUF + w * RB;
#pragma unroll
    for (int m = 0; m < 4; ++m) sts64(p + new_off<LPR>(lane, m), v[m]);
  };
  mbar_wait(bar_x + 0 * 8, 0);
  lds256(sm + (uint32_t)L::off_x + 0 * L::XSTAGE + own, lane, B, swz);
  mbar_wait(bar_x + (1 % SX) * 8, 0);
  lds256(sm + (uint32_t)L::off_x + (1 % SX) * L::XSTAGE + own, lane, C, swz);
  publish(K0 - 1, B); // phase B of plane K0-1 is never stored: any finite values will do
  __syncthreads();
  if (threadIdx.x == 0 && SX < nq) issue_x(SX); // slot of q = 0 is free again

  // one plane step for a row whose first-colour columns of plane kk are M = P, P+2; xp1 receives plane kk+1
  auto step = [&](auto ptag, int s, const double (&xm2)[4], double (&xm1)[4], double (&x0)[4], double (&xp1)[4]) {
    constexpr int P = decltype(ptag)::value;
    const int kk = K0 + s;
    const int q0 = s + 1, q1 = s + 2; // x sequence numbers of planes kk, kk+1
    const uint32_t xs0 = sm + (uint32_t)L::off_x + (q0 % SX) * L::XSTAGE, xs1 = sm + (uint32_t)L::off_x + (q1 % SX) * L::XSTAGE;
    double bb[4] = {0, 0, 0, 0};
    mbar_wait(bar_x + (q1 % SX) * 8, (uint32_t)(q1 / SX) & 1u);
    lds256(xs1 + own, lane, xp1, swz);
    if (a.has_b) {
      mbar_wait(bar_b + (s % SB) * 8, (uint32_t)(s / SB) & 1u);
      lds256(sm + (uint32_t)L::off_b + (s % SB) * L::BSTAGE + w * RB, lane, bb, swz);
    }
    int ci[4] = {0, 0, 0, 0};
    classes(kk, ci);
    // w = b + sqrtdiag z (src/pc_mcgibbs.c:124-126: two roundings)
    double wv[4];
    if (NOISE == NOISE_NONE) {
#pragma unroll
      for (int m = 0; m < 4; ++m) wv[m] = bb[m];
    } else if (WS) {
      const uint32_t p = sm + (uint32_t)L::off_z + (s & 1) * L::NEWBUF + w * RB;
#pragma unroll
      for (int m = 0; m < 4; ++m) wv[m] = __dadd_rn(lds64(p + new_off<LPR>(lane, m)), bb[m]);
    } else {
      double z[4];
      if (NOISE == NOISE_TAPE) {
        const bool    ok = INTERIOR || (rowok && kk >= a.slo && kk < a.shi);
        const double *p  = a.tape + ((long long)(kk - a.slo) * a.ny + y) * a.nx + c;
#pragma unroll
        for (int m = 0; m < 4; ++m) z[m] = (ok && (INTERIOR || colok[m])) ? p[m] : 0.0;
      } else {
        const long long quad = (((long long)kk * a.ny + y) * a.pitch + c) >> 2;
        uint32_t        w0, w1, w2, w3;
        philox4x32_10_keys((uint32_t)quad, (uint32_t)((unsigned long long)quad >> 32), a.call_lo, a.call_hi, a.pk, w0, w1, w2, w3);
        fastnormal::box_muller(ft, w0, w1, z[0], z[1]);
        fastnormal::box_muller(ft, w2, w3, z[2], z[3]);
      }
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const double sd = INTERIOR ? a.sd : coef[ci[m]].sd;
        wv[m]           = __dadd_rn(__dmul_rn(z[m], sd), bb[m]);
      }
    }
    const uint32_t old_s = xs0 + own - RB, old_n = xs0 + own + RB;
    const uint32_t nb    = sm + (uint32_t)L::off_new + ((kk - 1) & 1) * L::NEWBUF;
    const uint32_t new_s = nb + (inner_row ? w - 1 : w) * RB, new_n = nb + (inner_row ? w + 1 : w) * RB;
    phases<P, INTERIOR, LPR>(a, coef, lane, swz, inner_row, xm2, xm1, x0, xp1, old_s, old_n, new_s, new_n, wv, wk, ci, cis);
    wk[0] = wv[1 - P];
    wk[1] = wv[3 - P];
#pragma unroll
    for (int m = 0; m < 4; ++m) cis[m] = ci[m];
    const int ko = kk - 1; // plane kk-1 of this row is final
    if (out_thread && ko >= it.ka && ko < it.kb) st256(a.xout + (long long)(ko - a.tlo) * a.pplane + (long long)y * a.pitch + c, xm1);
    publish(kk, x0); // the half-updated plane kk, for the neighbouring rows
    __syncthreads();
    if (threadIdx.x == 0) { // the boxes of plane kk (x) and of this step (b) are free
      if (q0 + SX < nq) issue_x(q0 + SX);
      if (a.has_b && s + SB < nsteps) issue_b(s + SB);
    }
  };
  // the row parity alternates from plane to plane and the rolling window has four slots: steps are unrolled in fours so that
  // the colour pattern is a compile-time constant and the window never moves between registers; every warp of the CTA runs
  // the same number of steps (one barrier each)
  auto pairs = [&](auto qtag) {
    constexpr int Q = decltype(qtag)::value;
    using T0 = std::integral_constant<int, Q>;
    using T1 = std::integral_constant<int, 1 - Q>;
    int s = 0;
    for (; s + 3 < nsteps; s += 4) {
      step(T0{}, s, A, B, C, D);
      step(T1{}, s + 1, B, C, D, A);
      step(T0{}, s + 2, C, D, A, B);
      step(T1{}, s + 3, D, A, B, C);
    }
    if (s < nsteps) {
      step(T0{}, s, A, B, C, D);
      if (s + 1 < nsteps) {
        step(T1{}, s + 1, B, C, D, A);
        if (s + 2 < nsteps) step(T0{}, s + 2, C, D, A, B);
      }
    }
  };
  if (((y + K0 + a.flip) & 1) == 0) pairs(std::integral_constant<int, 0>{});
  else pairs(std::integral_constant<int, 1>{});
}

template <int NOISE, int NW, int SX, int SB, int MINB, bool WS = false> __global__ void __launch_bounds__(NW * 32 * (WS ? 2 : 1), MINB) sweep3d_kernel(const __grid_constant__ Args a)
{
  using L = Smem<NW, SX, SB, WS>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char *base = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  auto *fts = reinterpret_cast<fastnormal::SharedTablesT<L::TREP> *>(base + L::off_tab);
  Coef                     *coef = reinterpret_cast<Coef *>(base + L::off_coef);
  const auto ft = fastnormal::load_tables(*fts);
  if (threadIdx.x < 8) coef[threadIdx.x] = a.coef[threadIdx.x];
  const uint32_t sm = smem_u32(base);
  if (threadIdx.x == 0) {
    for (int s = 0; s < SX + SB; ++s) mbar_init(sm + (uint32_t)L::off_bar + s * 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  const Item it = a.items[blockIdx.x];
  const int  c0 = it.strip * STRIP_OUT - 4;
  // every node the CTA updates exists and has all six neighbours, and every row / plane it reads exists and is owned
  const bool zconst = it.ka - 1 >= 1 && it.kb <= a.nz - 2 && it.ka - 2 >= a.tlo && it.kb + 1 < a.thi; // every plane the tile touches has both z neighbours
  const bool interior = zconst && !it.narrow && c0 >= 1 && c0 + 127 <= a.nx - 2 && it.ya - 1 >= 1 && it.ya + NW - 2 <= a.ny - 2;
  if (interior) run_cta<NOISE, true, NW, SX, SB, WS>(a, ft, coef, sm, it);
  else if (it.narrow) run_cta<NOISE, false, NW, SX, SB, WS, true, 16>(a, ft, coef, sm, it); // the host makes narrow tiles only where zconst holds
  else if (zconst) run_cta<NOISE, false, NW, SX, SB, WS, true>(a, ft, coef, sm, it);
  else run_cta<NOISE, false, NW, SX, SB, WS>(a, ft, coef, sm, it);
}

} // namespace sweep3d
