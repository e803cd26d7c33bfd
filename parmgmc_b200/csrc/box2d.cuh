// box2d.cuh -- TMA-fed one-pass kernels for the 2D 9-point Galerkin levels of the V-cycle.
//
// One launch does what PCMG does in up to seven passes on such a level (src/pc_gamgmc.c:242-259):
//
//   MODE_RESTRICT  pre-sample (four-colour SOR-Gibbs sweep, src/mc_sor.c:257-285 on w = b + sqrtdiag z, src/pc_sorgibbs.c:81-83),
//                  r = b - A x, b_c = P^T r                     -- r is never stored
//   MODE_PROLONG   x += P x_c, post-sample
//   MODE_PLAIN     one sweep (further sweeps of a level KSP with max_it > 1, symmetric sweeps)
//
// Layout: the level's vectors are PITCHED (row stride = nx rounded up to 4, pad columns kept at zero), so a warp's row
// segment is a legal TMA box and a lane's four columns are one aligned 32-byte access.
//
// Sweep order (box_stream.cuh): colours (i mod 2) + 2 (j mod 2).  Rows of the first parity ("A rows") need only OLD rows
// above and below, rows of the other parity ("B rows") need NEW ones, so a warp walks down its band two rows per step:
// A(e), then B(e-1).  A warp owns 128 columns (lane l: columns c0+4l .. c0+4l+3; lanes 1..30 are written, lanes 0 / 31
// recompute the neighbouring strips' edge columns), keeps the row window in registers and takes east / west / diagonal
// neighbours from warp shuffles.  Rows arrive through the TMA (cp.async.bulk.tensor, one 128 x 2 box of x and of b per step,
// in a per-warp ring of shared-memory stages with mbarriers): nothing in flight costs registers, out-of-grid rows and columns
// are zero-filled by the hardware.
//
// Boundary handling without predicates: the Galerkin coarsening of the constant-coefficient fine operator yields the same
// stencil at every node of a boundary CLASS (row class x column class, each first / interior / last; verified bitwise at
// set-up, BoxOp::detect_classes).  Interior warps read the interior stencil from kernel parameters; the others look up the
// class table in shared memory.  Nodes outside the grid belong to a null class (all zeros): with zero-filled loads they
// compute exact zeros, and a structurally absent neighbour contributes fma(0, 0, s) = s, so the arithmetic per node is
// box_sweep_kernel's / box_apply_kernel's / restrict_kernel's / prolong_kernel's, fma for fma (bit-identical results).
#pragma once
#include <cuda.h>

#include "common.hpp"
#include "fastnormal.cuh"
#include "philox.cuh"
#include "sweep2d.cuh"

namespace box2d {

using sweep2d::lds256;
using sweep2d::mbar_expect_tx;
using sweep2d::mbar_init;
using sweep2d::mbar_wait;
using sweep2d::shfl_dn1;
using sweep2d::shfl_up1;
using sweep2d::smem_u32;
using sweep2d::st256;
using sweep2d::tma_load_3d;

// Columns written per warp.  One sweep invalidates four halo columns on either side of a strip (a colour-2 node of a B row
// needs the new colour-1 values of the A rows next to it, which need the new colour-0 values beside them ...); the fused
// residual + restriction reads the swept iterate two more columns out, so its strips keep eight halo columns per side.
template <int MODE> struct Strip {
  static constexpr int HALO = MODE == 2 ? 8 : 4;
  static constexpr int OUT  = 128 - 2 * HALO;
};
constexpr int ROW_BYTES   = 128 * 8;
constexpr int STAGE_BYTES = 4 * ROW_BYTES; // x rows e+1, e+2 | b rows e-1, e

enum { MODE_PLAIN = 0, MODE_PROLONG = 1, MODE_RESTRICT = 2 };
enum { NOISE_RT = 0, NOISE_PHILOX = 1 }; // RT: none / injected tape, chosen at run time (parity tests)

struct Item {
  int strip, ja, jb; // output columns of strip `strip`, output rows [ja, jb)
};

// one boundary class: NEGATED off-diagonal coefficients in ascending stencil order (centre left out), omega / a_ii,
// sqrt((2-omega)/omega) sqrt(a_ii), 1 - omega, -a_ii
struct __align__(16) Cls {
  double nc[8];
  double idiag, sd, omo, ndiag;
};

struct Args {
  CUtensorMap   tm_x, tm_b; // {pitch, rows held} FP64 tensors, box 128 x 2
  int           nx, ny, pitch; // GLOBAL grid; row stride
  int           tlo;           // first grid row held by the tensors and by xout (a slab holds its rows plus four ghost rows per side)
  int           ctlo;          // first coarse row held by xc / bc
  const Item   *items;
  int           nitems;
  int           has_x, has_b; // 0: the iterate / right-hand side is zero and is not read
  double       *xout;
  const double *xc;                      // MODE_PROLONG: coarse iterate
  double       *bc;                      // MODE_RESTRICT: coarse right-hand side
  int           cnx, cny, cpitch, ccols; // coarse grid, row stride of xc / bc, columns of bc to store (pads included when pitched)
  int           mode;                    // NOISE_RT: PMG_NOISE_NONE | PMG_NOISE_INJECTED
  const double *tape;                    // natural layout (row stride nx)
  Cls           in;                      // the interior class
  Cls           cls[16];                 // [4 row class + column class]; class 3 = outside the grid
  PhiloxKeys    pk;
  uint32_t      call_lo, call_hi;
};

__device__ __forceinline__ void st128(double *p, double v0, double v1) { asm volatile("st.global.v2.f64 [%2], {%0,%1};" ::"d"(v0), "d"(v1), "l"(p) : "memory"); }

template <int NOISE, int MODE, int PC, bool GENERAL> struct Warp {
  const Args              &a;
  const fastnormal::Tables ft;
  const Cls               *cls; // shared-memory copy of a.cls
  int                      lane, c;
  int                      colcls[4];

  __device__ __forceinline__ Warp(const Args &a_, const fastnormal::Tables &ft_, const Cls *cls_, int lane_, int c_) : a(a_), ft(ft_), cls(cls_), lane(lane_), c(c_)
  {
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int i = c + m;
      colcls[m]   = (i < 0 || i >= a.nx) ? 3 : (i == 0 ? 0 : (i == a.nx - 1 ? 2 : 1));
    }
  }
  __device__ __forceinline__ int rowcls(int j) const { return (j < 0 || j >= a.ny) ? 3 : (j == 0 ? 0 : (j == a.ny - 1 ? 2 : 1)); }

  // the four normals of row j
  __device__ __forceinline__ void noise_row(int j, double (&z)[4]) const
  {
    if (NOISE == NOISE_PHILOX) {
      const long long quad = ((long long)j * a.pitch + c) >> 2;
      uint32_t        w0, w1, w2, w3;
      philox4x32_10_keys((uint32_t)quad, (uint32_t)((unsigned long long)quad >> 32), a.call_lo, a.call_hi, a.pk, w0, w1, w2, w3);
      fastnormal::box_muller(ft, w0, w1, z[0], z[1]);
      fastnormal::box_muller(ft, w2, w3, z[2], z[3]);
    } else {
      const bool    rowok = a.mode == PMG_NOISE_INJECTED && j >= 0 && j < a.ny;
      const double *p     = a.tape + (long long)j * a.nx + c;
#pragma unroll
      for (int m = 0; m < 4; ++m) z[m] = (rowok && c + m >= 0 && c + m < a.nx) ? p[m] : 0.0;
    }
  }

  template <int Q> static __device__ __forceinline__ double col(const double (&X)[4], double xw, double xe) { return Q < 0 ? xw : (Q > 3 ? xe : X[Q < 0 ? 0 : (Q > 3 ? 3 : Q)]); }

  // one node (box_sweep_kernel): sum = w - sum_s c_s x_s in ascending stencil order, x = omo x + idiag sum
  template <int M> __device__ __forceinline__ void node(const Cls &k, double (&T)[4], const double (&S)[4], const double (&N)[4], double sw, double se, double tw, double te, double nw, double ne, double bval, double z) const
  {
    const double v0 = col<M - 1>(S, sw, se), v1 = col<M>(S, sw, se), v2 = col<M + 1>(S, sw, se);
    const double v3 = col<M - 1>(T, tw, te), v5 = col<M + 1>(T, tw, te);
    const double v6 = col<M - 1>(N, nw, ne), v7 = col<M>(N, nw, ne), v8 = col<M + 1>(N, nw, ne);
    double       sum;
    if (NOISE == NOISE_PHILOX) sum = __dadd_rn(__dmul_rn(z, k.sd), bval);
    else sum = a.mode == PMG_NOISE_NONE ? bval : __dadd_rn(__dmul_rn(z, k.sd), bval);
    sum = fma(k.nc[0], v0, sum);
    sum = fma(k.nc[1], v1, sum);
    sum = fma(k.nc[2], v2, sum);
    sum = fma(k.nc[3], v3, sum);
    sum = fma(k.nc[4], v5, sum);
    sum = fma(k.nc[5], v6, sum);
    sum = fma(k.nc[6], v7, sum);
    sum = fma(k.nc[7], v8, sum);
    const double t0 = __dmul_rn(k.omo, T[M]);
    T[M]            = fma(k.idiag, sum, t0);
  }
  template <int M> __device__ __forceinline__ const Cls &cls_of(int rc) const { return GENERAL ? cls[4 * rc + colcls[M]] : a.in; }

  // both colours of row j: columns of parity PC first, then the others (which see the first ones updated)
  __device__ __forceinline__ void row_update(int j, double (&T)[4], const double (&S)[4], const double (&N)[4], const double (&bv)[4], const double (&z)[4]) const
  {
    const int    rc = GENERAL ? rowcls(j) : 1;
    const double sw = shfl_up1(S[3]), se = shfl_dn1(S[0]), nw = shfl_up1(N[3]), ne = shfl_dn1(N[0]);
    if (PC == 0) {
      const double tw = shfl_up1(T[3]);
      node<0>(cls_of<0>(rc), T, S, N, sw, se, tw, 0.0, nw, ne, bv[0], z[0]);
      node<2>(cls_of<2>(rc), T, S, N, sw, se, tw, 0.0, nw, ne, bv[2], z[2]);
      const double te = shfl_dn1(T[0]);
      node<1>(cls_of<1>(rc), T, S, N, sw, se, 0.0, te, nw, ne, bv[1], z[1]);
      node<3>(cls_of<3>(rc), T, S, N, sw, se, 0.0, te, nw, ne, bv[3], z[3]);
    } else {
      const double te = shfl_dn1(T[0]);
      node<1>(cls_of<1>(rc), T, S, N, sw, se, 0.0, te, nw, ne, bv[1], z[1]);
      node<3>(cls_of<3>(rc), T, S, N, sw, se, 0.0, te, nw, ne, bv[3], z[3]);
      const double tw = shfl_up1(T[3]);
      node<0>(cls_of<0>(rc), T, S, N, sw, se, tw, 0.0, nw, ne, bv[0], z[0]);
      node<2>(cls_of<2>(rc), T, S, N, sw, se, tw, 0.0, nw, ne, bv[2], z[2]);
    }
  }

  // r = b - A x of column M (box_apply_kernel: ax accumulated from 0 over the nine entries in ascending order; the negated
  // coefficients give -ax with the same roundings)
  template <int M> __device__ __forceinline__ double resid(const Cls &k, const double (&T)[4], const double (&S)[4], const double (&N)[4], double sw, double se, double tw, double te, double nw, double ne, double bval) const
  {
    double acc = 0.0;
    acc = fma(k.nc[0], col<M - 1>(S, sw, se), acc);
    acc = fma(k.nc[1], col<M>(S, sw, se), acc);
    acc = fma(k.nc[2], col<M + 1>(S, sw, se), acc);
    acc = fma(k.nc[3], col<M - 1>(T, tw, te), acc);
    acc = fma(k.ndiag, T[M], acc);
    acc = fma(k.nc[4], col<M + 1>(T, tw, te), acc);
    acc = fma(k.nc[5], col<M - 1>(N, nw, ne), acc);
    acc = fma(k.nc[6], col<M>(N, nw, ne), acc);
    acc = fma(k.nc[7], col<M + 1>(N, nw, ne), acc);
    return __dadd_rn(bval, acc);
  }
  __device__ __forceinline__ void resid_row(int j, const double (&T)[4], const double (&S)[4], const double (&N)[4], const double (&bv)[4], double (&r)[4]) const
  {
    const int    rc = GENERAL ? rowcls(j) : 1;
    const double sw = shfl_up1(S[3]), se = shfl_dn1(S[0]), tw = shfl_up1(T[3]), te = shfl_dn1(T[0]), nw = shfl_up1(N[3]), ne = shfl_dn1(N[0]);
    r[0] = resid<0>(cls_of<0>(rc), T, S, N, sw, se, tw, te, nw, ne, bv[0]);
    r[1] = resid<1>(cls_of<1>(rc), T, S, N, sw, se, tw, te, nw, ne, bv[1]);
    r[2] = resid<2>(cls_of<2>(rc), T, S, N, sw, se, tw, te, nw, ne, bv[2]);
    r[3] = resid<3>(cls_of<3>(rc), T, S, N, sw, se, tw, te, nw, ne, bv[3]);
  }

  // b_c row J (fine centre row 2J) from the residual rows 2J-1, 2J, 2J+1 (restrict_kernel: ascending fine index)
  __device__ __forceinline__ void emit_coarse(int J, const double (&rs)[4], const double (&rm)[4], const double (&rn)[4], bool out_lane) const
  {
    const double rsw = shfl_up1(rs[3]), rmw = shfl_up1(rm[3]), rnw = shfl_up1(rn[3]);
    if (!out_lane) return;
    double acc[2];
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
      acc[q] = s;
    }
    const int I0 = c >> 1;
    double   *p  = a.bc + (long long)(J - a.ctlo) * a.cpitch + I0;
    if (!GENERAL && (a.cpitch & 1) == 0) st128(p, acc[0], acc[1]);
    else { // pad columns of a pitched coarse vector are written as zeros (an even fine row length leaves a real residual beside them)
      if (I0 < a.ccols) p[0] = I0 < a.cnx ? acc[0] : 0.0;
      if (I0 + 1 < a.ccols) p[1] = I0 + 1 < a.cnx ? acc[1] : 0.0;
    }
  }

  // x_old row j = xin row j + (P xc) row j, in prolong_kernel's order: s = x; s = fma(w, xc_J, s) over ascending coarse index
  __device__ __forceinline__ void prolong(int j, double (&out)[4]) const
  {
    const int    Jlo = j >> 1, nJ = (j & 1) ? 2 : 1;
    const double wj = (j & 1) ? 0.5 : 1.0, wh = 0.5 * wj;
    const int    I0 = c >> 1; // c = 0 mod 4: fine columns c .. c+3 see coarse columns I0, I0+1, I0+2
    if (!GENERAL) {
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
  }
  __device__ __forceinline__ void prefetch_coarse(int j) const // the coarse rows that fine rows j, j+1 will read
  {
    if (GENERAL) return;
    const double *p = a.xc + (long long)((j >> 1) - a.ctlo) * a.cpitch + (c >> 1);
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p + 2));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p + a.cpitch));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p + a.cpitch + 2));
  }
};

// first / last A row of a band that must leave rows [sa, sb) final
template <int PC> __device__ __forceinline__ void band_steps(int sa, int sb, int &e0, int &elast)
{
  e0 = sa - 1;
  if ((e0 & 1) != PC) ++e0;
  elast = sb;
  if ((elast & 1) != PC) --elast;
}
// rows that must be final for the outputs of band [ja, jb): the fused residual + restriction needs one more row below and
// reaches further up (the coarse row of centre 2J is complete one (PC = 0) or two (PC = 1) steps after row 2J is swept)
template <int MODE, int PC> __device__ __forceinline__ void band_range(int ja, int jb, int &sa, int &sb)
{
  sa = MODE == MODE_RESTRICT ? ja - 2 : ja;
  sb = MODE == MODE_RESTRICT ? jb + 1 + PC : jb;
}

template <int WARPS, int STAGES> constexpr size_t smem_bytes() { return (size_t)WARPS * STAGES * STAGE_BYTES + sizeof(fastnormal::SharedTables) + 16 * sizeof(Cls) + (size_t)WARPS * (STAGES + 1) * 8 + 1024; }

template <int NOISE, int MODE, int PC, bool GENERAL, int STAGES>
__device__ __forceinline__ void run_warp(const Args &a, const fastnormal::Tables &ft, const Cls *cls, uint32_t ring, uint32_t bars, int lane, const Item it)
{
  constexpr int HL = Strip<MODE>::HALO / 4; // halo lanes per side
  const int     c0 = it.strip * Strip<MODE>::OUT - Strip<MODE>::HALO, c = c0 + 4 * lane;
  Warp<NOISE, MODE, PC, GENERAL> W(a, ft, cls, lane, c);
  int sa, sb, e0, elast;
  band_range<MODE, PC>(it.ja, it.jb, sa, sb);
  band_steps<PC>(sa, sb, e0, elast);
  const int      T       = (elast - e0) / 2 + 1; // stage t feeds step e = e0 + 2t with x rows e+1, e+2 and b rows e-1, e
  const uint32_t bytes   = (a.has_x ? 2 * ROW_BYTES : 0) + (a.has_b ? 2 * ROW_BYTES : 0);
  const uint32_t bar_pro = bars + STAGES * 8;
  const bool     out_lane = lane >= HL && lane <= 31 - HL && c < a.pitch;

  auto issue = [&](int t) {
    const int      s   = t % STAGES;
    const uint32_t dst = ring + s * STAGE_BYTES, bar = bars + s * 8;
    mbar_expect_tx(bar, bytes);
    if (a.has_x) tma_load_3d(dst, &a.tm_x, c0, e0 + 2 * t + 1 - a.tlo, 0, bar);
    if (a.has_b) tma_load_3d(dst + 2 * ROW_BYTES, &a.tm_b, c0, e0 + 2 * t - 1 - a.tlo, 0, bar);
  };
  double Rm3[4] = {0, 0, 0, 0}, Rm2[4] = {0, 0, 0, 0}, Rm1[4] = {0, 0, 0, 0}, R0[4] = {0, 0, 0, 0};
  if (bytes == 0) { // nothing to fetch (zero iterate, zero right-hand side): the mbarriers are never armed
  } else if (a.has_x) {
    // prologue rows e0-1, e0 travel through the x half of the LAST ring slot, whose first real stage is issued afterwards
    if (lane == 0) {
      mbar_expect_tx(bar_pro, 2 * ROW_BYTES);
      tma_load_3d(ring + (STAGES - 1) * STAGE_BYTES, &a.tm_x, c0, e0 - 1 - a.tlo, 0, bar_pro);
      for (int t = 0; t < STAGES - 1 && t < T; ++t) issue(t);
    }
    mbar_wait(bar_pro, 0);
    const uint32_t p = ring + (STAGES - 1) * STAGE_BYTES;
    lds256(p, lane, Rm1, false);
    lds256(p + ROW_BYTES, lane, R0, false);
    __syncwarp();
    if (lane == 0 && STAGES - 1 < T) issue(STAGES - 1);
  } else if (lane == 0) {
    for (int t = 0; t < STAGES && t < T; ++t) issue(t);
  }
  if (MODE == MODE_PROLONG) {
    W.prolong(e0 - 1, Rm1);
    W.prolong(e0, R0);
    W.prefetch_coarse(e0 + 1);
  }
  double bPrev[4] = {0, 0, 0, 0};                                          // b row e-2 (MODE_RESTRICT)
  double rP1[4] = {0, 0, 0, 0}, rP2[4] = {0, 0, 0, 0};                     // residual rows e-3, e-4 of the step being done

  // injected noise (a tape in global memory, e.g. the batched prefill of pc.cu) is fetched one step ahead: a load issued in the
  // step that uses it would put a global-memory latency on every band step
  double zAn[4] = {0, 0, 0, 0}, zBn[4] = {0, 0, 0, 0};
  if (NOISE == NOISE_RT) {
    W.noise_row(e0, zAn);
    W.noise_row(e0 - 1, zBn);
  }
  int e = e0;
  for (int t = 0; t < T; ++t, e += 2) {
    const int      s   = t % STAGES;
    const uint32_t src = ring + s * STAGE_BYTES;
    double         xa[4] = {0, 0, 0, 0}, xb[4] = {0, 0, 0, 0}, bA[4] = {0, 0, 0, 0}, bB[4] = {0, 0, 0, 0};
    if (bytes) {
      mbar_wait(bars + s * 8, (uint32_t)(t / STAGES) & 1u);
      if (a.has_x) {
        lds256(src, lane, xa, false);
        lds256(src + ROW_BYTES, lane, xb, false);
      }
      if (a.has_b) {
        lds256(src + 2 * ROW_BYTES, lane, bB, false);
        lds256(src + 3 * ROW_BYTES, lane, bA, false);
      }
      __syncwarp();
      if (lane == 0 && t + STAGES < T) issue(t + STAGES);
    }
    if (MODE == MODE_PROLONG) {
      W.prolong(e + 1, xa);
      W.prolong(e + 2, xb);
      if (e + 2 <= elast) W.prefetch_coarse(e + 3);
    }
    double zA[4], zB[4];
    if (NOISE == NOISE_RT) {
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        zA[m] = zAn[m];
        zB[m] = zBn[m];
      }
      if (t + 1 < T) {
        W.noise_row(e + 2, zAn);
        W.noise_row(e + 1, zBn);
      }
    } else {
      W.noise_row(e, zA);
      W.noise_row(e - 1, zB);
    }
    W.row_update(e, R0, Rm1, xa, bA, zA);      // A row e: neighbours rows e-1, e+1 old
    W.row_update(e - 1, Rm1, Rm2, R0, bB, zB); // B row e-1: neighbour rows e-2, e final
    if (out_lane) {
      if (e >= it.ja && e < it.jb) st256(a.xout + (long long)(e - a.tlo) * a.pitch + c, R0);
      if (e - 1 >= it.ja && e - 1 < it.jb) st256(a.xout + (long long)(e - 1 - a.tlo) * a.pitch + c, Rm1);
    }
    if (MODE == MODE_RESTRICT) {
      double rX[4], rY[4]; // residual rows e-2, e-1 (rows <= e are final)
      W.resid_row(e - 2, Rm2, Rm3, Rm1, bPrev, rX);
      W.resid_row(e - 1, Rm1, Rm2, R0, bB, rY);
      if (PC == 0) { // A rows are the even rows: coarse row of centre e-2
        const int jc = e - 2;
        if (jc >= it.ja && jc < it.jb) W.emit_coarse(jc >> 1, rP1, rX, rY, out_lane);
      } else { // centre e-3
        const int jc = e - 3;
        if (jc >= it.ja && jc < it.jb) W.emit_coarse(jc >> 1, rP2, rP1, rX, out_lane);
      }
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        rP2[m]   = rX[m];
        rP1[m]   = rY[m];
        bPrev[m] = bA[m];
        Rm3[m]   = Rm1[m];
      }
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      Rm2[m] = R0[m];
      Rm1[m] = xa[m];
      R0[m]  = xb[m];
    }
  }
}

template <int NOISE, int MODE, int PC, int WARPS, int STAGES, int MINB> __global__ void __launch_bounds__(WARPS * 32, MINB) box2d_kernel(const __grid_constant__ Args a)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char *base = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  fastnormal::SharedTables *fts = reinterpret_cast<fastnormal::SharedTables *>(base + (size_t)WARPS * STAGES * STAGE_BYTES);
  Cls                      *cls = reinterpret_cast<Cls *>(fts + 1);
  unsigned long long       *bar = reinterpret_cast<unsigned long long *>(cls + 16);
  const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
  pdl_launch_dependents();
  const fastnormal::Tables ft = fastnormal::load_tables(*fts);
  for (int q = threadIdx.x; q < 16 * (int)(sizeof(Cls) / sizeof(double)); q += blockDim.x) reinterpret_cast<double *>(cls)[q] = reinterpret_cast<const double *>(a.cls)[q];
  if (lane == 0) {
    for (int s = 0; s <= STAGES; ++s) mbar_init(smem_u32(bar + wl * (STAGES + 1) + s), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  const int w = blockIdx.x * WARPS + wl;
  if (w >= a.nitems) return;
  const Item it = a.items[w];
  const int  c0 = it.strip * Strip<MODE>::OUT - Strip<MODE>::HALO;
  int        sa, sb, e0, elast;
  band_range<MODE, PC>(it.ja, it.jb, sa, sb);
  band_steps<PC>(sa, sb, e0, elast);
  // every node the warp touches (rows e0-2 .. elast+2, columns c0 .. c0+127) is an interior-class node
  const bool interior = c0 >= 1 && c0 + 127 <= a.nx - 2 && e0 - 2 >= 1 && elast + 2 <= a.ny - 2;
  const uint32_t ring = smem_u32(base) + wl * STAGES * STAGE_BYTES, bars = smem_u32(bar + wl * (STAGES + 1));
  pdl_wait(); // everything above reads launch constants only; the level's vectors belong to the preceding kernels
  if (interior) run_warp<NOISE, MODE, PC, false, STAGES>(a, ft, cls, ring, bars, lane, it);
  else run_warp<NOISE, MODE, PC, true, STAGES>(a, ft, cls, ring, bars, lane, it);
}

} // namespace box2d
