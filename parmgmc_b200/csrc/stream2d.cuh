// stream2d.cuh -- fused, warp-streaming red-black Gibbs sweep for the matrix-free 2D 5-point operator.
//
// One pass over memory does what the reference does in up to seven (VecSetRandomStandardNormal, VecPointwiseMult,
// VecAXPY, the two colour phases of MCSORApply, MatResidual, MatRestrict / MatInterpolateAdd):
//
//   guess   x_old = xin                      (LOAD)      | 0 (ZERO, PCMG zeroes the coarse iterate)
//                 | xin + P xc               (PROLONG, MatInterpolateAdd fused into the post-smoother)
//   sweep   red nodes of row j, then black nodes of row j-1 (all their red neighbours are then new): the forward
//           red-black sweep of src/mc_sor.c:257-271 on w = b + sqrtdiag z (src/pc_mcgibbs.c:124-126), z drawn on the fly
//   post    optionally r = b - A x_new and b_c = P^T r (MatResidual + MatRestrict) without ever storing r
//
// Work decomposition: a warp owns a strip of 128 columns (lane l owns columns c0+4l .. c0+4l+3, of which lanes 1..30 are
// written back; lanes 0 and 31 recompute the neighbouring strips' edge columns) and a band of `by` rows, and walks down
// the band keeping a rolling window of rows in registers.  East/west neighbours come from warp shuffles, north/south
// neighbours are the thread's own registers: no shared memory, no block barrier, every global access is a coalesced
// row segment.  The result is written out of place, so strips and bands are independent; band and strip edges are
// recomputed redundantly (halo of 1 row/side for a plain sweep, 3 rows/side with the fused residual+restriction).
// Arithmetic per node is exactly that of the per-colour kernels (same fma order), so the result is bit-identical.
#pragma once
#include "common.hpp"
#include "fastnormal.cuh"
#include "philox.cuh"

namespace stream2d {



struct Geom2 {
  int nx, ny;   // global grid
  int slo, shi; // owned rows
};

struct LapTab2 {
  double diag[5], idiag[5], sqrtdiag[5]; // by number of existing neighbours
  double h, omo;
};

enum Guess { GUESS_LOAD = 0, GUESS_ZERO = 1, GUESS_PROLONG = 2 };

// one warp's work: output columns of strip `strip`, output rows [ja, jb)
struct Item {
  int strip, ja, jb;
};

struct Args {
  Geom2         g, gc; // fine grid; coarse grid (PROLONG / RESTRICT)
  const Item   *items; // work list (built on the host: LapOp::stream_items); nitems warps are launched
  int           nitems;
  int           pitch; // row stride of xin / xout / b: nx rounded up to 4, so that a lane's four columns are one aligned
                       // 32-byte access (LDG.256 / STG.256) and a warp row is one contiguous kilobyte
  int           by, nstrips, nbands;
  int           cpitch; // row stride of xc / bc: the coarse row length, or the coarse level's pitch when its vectors are pitched
  int           flip; // 0: forward sweep (colour (i+j) even first); 1: backward sweep (colour (i+j) odd first)
  const double *xin, *b, *xc;
  double       *xout, *bc;
  LapTab2       tab;
  NoiseArgs     na;
};

constexpr int STRIP_OUT = 120; // columns written per warp

// pull the 32 bytes at p (and hence its 128-byte line) into L1 ahead of use
__device__ __forceinline__ void prefetch_l1(const double *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

__device__ __forceinline__ double shfl_up1(double v) { return __shfl_up_sync(0xffffffffu, v, 1); }
__device__ __forceinline__ double shfl_dn1(double v) { return __shfl_down_sync(0xffffffffu, v, 1); }

__device__ __forceinline__ void ld256(const double *p, double (&out)[4]) { asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(out[0]), "=d"(out[1]), "=d"(out[2]), "=d"(out[3]) : "l"(p)); }
__device__ __forceinline__ void st256(double *p, const double (&v)[4]) { asm volatile("st.global.v4.f64 [%4], {%0,%1,%2,%3};" ::"d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]), "l"(p) : "memory"); }

// four consecutive entries of row j starting at column c (zero where the node does not exist / is not owned); `stride`
// is the row stride of v.  INTERIOR: the caller guarantees that all four exist; ALIGNED: stride and c are multiples of 4
// and v is 32-byte aligned (the pitched fine-level vectors).
template <bool INTERIOR = false, bool ALIGNED = false>
__device__ __forceinline__ void load4(const double *__restrict__ v, const Geom2 &g, int stride, int j, int c, double (&out)[4])
{
  if (INTERIOR) {
    if (v == nullptr) {
      out[0] = out[1] = out[2] = out[3] = 0.0;
      return;
    }
    const double *p = v + (size_t)(j - g.slo) * stride + c;
    if (ALIGNED) ld256(p, out);
    else {
#pragma unroll
      for (int m = 0; m < 4; ++m) out[m] = p[m];
    }
    return;
  }
  const bool rowok = v != nullptr && j >= g.slo && j < g.shi;
  const double *p  = v + (size_t)(j - g.slo) * stride + c;
#pragma unroll
  for (int m = 0; m < 4; ++m) out[m] = (rowok && c + m >= 0 && c + m < g.nx) ? p[m] : 0.0;
}

// the four normals of generator indices g0 .. g0+3, g0 = 0 mod 4 (padded indices, philox.cuh): one Philox call per lane
__device__ __forceinline__ void philox_normals4(const fastnormal::Tables &ft, const NoiseArgs &na, long long g0, double (&z)[4])
{
  const long long quad = g0 >> 2; // may be negative for halo lanes outside the grid: those values are never used
  uint32_t        w0, w1, w2, w3;
  philox4x32_10((uint32_t)quad, (uint32_t)((uint64_t)quad >> 32), (uint32_t)na.call, (uint32_t)(na.call >> 32), (uint32_t)na.seed, (uint32_t)(na.seed >> 32), w0, w1, w2, w3);
  fastnormal::box_muller(ft, w0, w1, z[0], z[1]);
  fastnormal::box_muller(ft, w2, w3, z[2], z[3]);
}

// the four normals of columns c..c+3 of row j (global rows g0 = c + nx j ...), one Philox quad per lane plus shuffles
template <bool INTERIOR = false>
__device__ __forceinline__ void noise4(const fastnormal::Tables &ft, const NoiseArgs &na, const Geom2 &g, int pitch, int j, int c, double (&z)[4])
{
  if (na.mode == PMG_NOISE_NONE) {
    z[0] = z[1] = z[2] = z[3] = 0.0;
    return;
  }
  if (na.mode == PMG_NOISE_INJECTED) { // the tape is the caller's: natural row stride, no alignment
    load4<INTERIOR, false>(na.tape, g, g.nx, j, c, z);
    return;
  }
  philox_normals4(ft, na, (long long)j * pitch + c, z);
}

// one node update (src/mc_sor.c:260-268): column M of `row` (row index j), south/north rows, west/east values for M = 0 / 3
template <int M, bool INTERIOR = false>
__device__ __forceinline__ void update(const Geom2 &g, const LapTab2 &t, int j, int c, double (&row)[4], const double (&south)[4], const double (&north)[4], double west, double east, double bval, double z, bool noisy)
{
  if (INTERIOR) { // every neighbour exists
    const double xw = M == 0 ? west : row[M == 0 ? 0 : M - 1];
    const double xe = M == 3 ? east : row[M == 3 ? 3 : M + 1];
    double       sum = noisy ? __dadd_rn(__dmul_rn(z, t.sqrtdiag[4]), bval) : bval;
    sum = fma(t.h, south[M], sum);
    sum = fma(t.h, xw, sum);
    sum = fma(t.h, xe, sum);
    sum = fma(t.h, north[M], sum);
    const double t0 = __dmul_rn(t.omo, row[M]);
    row[M]          = fma(t.idiag[4], sum, t0);
    return;
  }
  const int i = c + M;
  if (i < 0 || i >= g.nx || j < 0 || j >= g.ny) return;
  const bool   W = i > 0, E = i < g.nx - 1, S = j > 0, N = j < g.ny - 1;
  const int    deg = (int)W + (int)E + (int)S + (int)N;
  const double xw = M == 0 ? west : row[M == 0 ? 0 : M - 1];
  const double xe = M == 3 ? east : row[M == 3 ? 3 : M + 1];
  double       sum = noisy ? __dadd_rn(__dmul_rn(z, t.sqrtdiag[deg]), bval) : bval;
  if (S) sum = fma(t.h, south[M], sum);
  if (W) sum = fma(t.h, xw, sum);
  if (E) sum = fma(t.h, xe, sum);
  if (N) sum = fma(t.h, north[M], sum);
  const double t0 = __dmul_rn(t.omo, row[M]);
  row[M]          = fma(t.idiag[deg], sum, t0);
}

// residual of column M of row j: r = b - A x with the assembled row's accumulation order (S, W, C, E, N)
template <int M, bool INTERIOR = false>
__device__ __forceinline__ double resid(const Geom2 &g, const LapTab2 &t, int j, int c, const double (&row)[4], const double (&south)[4], const double (&north)[4], double west, double east, double bval)
{
  if (INTERIOR) {
    const double xw = M == 0 ? west : row[M == 0 ? 0 : M - 1];
    const double xe = M == 3 ? east : row[M == 3 ? 3 : M + 1];
    const double mh = -t.h;
    double       ax = 0.0;
    ax = fma(mh, south[M], ax);
    ax = fma(mh, xw, ax);
    ax = fma(t.diag[4], row[M], ax);
    ax = fma(mh, xe, ax);
    ax = fma(mh, north[M], ax);
    return __dsub_rn(bval, ax);
  }
  const int i = c + M;
  if (i < 0 || i >= g.nx || j < 0 || j >= g.ny) return 0.0;
  const bool   W = i > 0, E = i < g.nx - 1, S = j > 0, N = j < g.ny - 1;
  const int    deg = (int)W + (int)E + (int)S + (int)N;
  const double xw = M == 0 ? west : row[M == 0 ? 0 : M - 1];
  const double xe = M == 3 ? east : row[M == 3 ? 3 : M + 1];
  const double mh = -t.h;
  double       ax = 0.0;
  if (S) ax = fma(mh, south[M], ax);
  if (W) ax = fma(mh, xw, ax);
  ax = fma(t.diag[deg], row[M], ax);
  if (E) ax = fma(mh, xe, ax);
  if (N) ax = fma(mh, north[M], ax);
  return __dsub_rn(bval, ax);
}

__device__ __forceinline__ void copy4(double (&d)[4], const double (&s)[4])
{
#pragma unroll
  for (int m = 0; m < 4; ++m) d[m] = s[m];
}

// x_old row j = xin row j (+ P xc): MatInterpolateAdd's order, s = x; s = fma(w, xc_J, s) over ascending coarse index
template <int GUESS, bool INTERIOR = false>
__device__ __forceinline__ void guess_row(const Args &a, int j, int c, double (&out)[4])
{
  if (GUESS == GUESS_ZERO) {
    out[0] = out[1] = out[2] = out[3] = 0.0;
    return;
  }
  load4<INTERIOR, true>(a.xin, a.g, a.pitch, j, c, out);
  if (GUESS == GUESS_PROLONG && INTERIOR) { // all parents exist
    const int    J0 = j >> 1, nJ = (j & 1) ? 2 : 1;
    const double wj = (j & 1) ? 0.5 : 1.0, wh = 0.5 * wj;
    const double *p = a.xc + (size_t)(J0 - a.gc.slo) * a.cpitch + (c >> 1);
    for (int q = 0; q < nJ; ++q, p += a.cpitch) {
      const double c0 = p[0], c1 = p[1], c2 = p[2];
      out[0] = fma(wj, c0, out[0]);
      out[1] = fma(wh, c1, fma(wh, c0, out[1]));
      out[2] = fma(wj, c1, out[2]);
      out[3] = fma(wh, c2, fma(wh, c1, out[3]));
    }
    return;
  }
  if (GUESS == GUESS_PROLONG) {
    if (j < 0 || j >= a.g.ny) return;
    const int    J0 = j >> 1, nJ = (j & 1) ? 2 : 1;
    const double wj = (j & 1) ? 0.5 : 1.0;
    const int    I0 = c >> 1; // c = 0 mod 4; fine columns c..c+3 see coarse columns I0, I0+1, I0+2
    for (int q = 0; q < nJ; ++q) {
      const int J = J0 + q;
      if (J >= a.gc.ny) continue;
      double       cv[3];
      const double *p = a.xc + (size_t)(J - a.gc.slo) * a.cpitch + I0;
#pragma unroll
      for (int m = 0; m < 3; ++m) cv[m] = (I0 + m >= 0 && I0 + m < a.gc.nx) ? p[m] : 0.0;
      // column c (even): parent I0; c+1 (odd): I0, I0+1; c+2 (even): I0+1; c+3 (odd): I0+1, I0+2
      if (c >= 0 && c < a.g.nx) out[0] = fma(wj, cv[0], out[0]);
      if (c + 1 >= 0 && c + 1 < a.g.nx) {
        out[1] = fma(0.5 * wj, cv[0], out[1]);
        if (I0 + 1 < a.gc.nx) out[1] = fma(0.5 * wj, cv[1], out[1]);
      }
      if (c + 2 >= 0 && c + 2 < a.g.nx) out[2] = fma(wj, cv[1], out[2]);
      if (c + 3 >= 0 && c + 3 < a.g.nx) {
        out[3] = fma(0.5 * wj, cv[1], out[3]);
        if (I0 + 2 < a.gc.nx) out[3] = fma(0.5 * wj, cv[2], out[3]);
      }
    }
  }
}

// ---- the streaming loop of one warp --------------------------------------------------------------------------
// INTERIOR: every node this warp touches has all four neighbours (about 90% of the warps of a large grid), so no
// existence predicates are evaluated.  The window of rows is advanced by register copies: a fully unrolled ring of
// slots removes those ~80 moves per row but the larger loop body then misses in the instruction cache and runs slower
// (profiles/stream2d_notes.md), so the compact loop is kept.
template <int GUESS, bool RESTRICT, bool INTERIOR>
__device__ __forceinline__ void run_warp(const Args &a, const fastnormal::Tables &ft, int lane, int c, int ja, int jb)
{
  const Geom2 &g = a.g;
  // phase A (first colour) rows jA0..jA1; phase B (second colour) rows jA0+1..jA1-1 are exact
  const int  jA0 = ja - (RESTRICT ? 3 : 1), jA1 = jb + (RESTRICT ? 2 : 0);
  const bool out_lane = lane >= 1 && lane <= 30;
  const bool noisy    = a.na.mode != PMG_NOISE_NONE;
  constexpr int PF = 3; // L1 prefetch distance in rows (interior warps only)

  double xm3[4] = {0, 0, 0, 0}, xm2[4] = {0, 0, 0, 0}, xm1[4], x0[4]; // rows jj-3 .. jj
  double bm1[4] = {0, 0, 0, 0}, bm2[4] = {0, 0, 0, 0};                 // rhs rows jj-1, jj-2 (RESTRICT: the residual needs whole rows)
  double bk[2] = {0, 0}, zk[2] = {0, 0};                               // rhs / noise of row jj-1 at the two second-colour columns phase B updates
  double r1[4] = {0, 0, 0, 0}, r2[4] = {0, 0, 0, 0};                   // residual rows jj-3, jj-4
  double r1w = 0, r2w = 0;                                             // their column c-1 (from the lane to the west)

  guess_row<GUESS, INTERIOR>(a, jA0 - 1, c, xm1);
  guess_row<GUESS, INTERIOR>(a, jA0, c, x0);

  for (int jj = jA0; jj <= jA1; ++jj) {
    // this iteration's incoming rows: requested first, consumed after the normals have been computed, so that the
    // generator hides their latency; keeping the window this short is what lets 20 warps stay resident per SM
    double xp1[4], b0[4];
    guess_row<GUESS, INTERIOR>(a, jj + 1, c, xp1);
    load4<INTERIOR, true>(a.b, g, a.pitch, jj, c, b0);
    if (INTERIOR) { // rows further ahead go to L2 / L1 now
      if (GUESS != GUESS_ZERO) prefetch_l1(a.xin + (size_t)(jj + 1 + PF - g.slo) * a.pitch + c);
      if (a.b) prefetch_l1(a.b + (size_t)(jj + PF - g.slo) * a.pitch + c);
    }
    double z[4];
    noise4<INTERIOR>(ft, a.na, g, a.pitch, jj, c, z);

    const bool even = ((jj + a.flip) & 1) == 0; // first-colour columns of row jj are M = 0,2 (else 1,3)
    { // ---- phase A: first-colour nodes of row jj ----
      const double west = shfl_up1(x0[3]), east = shfl_dn1(x0[0]);
      if (even) {
        update<0, INTERIOR>(g, a.tab, jj, c, x0, xm1, xp1, west, east, b0[0], z[0], noisy);
        update<2, INTERIOR>(g, a.tab, jj, c, x0, xm1, xp1, west, east, b0[2], z[2], noisy);
      } else {
        update<1, INTERIOR>(g, a.tab, jj, c, x0, xm1, xp1, west, east, b0[1], z[1], noisy);
        update<3, INTERIOR>(g, a.tab, jj, c, x0, xm1, xp1, west, east, b0[3], z[3], noisy);
      }
    }
    { // ---- phase B: second-colour nodes of row jj-1 (same columns); all their neighbours are new ----
      const double west = shfl_up1(xm1[3]), east = shfl_dn1(xm1[0]);
      if (even) {
        update<0, INTERIOR>(g, a.tab, jj - 1, c, xm1, xm2, x0, west, east, bk[0], zk[0], noisy);
        update<2, INTERIOR>(g, a.tab, jj - 1, c, xm1, xm2, x0, west, east, bk[1], zk[1], noisy);
      } else {
        update<1, INTERIOR>(g, a.tab, jj - 1, c, xm1, xm2, x0, west, east, bk[0], zk[0], noisy);
        update<3, INTERIOR>(g, a.tab, jj - 1, c, xm1, xm2, x0, west, east, bk[1], zk[1], noisy);
      }
    }
    // what phase B of the next iteration (whose columns are the other two) needs of this row
    bk[0] = even ? b0[1] : b0[0]; bk[1] = even ? b0[3] : b0[2];
    zk[0] = even ? z[1] : z[0];   zk[1] = even ? z[3] : z[2];
    const int jo = jj - 1; // row jj-1 is final
    if (out_lane && jo >= ja && jo < jb) {
      double *p = a.xout + (size_t)(jo - g.slo) * a.pitch + c;
      if (INTERIOR) st256(p, xm1);
      else {
#pragma unroll
        for (int m = 0; m < 4; ++m)
          if (c + m < g.nx) p[m] = xm1[m];
      }
    }
    if (RESTRICT) {
      const int    jr = jj - 2; // residual of row jj-2 (rows jj-3, jj-2, jj-1 are final)
      const double west = shfl_up1(xm2[3]), east = shfl_dn1(xm2[0]);
      double       r0[4];
      r0[0] = resid<0, INTERIOR>(g, a.tab, jr, c, xm2, xm3, xm1, west, east, bm2[0]);
      r0[1] = resid<1, INTERIOR>(g, a.tab, jr, c, xm2, xm3, xm1, west, east, bm2[1]);
      r0[2] = resid<2, INTERIOR>(g, a.tab, jr, c, xm2, xm3, xm1, west, east, bm2[2]);
      r0[3] = resid<3, INTERIOR>(g, a.tab, jr, c, xm2, xm3, xm1, west, east, bm2[3]);
      const double r0w = shfl_up1(r0[3]);
      const int    jc  = jr - 1; // centre row 2J of the coarse row that is now complete
      if ((jc & 1) == 0 && jc >= ja && jc < jb && out_lane) {
        const int  J = jc >> 1, I0 = c >> 1;
        const bool hasS = INTERIOR || jc - 1 >= 0, hasN = INTERIOR || jc + 1 < g.ny;
#pragma unroll
        for (int q = 0; q < 2; ++q) { // coarse columns I0 (fine c) and I0+1 (fine c+2)
          const int I = I0 + q, fc = c + 2 * q;
          if (!INTERIOR && (I >= a.gc.nx || fc >= g.nx)) continue;
          const bool   hasW = INTERIOR || fc - 1 >= 0, hasE = INTERIOR || fc + 1 < g.nx;
          const double sW = q == 0 ? r2w : r2[1], sC = q == 0 ? r2[0] : r2[2], sE = q == 0 ? r2[1] : r2[3];
          const double cW = q == 0 ? r1w : r1[1], cC = q == 0 ? r1[0] : r1[2], cE = q == 0 ? r1[1] : r1[3];
          const double nW = q == 0 ? r0w : r0[1], nC = q == 0 ? r0[0] : r0[2], nE = q == 0 ? r0[1] : r0[3];
          double       acc = 0.0; // ascending fine index = MatMultTranspose's order
          if (hasS) {
            if (hasW) acc = fma(0.25, sW, acc);
            acc = fma(0.5, sC, acc);
            if (hasE) acc = fma(0.25, sE, acc);
          }
          if (hasW) acc = fma(0.5, cW, acc);
          acc = fma(1.0, cC, acc);
          if (hasE) acc = fma(0.5, cE, acc);
          if (hasN) {
            if (hasW) acc = fma(0.25, nW, acc);
            acc = fma(0.5, nC, acc);
            if (hasE) acc = fma(0.25, nE, acc);
          }
          a.bc[(size_t)(J - a.gc.slo) * a.cpitch + I] = acc;
        }
      }
      copy4(r2, r1); r2w = r1w;
      copy4(r1, r0); r1w = r0w;
    }
    // advance the window
    if (RESTRICT) {
      copy4(xm3, xm2);
      copy4(bm2, bm1); copy4(bm1, b0);
    }
    copy4(xm2, xm1); copy4(xm1, x0); copy4(x0, xp1);
  }
}

template <int GUESS, bool RESTRICT> __global__ void __launch_bounds__(128, RESTRICT ? 3 : 5) lap_stream_kernel(const Args a)
{
  __shared__ fastnormal::SharedTables fts;
  const fastnormal::Tables ft = fastnormal::load_tables(fts);
  __syncthreads();
  const Geom2 &g    = a.g;
  const int    lane = threadIdx.x & 31;
  const int    w    = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (w >= a.nitems) return;
  const Item it = a.items[w];
  const int  c0 = it.strip * STRIP_OUT - 4;
  const int  ja = it.ja, jb = it.jb;
  const int jlo = ja - (RESTRICT ? 3 : 1) - 1, jhi = jb + (RESTRICT ? 2 : 0) + 2; // first / last row touched
  const bool interior = c0 >= 1 && c0 + 127 <= g.nx - 2 && jlo >= 1 && jhi <= g.ny - 2 && jlo >= g.slo && jhi + 4 < g.shi;
  if (interior) run_warp<GUESS, RESTRICT, true>(a, ft, lane, c0 + 4 * lane, ja, jb);
  else run_warp<GUESS, RESTRICT, false>(a, ft, lane, c0 + 4 * lane, ja, jb);
}

} // namespace stream2d
