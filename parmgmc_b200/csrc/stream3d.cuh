// stream3d.cuh -- fused red-black Gibbs sweep for the matrix-free 3D 7-point operator: both colours, the noise and the
// right-hand-side perturbation in ONE pass over memory (the reference needs VecSetRandomStandardNormal, VecPointwiseMult,
// VecAXPY and the two colour phases of MCSORApply: src/pc_mcgibbs.c:119-128, src/mc_sor.c:257-271).
//
// 2.5D blocking.  A CTA of NW warps owns a tile of 120 columns (x) by NW-2 rows (y) and walks along z through a band of
// planes.  A warp is one grid row of the tile (lane l owns columns c0+4l .. c0+4l+3, lanes 0 / 31 are halo columns, as
// in stream2d.cuh); warps 0 and NW-1 are halo rows.  Along z a thread keeps a rolling window of four planes in registers;
// east / west neighbours come from warp shuffles; north / south neighbours come from the neighbouring warps through two
// small double-buffered shared-memory rings (the old plane k+1 and the half-updated plane k), one __syncthreads per plane.
// At plane step k the first colour of plane k is updated (all its neighbours are still old), then the second colour of
// plane k-1 (all its neighbours are new by then), and plane k-1 is written out of place: 8 B read of x, 8 B read of b and
// 8 B written per DOF-update instead of the 48 B a colour-by-colour sweep moves.  Halo columns, halo rows and one halo
// plane per band end are recomputed redundantly; the arithmetic per node is exactly lap_sweep_kernel<3>'s (stencil_op.cu),
// fma for fma, so the result is bit-identical to the colour-by-colour path.
#pragma once
#include "common.hpp"
#include "fastnormal.cuh"
#include "philox.cuh"
#include "stream2d.cuh"

namespace stream3d {

using stream2d::ld256;
using stream2d::shfl_dn1;
using stream2d::shfl_up1;
using stream2d::st256;

struct Geom3 {
  int nx, ny, nz;
  int slo, shi; // owned planes
};

struct LapTab3 {
  double diag[7], idiag[7], sqrtdiag[7]; // by number of existing neighbours
  double h, omo;
};

// one CTA's work: output columns of strip `strip`, output rows [ya, ya + NW - 2), output planes [ka, kb)
struct Item {
  int strip, ya, ka, kb;
};

struct Args {
  Geom3         g;
  int           pitch;  // row stride of xin / xout / b (nx rounded up to 4)
  long long     pplane; // plane stride = pitch * ny
  const Item   *items;
  int           flip; // 0: forward sweep (colour (i+j+k) even first); 1: backward
  const double *xin, *b;
  double       *xout;
  LapTab3       tab;
  NoiseArgs     na;
};

constexpr int STRIP_OUT = 120;

// four consecutive entries of row (y, k) from column c; zero where the node does not exist / is not owned
template <bool INTERIOR, bool ALIGNED>
__device__ __forceinline__ void load4(const double *__restrict__ v, const Geom3 &g, int stride, long long pstride, int y, int k, int c, double (&out)[4])
{
  if (v == nullptr) {
    out[0] = out[1] = out[2] = out[3] = 0.0;
    return;
  }
  const double *p = v + (long long)(k - g.slo) * pstride + (long long)y * stride + c;
  if (INTERIOR) {
    if (ALIGNED) ld256(p, out);
    else {
#pragma unroll
      for (int m = 0; m < 4; ++m) out[m] = p[m];
    }
    return;
  }
  const bool ok = y >= 0 && y < g.ny && k >= g.slo && k < g.shi;
#pragma unroll
  for (int m = 0; m < 4; ++m) out[m] = (ok && c + m >= 0 && c + m < g.nx) ? p[m] : 0.0;
}

template <bool INTERIOR>
__device__ __forceinline__ void noise4(const fastnormal::Tables &ft, const NoiseArgs &na, const Geom3 &g, int pitch, int y, int k, int c, double (&z)[4])
{
  if (na.mode == PMG_NOISE_NONE) {
    z[0] = z[1] = z[2] = z[3] = 0.0;
    return;
  }
  if (na.mode == PMG_NOISE_INJECTED) { // the caller's tape: natural strides, local planes
    load4<INTERIOR, false>(na.tape, g, g.nx, (long long)g.nx * g.ny, y, k, c, z);
    return;
  }
  stream2d::philox_normals4(ft, na, ((long long)k * g.ny + y) * pitch + c, z);
}

// one node update, column M of `row` = row (y, k); accumulation order of the assembled row: down, south, west, east, north, up
template <int M, bool INTERIOR>
__device__ __forceinline__ void update(const Geom3 &g, const LapTab3 &t, int y, int k, int c, double (&row)[4], const double (&down)[4], const double (&south)[4], const double (&north)[4], const double (&up)[4], double west,
                                       double east, double bval, double z, bool noisy)
{
  const double xw = M == 0 ? west : row[M == 0 ? 0 : M - 1];
  const double xe = M == 3 ? east : row[M == 3 ? 3 : M + 1];
  if (INTERIOR) {
    double sum = noisy ? __dadd_rn(__dmul_rn(z, t.sqrtdiag[6]), bval) : bval;
    sum = fma(t.h, down[M], sum);
    sum = fma(t.h, south[M], sum);
    sum = fma(t.h, xw, sum);
    sum = fma(t.h, xe, sum);
    sum = fma(t.h, north[M], sum);
    sum = fma(t.h, up[M], sum);
    const double t0 = __dmul_rn(t.omo, row[M]);
    row[M]          = fma(t.idiag[6], sum, t0);
    return;
  }
  const int i = c + M;
  if (i < 0 || i >= g.nx || y < 0 || y >= g.ny || k < 0 || k >= g.nz) return;
  const bool W = i > 0, E = i < g.nx - 1, S = y > 0, N = y < g.ny - 1, D = k > 0, U = k < g.nz - 1;
  const int  deg = (int)W + (int)E + (int)S + (int)N + (int)D + (int)U;
  double     sum = noisy ? __dadd_rn(__dmul_rn(z, t.sqrtdiag[deg]), bval) : bval;
  if (D) sum = fma(t.h, down[M], sum);
  if (S) sum = fma(t.h, south[M], sum);
  if (W) sum = fma(t.h, xw, sum);
  if (E) sum = fma(t.h, xe, sum);
  if (N) sum = fma(t.h, north[M], sum);
  if (U) sum = fma(t.h, up[M], sum);
  const double t0 = __dmul_rn(t.omo, row[M]);
  row[M]          = fma(t.idiag[deg], sum, t0);
}

struct __align__(32) Row4 {
  double v[4];
};

template <int NW, bool INTERIOR>
__device__ __forceinline__ void run_cta(const Args &a, const fastnormal::Tables &ft, Row4 (*sm_old)[NW][32], Row4 (*sm_new)[NW][32], const Item it)
{
  const Geom3 &g    = a.g;
  const int    lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int    c    = it.strip * STRIP_OUT - 4 + 4 * lane;
  const int    y    = it.ya - 1 + w;
  const bool   halo_row = w == 0 || w == NW - 1;
  const bool   out_lane = lane >= 1 && lane <= 30;
  const bool   noisy    = a.na.mode != PMG_NOISE_NONE;
  const int    kA0 = it.ka - 1, kA1 = it.kb; // phase A planes; phase B planes kA0+1 .. kA1-1 are exact
  // the outer y-neighbour of a halo row lives in another CTA's tile: its OLD value is read from xin (the sweep is out of place)
  const int yo = w == 0 ? y - 1 : y + 1;

  double xm2[4] = {0, 0, 0, 0}, xm1[4], x0[4]; // planes kk-2 (final), kk-1 (first colour done), kk
  double bk[2] = {0, 0}, zk[2] = {0, 0};       // rhs / noise of plane kk-1 at the two second-colour columns phase B updates

  load4<INTERIOR, true>(a.xin, g, a.pitch, a.pplane, y, kA0 - 1, c, xm1);
  load4<INTERIOR, true>(a.xin, g, a.pitch, a.pplane, y, kA0, c, x0);
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    sm_new[(kA0 - 1) & 1][w][lane].v[m] = xm1[m];
    sm_old[kA0 & 1][w][lane].v[m]       = x0[m];
  }
  __syncthreads();

  for (int kk = kA0; kk <= kA1; ++kk) {
    double xp1[4], b0[4], xo[4] = {0, 0, 0, 0};
    load4<INTERIOR, true>(a.xin, g, a.pitch, a.pplane, y, kk + 1, c, xp1);
    load4<INTERIOR, true>(a.b, g, a.pitch, a.pplane, y, kk, c, b0);
    if (halo_row) load4<INTERIOR, true>(a.xin, g, a.pitch, a.pplane, yo, kk, c, xo);
    double z[4];
    noise4<INTERIOR>(ft, a.na, g, a.pitch, y, kk, c, z);

    const bool even = ((y + kk + a.flip) & 1) == 0; // first-colour columns of row (y, kk) are M = 0,2 (else 1,3)
    { // ---- phase A: first-colour nodes of plane kk; every neighbour is still old ----
      double south[4], north[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        south[m] = w == 0 ? xo[m] : sm_old[kk & 1][w == 0 ? 0 : w - 1][lane].v[m];
        north[m] = w == NW - 1 ? xo[m] : sm_old[kk & 1][w == NW - 1 ? w : w + 1][lane].v[m];
      }
      const double west = shfl_up1(x0[3]), east = shfl_dn1(x0[0]);
      if (even) {
        update<0, INTERIOR>(g, a.tab, y, kk, c, x0, xm1, south, north, xp1, west, east, b0[0], z[0], noisy);
        update<2, INTERIOR>(g, a.tab, y, kk, c, x0, xm1, south, north, xp1, west, east, b0[2], z[2], noisy);
      } else {
        update<1, INTERIOR>(g, a.tab, y, kk, c, x0, xm1, south, north, xp1, west, east, b0[1], z[1], noisy);
        update<3, INTERIOR>(g, a.tab, y, kk, c, x0, xm1, south, north, xp1, west, east, b0[3], z[3], noisy);
      }
    }
    if (!halo_row) { // ---- phase B: second-colour nodes of plane kk-1 (same columns); all their neighbours are new ----
      double south[4], north[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        south[m] = sm_new[(kk - 1) & 1][w - 1][lane].v[m];
        north[m] = sm_new[(kk - 1) & 1][w + 1][lane].v[m];
      }
      const double west = shfl_up1(xm1[3]), east = shfl_dn1(xm1[0]);
      if (even) {
        update<0, INTERIOR>(g, a.tab, y, kk - 1, c, xm1, xm2, south, north, x0, west, east, bk[0], zk[0], noisy);
        update<2, INTERIOR>(g, a.tab, y, kk - 1, c, xm1, xm2, south, north, x0, west, east, bk[1], zk[1], noisy);
      } else {
        update<1, INTERIOR>(g, a.tab, y, kk - 1, c, xm1, xm2, south, north, x0, west, east, bk[0], zk[0], noisy);
        update<3, INTERIOR>(g, a.tab, y, kk - 1, c, xm1, xm2, south, north, x0, west, east, bk[1], zk[1], noisy);
      }
      const int ko = kk - 1; // plane kk-1 of this row is final
      if (out_lane && ko >= it.ka && ko < it.kb && (INTERIOR || (y >= 0 && y < g.ny))) {
        double *p = a.xout + (long long)(ko - g.slo) * a.pplane + (long long)y * a.pitch + c;
        if (INTERIOR) st256(p, xm1);
        else {
#pragma unroll
          for (int m = 0; m < 4; ++m)
            if (c + m < g.nx) p[m] = xm1[m];
        }
      }
    }
    bk[0] = even ? b0[1] : b0[0]; bk[1] = even ? b0[3] : b0[2];
    zk[0] = even ? z[1] : z[0];   zk[1] = even ? z[3] : z[2];
    // publish the half-updated plane kk and the old plane kk+1 for the neighbouring rows
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      sm_new[kk & 1][w][lane].v[m]       = x0[m];
      sm_old[(kk + 1) & 1][w][lane].v[m] = xp1[m];
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      xm2[m] = xm1[m];
      xm1[m] = x0[m];
      x0[m]  = xp1[m];
    }
  }
}

template <int NW> __global__ void __launch_bounds__(NW * 32, NW <= 10 ? 2 : 1) lap_stream3d_kernel(const Args a)
{
  extern __shared__ __align__(32) unsigned char smem_raw[];
  fastnormal::SharedTables *fts = reinterpret_cast<fastnormal::SharedTables *>(smem_raw);
  Row4(*sm_old)[NW][32]         = reinterpret_cast<Row4(*)[NW][32]>(smem_raw + sizeof(fastnormal::SharedTables));
  Row4(*sm_new)[NW][32]         = sm_old + 2;
  const fastnormal::Tables ft   = fastnormal::load_tables(*fts);
  __syncthreads();
  const Geom3 &g  = a.g;
  const Item   it = a.items[blockIdx.x];
  const int    c0 = it.strip * STRIP_OUT - 4;
  // every node the CTA computes has all six neighbours, and every row / plane it reads exists and is owned
  const bool interior = c0 >= 1 && c0 + 127 <= g.nx - 2 && it.ya - 2 >= 0 && it.ya + NW - 2 <= g.ny - 2 && it.ka - 2 >= 1 && it.kb + 1 <= g.nz - 2 && it.ka - 2 >= g.slo && it.kb + 1 < g.shi;
  if (interior) run_cta<NW, true>(a, ft, sm_old, sm_new, it);
  else run_cta<NW, false>(a, ft, sm_old, sm_new, it);
}

template <int NW> constexpr size_t smem_bytes() { return sizeof(fastnormal::SharedTables) + 4 * NW * 32 * sizeof(Row4); }

} // namespace stream3d
