// resrest3d.cuh -- b_c = P^T (b - A x) for the finest level of the 3D V-cycle in ONE pass over x and b (MatResidual +
// MatRestrict of PCMG's down-sweep, src/pc_gamgmc.c:242-259 with the Q1 restriction of SURVEY Appendix A.4): the fine
// residual never goes to memory.  The two-kernel form (lap_residual3_pitched_kernel + restrict3_pitched_kernel) moves
// 24 + 9 B per fine node, this one 16 + 1.
//
// A CTA owns TI x TJ coarse nodes of the (I, J) plane and walks up a segment of coarse planes.  Per coarse plane K it
// computes the residual of the fine planes 2K and 2K+1 of its footprint (fine columns 2 I0 - 4 .. 2 I0 + 2 TI + 3, fine rows
// 2 J0 - 1 .. 2 J0 + 2 TJ - 1) into a five-slot ring of planes in shared memory -- one task = four columns of one row, the
// loads and the fma sequence of lap_residual3_pitched_kernel -- and, after one barrier, restricts the planes 2K-1, 2K,
// 2K+1 with restrict3_pitched_kernel's accumulation order (four coarse nodes per thread).  Plane 2K-1 is the previous
// step's 2(K-1)+1; a segment starts by computing it once.  x is read from global memory five times per node (centre, north /
// south, up / down), four of them L1 hits: the three fine planes a step touches are 60 KB of a footprint.
// The result is bit-identical to the two kernels it replaces (tests/test_gpu_parity.py).
#pragma once
#include "common.hpp"

namespace resrest3d {

constexpr int TI = 60, TJ = 8, NT = 384, RING = 5;
constexpr int W = 2 * TI + 8;      // fine columns of a footprint row (32 quads)
constexpr int R = 2 * TJ + 1;      // fine rows of a footprint
constexpr int PLANE = W * R;       // doubles per ring slot
constexpr size_t SMEM = (size_t)RING * PLANE * sizeof(double);

struct Args {
  int           n0, n1, n2;   // fine grid
  int           ld;           // fine row stride (pitch, a multiple of 4)
  long long     unit;         // fine plane stride = ld * n1
  int           c0, c1, c2;   // coarse grid
  int           cld;          // coarse row stride
  int           tiles_i, tiles_j, kseg; // tiles per coarse plane, coarse planes per segment
  double        mh;           // -h
  double        diag[7];      // by number of existing neighbours
  const double *x, *b;
  double       *bc;
};

__device__ __forceinline__ void ldg256(const double *p, double (&v)[4]) { asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p)); }

// residual of the fine columns i0 .. i0+3 of row j, plane k (lap_residual3_pitched_kernel's arithmetic, fma for fma)
__device__ __forceinline__ void residual_quad(const Args &a, int i0, int j, int k, double (&r)[4])
{
  const long long idx = i0 + (long long)a.ld * j + a.unit * k;
  const bool      S = j > 0, N = j < a.n1 - 1, D = k > 0, U = k < a.n2 - 1;
  double          xc[4], xs[4] = {0, 0, 0, 0}, xn[4] = {0, 0, 0, 0}, xd[4] = {0, 0, 0, 0}, xu[4] = {0, 0, 0, 0}, bb[4];
  ldg256(a.x + idx, xc);
  ldg256(a.b + idx, bb);
  if (S) ldg256(a.x + idx - a.ld, xs);
  if (N) ldg256(a.x + idx + a.ld, xn);
  if (D) ldg256(a.x + idx - a.unit, xd);
  if (U) ldg256(a.x + idx + a.unit, xu);
  const double xw = i0 > 0 ? __ldg(a.x + idx - 1) : 0.0, xe = i0 + 4 < a.n0 ? __ldg(a.x + idx + 4) : 0.0;
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const int  i = i0 + m;
    const bool Wn = i > 0, E = i < a.n0 - 1;
    const int  deg = (int)Wn + (int)E + (int)S + (int)N + (int)D + (int)U;
    double     ax  = 0.0;
    if (D) ax = fma(a.mh, xd[m], ax);
    if (S) ax = fma(a.mh, xs[m], ax);
    if (Wn) ax = fma(a.mh, m == 0 ? xw : xc[m == 0 ? 0 : m - 1], ax);
    ax = fma(a.diag[deg], xc[m], ax);
    if (E) ax = fma(a.mh, m == 3 ? xe : xc[m == 3 ? 3 : m + 1], ax);
    if (N) ax = fma(a.mh, xn[m], ax);
    if (U) ax = fma(a.mh, xu[m], ax);
    r[m] = i < a.n0 ? __dsub_rn(bb[m], ax) : 0.0;
  }
}

__global__ void __launch_bounds__(NT) resrest3d_kernel(const __grid_constant__ Args a)
{
  extern __shared__ __align__(32) double ring[]; // RING planes of R rows of W doubles
  const int tile = blockIdx.x, seg = blockIdx.y;
  const int I0 = (tile % a.tiles_i) * TI, J0 = (tile / a.tiles_i) * TJ;
  const int Ka = seg * a.kseg, Kb = min(a.c2, Ka + a.kseg);
  const int fi0 = 2 * I0 - 4, fj0 = 2 * J0 - 1; // first fine column / row of the footprint

  // residual of fine plane k of the footprint into its ring slot (planes outside the grid are skipped: never read)
  auto fill = [&](int k, int t0, int nthreads) {
    if (k < 0 || k >= a.n2) return;
    double *slot = ring + (size_t)((k + RING) % RING) * PLANE;
    for (int t = t0; t < (W / 4) * R; t += nthreads) {
      const int rj = t / (W / 4), q = t - rj * (W / 4);
      const int i0 = fi0 + 4 * q, j = fj0 + rj;
      if (i0 < 0 || i0 >= a.ld || j < 0 || j >= a.n1) continue;
      double r[4];
      residual_quad(a, i0, j, k, r);
      *reinterpret_cast<double4 *>(slot + rj * W + 4 * q) = make_double4(r[0], r[1], r[2], r[3]);
    }
  };
  fill(2 * Ka - 1, threadIdx.x, NT);
  for (int K = Ka; K < Kb; ++K) {
    // two planes per step: the tasks of both are dealt out as one list so that every thread has work
    for (int t = threadIdx.x; t < 2 * (W / 4) * R; t += NT) {
      const int p = t >= (W / 4) * R ? 1 : 0, tt = t - p * (W / 4) * R;
      const int k = 2 * K + p;
      if (k >= a.n2) continue;
      double   *slot = ring + (size_t)(k % RING) * PLANE;
      const int rj = tt / (W / 4), q = tt - rj * (W / 4);
      const int i0 = fi0 + 4 * q, j = fj0 + rj;
      if (i0 < 0 || i0 >= a.ld || j < 0 || j >= a.n1) continue;
      double r[4];
      residual_quad(a, i0, j, k, r);
      *reinterpret_cast<double4 *>(slot + rj * W + 4 * q) = make_double4(r[0], r[1], r[2], r[3]);
    }
    __syncthreads();
    // restriction: four coarse nodes per thread, restrict3_pitched_kernel's order (k, j, i ascending)
    for (int t = threadIdx.x; t < (TI / 4) * TJ; t += NT) {
      const int cj = t / (TI / 4), ct = t - cj * (TI / 4);
      const int J = J0 + cj, Ic = I0 + 4 * ct; // first of the thread's four coarse nodes
      if (J >= a.c1 || Ic >= a.c0) continue;
      double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int dk = -1; dk <= 1; ++dk) {
        const int k = 2 * K + dk;
        if (k < 0 || k >= a.n2) continue;
        const double *slot = ring + (size_t)((k + RING) % RING) * PLANE;
#pragma unroll
        for (int dj = -1; dj <= 1; ++dj) {
          const int j = 2 * J + dj;
          if (j < 0 || j >= a.n1) continue;
          const double *row = slot + (j - fj0) * W + (2 * Ic - fi0); // fine column 2 Ic of this row
          const int     i0 = 2 * Ic;
          double        v[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}; // fine columns i0-1 .. i0+7; columns >= ld were not computed
          if (i0 > 0) v[0] = row[-1];
          {
            const double4 q = *reinterpret_cast<const double4 *>(row);
            v[1] = q.x; v[2] = q.y; v[3] = q.z; v[4] = q.w;
          }
          if (i0 + 4 < a.ld) {
            const double4 q = *reinterpret_cast<const double4 *>(row + 4);
            v[5] = q.x; v[6] = q.y; v[7] = q.z; v[8] = q.w;
          }
          const double wjk = (dj ? 0.5 : 1.0) * (dk ? 0.5 : 1.0);
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            const int ic = i0 + 2 * m; // fine column of coarse node Ic + m
            if (ic - 1 >= 0 && ic - 1 < a.n0) acc[m] = fma(0.5 * wjk, v[2 * m], acc[m]);
            if (ic < a.n0) acc[m] = fma(wjk, v[2 * m + 1], acc[m]);
            if (ic + 1 < a.n0) acc[m] = fma(0.5 * wjk, v[2 * m + 2], acc[m]);
          }
        }
      }
#pragma unroll
      for (int m = 0; m < 4; ++m)
        if (Ic + m < a.c0) a.bc[Ic + m + (long long)a.cld * (J + (long long)a.c1 * K)] = acc[m];
    }
    // no second barrier: the next step writes the slots of planes 2K+2 and 2K+3, which this step does not read, and the
    // step after that comes after the next barrier
  }
}

} // namespace resrest3d
