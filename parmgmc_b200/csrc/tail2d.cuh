// tail2d.cuh -- the coarse tail of the 2D V-cycle in ONE CTA, with every level vector in shared memory.
//
// Levels 0 .. nlev-1 of a geometric hierarchy (level 0: dense Cholesky sampler, levels >= 1: 9-point stencil-array levels,
// src/pc_gamgmc.c:242-259 / SURVEY Appendix A.3) are tiny -- 65 x 65 and below -- and were pure latency: one launch per
// pass (about 14 us each on B200, profiles/r2_summary.md) or, in grid_tail_kernel, a cluster barrier plus an L2 round trip
// per colour.  Here one CTA of 1024 threads keeps x and b of every level in its shared memory; a colour phase is a
// __syncthreads() and a few shared-memory loads.  Only the top level's right-hand side is read from global memory (it was
// written by the level above) and only the top level's iterate is written back.
//
// Arithmetic per node is box_sweep_kernel<2>'s / box_apply_kernel<2>'s / restrict_kernel<2>'s / prolong_kernel<2>'s /
// tri_gemv_kernel's, fma for fma, with box2d.cuh's boundary classes (a structurally absent neighbour has coefficient 0 and a
// clamped address): the result is bit-identical to the launch-per-colour path (tested).
#pragma once
#include "box2d.cuh"
#include "common.hpp"
#include "philox.cuh"

namespace tail2d {

constexpr int NT = 1024, MAX_LEVELS = 8, MAX_NOISE = 2 * MAX_LEVELS * 8 + 1;

struct Level {
  int    n0, n1, n;
  int    xoff, boff; // offsets (doubles) into the shared-memory arena; boff < 0: the right-hand side is Args::btop (global)
  int    ndirs, dirs[8];
  double omo;
};
struct Args {
  int               nlev;
  Level             lv[MAX_LEVELS];
  const box2d::Cls *cls; // [nlev][9] (3 row class + column class), device
  const double     *btop;
  double           *xtop;
  int               zoff, tmpoff; // scratch (noise / residual) and the coarsest sampler's intermediate vector
  int               nc;
  const double     *W, *WT;
  int               woff, wtoff; // >= 0: W / W^T are staged in the arena too (they fit), at these offsets
  int               mode;
  uint64_t          seed;
  TailNoise         ns[MAX_NOISE];
};

__device__ __forceinline__ int cls1(int i, int n) { return i == 0 ? 0 : (i == n - 1 ? 2 : 1); }

__device__ __forceinline__ void sweep(const Level &L, const box2d::Cls *cls, const fastnormal::Tables &ft, double *x, const double *b, double *zs, int dir, int mode, uint64_t seed, const TailNoise &tn)
{
  const int n0 = L.n0, n1 = L.n1, tid = threadIdx.x;
  if (mode == PMG_NOISE_PHILOX) { // the level's normals, four per generator call (padded index space, philox.cuh)
    const int qrow = (n0 + 3) >> 2, nq = qrow * n1;
    const float rq = 1.0f / (float)qrow;
    for (int q = tid; q < nq; q += NT) {
      const int row = (int)(((float)q + 0.5f) * rq), qi = q - row * qrow;
      double    z[4];
      uint32_t  w0, w1, w2, w3;
      philox4x32_10((uint32_t)q, 0u, (uint32_t)tn.call, (uint32_t)(tn.call >> 32), (uint32_t)seed, (uint32_t)(seed >> 32), w0, w1, w2, w3); // philox_normal_quad, tables in shared memory
      fastnormal::box_muller(ft, w0, w1, z[0], z[1]);
      fastnormal::box_muller(ft, w2, w3, z[2], z[3]);
#pragma unroll
      for (int m = 0; m < 4; ++m)
        if (4 * qi + m < n0) zs[row * n0 + 4 * qi + m] = z[m];
    }
    __syncthreads();
  }
  const box2d::Cls kin = cls[4]; // the interior class, in registers: interior nodes load no coefficients
  for (int s = 0; s < 4; ++s) {
    const int   c = dir == PMG_SOR_FORWARD_SWEEP ? s : 3 - s, ci = c & 1, cj = c >> 1;
    const int   ni = (n0 - ci + 1) / 2, nj = (n1 - cj + 1) / 2, total = ni * nj;
    const float rni = 1.0f / (float)ni; // t / ni without an integer division: (t + 0.5) / ni is never within rounding of an integer here
    for (int t = tid; t < total; t += NT) {
      const int    ty = (int)(((float)t + 0.5f) * rni), tx = t - ty * ni, i = 2 * tx + ci, j = 2 * ty + cj, idx = i + n0 * j;
      const double bv = b[idx];
      double       z  = 0.0;
      if (mode == PMG_NOISE_PHILOX) z = zs[idx];
      else if (mode == PMG_NOISE_INJECTED) z = tn.tape[idx];
      if (i > 0 && i < n0 - 1 && j > 0 && j < n1 - 1) {
        double       sum = mode == PMG_NOISE_NONE ? bv : __dadd_rn(__dmul_rn(z, kin.sd), bv); // noisy_rhs_id
        const double v0 = x[idx - n0 - 1], v1 = x[idx - n0], v2 = x[idx - n0 + 1], v3 = x[idx - 1], v5 = x[idx + 1], v6 = x[idx + n0 - 1], v7 = x[idx + n0], v8 = x[idx + n0 + 1];
        sum = fma(kin.nc[0], v0, sum);
        sum = fma(kin.nc[1], v1, sum);
        sum = fma(kin.nc[2], v2, sum);
        sum = fma(kin.nc[3], v3, sum);
        sum = fma(kin.nc[4], v5, sum);
        sum = fma(kin.nc[5], v6, sum);
        sum = fma(kin.nc[6], v7, sum);
        sum = fma(kin.nc[7], v8, sum);
        x[idx] = fma(kin.idiag, sum, __dmul_rn(kin.omo, x[idx]));
      } else {
        const box2d::Cls &k  = cls[3 * cls1(j, n1) + cls1(i, n0)];
        const int         iw = i > 0 ? -1 : 0, ie = i < n0 - 1 ? 1 : 0, rs = j > 0 ? -n0 : 0, rn = j < n1 - 1 ? n0 : 0;
        double            sum = mode == PMG_NOISE_NONE ? bv : __dadd_rn(__dmul_rn(z, k.sd), bv);
        const double v0 = x[idx + rs + iw], v1 = x[idx + rs], v2 = x[idx + rs + ie], v3 = x[idx + iw], v5 = x[idx + ie], v6 = x[idx + rn + iw], v7 = x[idx + rn], v8 = x[idx + rn + ie];
        sum = fma(k.nc[0], v0, sum);
        sum = fma(k.nc[1], v1, sum);
        sum = fma(k.nc[2], v2, sum);
        sum = fma(k.nc[3], v3, sum);
        sum = fma(k.nc[4], v5, sum);
        sum = fma(k.nc[5], v6, sum);
        sum = fma(k.nc[6], v7, sum);
        sum = fma(k.nc[7], v8, sum);
        x[idx] = fma(k.idiag, sum, __dmul_rn(k.omo, x[idx]));
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(NT, 1) tail2d_kernel(const __grid_constant__ Args a)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  fastnormal::SharedTables *fts  = reinterpret_cast<fastnormal::SharedTables *>(smem_raw);
  box2d::Cls               *clss = reinterpret_cast<box2d::Cls *>(fts + 1);
  double                   *arena = reinterpret_cast<double *>(clss + 9 * MAX_LEVELS);
  const int tid = threadIdx.x;
  pdl_launch_dependents();
  // everything constant comes into shared memory first, while the kernel before this one is still running: a kernel starts
  // with a cold L1, and every table line fetched on demand would cost an L2 round trip on the critical path of a phase
  const fastnormal::Tables ft = fastnormal::load_tables(*fts);
  for (int q = tid; q < 9 * a.nlev * (int)(sizeof(box2d::Cls) / sizeof(double)); q += NT) reinterpret_cast<double *>(clss)[q] = reinterpret_cast<const double *>(a.cls)[q];
  if (a.woff >= 0)
    for (int q = tid; q < a.nc * a.nc; q += NT) {
      arena[a.woff + q]  = a.W[q];
      arena[a.wtoff + q] = a.WT[q];
    }
  const int top = a.nlev - 1;
  int       kn  = 0; // cursor into the noise blocks, in the reference's consumption order (SURVEY 8(c) tape contract)
  double   *zs  = arena + a.zoff;
  for (int t = tid; t < a.lv[top].n; t += NT) arena[a.lv[top].xoff + t] = 0.0; // PCMG zeroes the iterate of the level it enters
  pdl_wait(); // btop was written by the kernel before this one
  __syncthreads();
  for (int l = top; l >= 1; --l) {
    const Level      &F = a.lv[l], &C = a.lv[l - 1];
    const box2d::Cls *cls = clss + 9 * l;
    double           *x = arena + F.xoff;
    const double     *b = F.boff >= 0 ? arena + F.boff : a.btop;
    for (int d = 0; d < F.ndirs; ++d, ++kn) sweep(F, cls, ft, x, b, zs, F.dirs[d], a.mode, a.seed, a.ns[kn]);
    const int n0 = F.n0, n1 = F.n1;
    {
      const box2d::Cls kin = cls[4];
      const float      rn0 = 1.0f / (float)n0;
      for (int t = tid; t < F.n; t += NT) { // r = b - A x (box2d::resid: negated coefficients, b + acc)
        const int j = (int)(((float)t + 0.5f) * rn0), i = t - j * n0;
        double    acc = 0.0;
        if (i > 0 && i < n0 - 1 && j > 0 && j < n1 - 1) {
          acc = fma(kin.nc[0], x[t - n0 - 1], acc);
          acc = fma(kin.nc[1], x[t - n0], acc);
          acc = fma(kin.nc[2], x[t - n0 + 1], acc);
          acc = fma(kin.nc[3], x[t - 1], acc);
          acc = fma(kin.ndiag, x[t], acc);
          acc = fma(kin.nc[4], x[t + 1], acc);
          acc = fma(kin.nc[5], x[t + n0 - 1], acc);
          acc = fma(kin.nc[6], x[t + n0], acc);
          acc = fma(kin.nc[7], x[t + n0 + 1], acc);
        } else {
          const box2d::Cls &k = cls[3 * cls1(j, n1) + cls1(i, n0)];
          const int         iw = i > 0 ? -1 : 0, ie = i < n0 - 1 ? 1 : 0, rs = j > 0 ? -n0 : 0, rn = j < n1 - 1 ? n0 : 0;
          acc = fma(k.nc[0], x[t + rs + iw], acc);
          acc = fma(k.nc[1], x[t + rs], acc);
          acc = fma(k.nc[2], x[t + rs + ie], acc);
          acc = fma(k.nc[3], x[t + iw], acc);
          acc = fma(k.ndiag, x[t], acc);
          acc = fma(k.nc[4], x[t + ie], acc);
          acc = fma(k.nc[5], x[t + rn + iw], acc);
          acc = fma(k.nc[6], x[t + rn], acc);
          acc = fma(k.nc[7], x[t + rn + ie], acc);
        }
        zs[t] = __dadd_rn(b[t], acc);
      }
    }
    __syncthreads();
    double *bc = arena + C.boff, *xc = arena + C.xoff;
    const float rcn0 = 1.0f / (float)C.n0;
    for (int t = tid; t < C.n; t += NT) { // b_c = P^T r (ascending fine index), x_c = 0
      const int J = (int)(((float)t + 0.5f) * rcn0), I = t - J * C.n0;
      double    acc = 0.0;
#pragma unroll
      for (int dj = -1; dj <= 1; ++dj)
#pragma unroll
        for (int di = -1; di <= 1; ++di) {
          const int i = 2 * I + di, j = 2 * J + dj;
          if (i < 0 || i >= n0 || j < 0 || j >= n1) continue;
          acc = fma((di ? 0.5 : 1.0) * (dj ? 0.5 : 1.0), zs[i + n0 * j], acc);
        }
      bc[t] = acc;
      xc[t] = 0.0;
    }
    __syncthreads();
  }
  { // coarsest level: y = W^T (W b + z), one warp per entry (tri_gemv_kernel's order)
    const Level     &C = a.lv[0];
    const TailNoise &tn = a.ns[kn];
    ++kn;
    const double   *b0 = arena + C.boff;
    double         *x0 = arena + C.xoff, *tmp = arena + a.tmpoff;
    const int       lane = tid & 31, gw = tid >> 5, nw = NT >> 5, n = a.nc;
    for (int i = gw; i < n; i += nw) {
      const double *row = (a.woff >= 0 ? arena + a.woff : a.W) + (size_t)i * n;
      double        acc = 0.0;
      for (int k = lane; k < i + 1; k += 32) acc = fma(row[k], b0[k], acc);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) {
        double zn = 0.0; // noise_value(na, i), with the shared-memory tables
        if (a.mode == PMG_NOISE_INJECTED) zn = tn.tape[i];
        else if (a.mode == PMG_NOISE_PHILOX) {
          uint32_t w0, w1, w2, w3;
          philox4x32_10((uint32_t)(i >> 2), 0u, (uint32_t)tn.call, (uint32_t)(tn.call >> 32), (uint32_t)a.seed, (uint32_t)(a.seed >> 32), w0, w1, w2, w3);
          double zc, zsn;
          if (i & 2) fastnormal::box_muller(ft, w2, w3, zc, zsn);
          else fastnormal::box_muller(ft, w0, w1, zc, zsn);
          zn = (i & 1) ? zsn : zc;
        }
        tmp[i] = a.mode != PMG_NOISE_NONE ? __dadd_rn(acc, zn) : acc;
      }
    }
    __syncthreads();
    for (int i = gw; i < n; i += nw) {
      const double *row = (a.wtoff >= 0 ? arena + a.wtoff : a.WT) + (size_t)i * n;
      double        acc = 0.0;
      for (int k = i + lane; k < n; k += 32) acc = fma(row[k], tmp[k], acc);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) x0[i] = acc;
    }
    __syncthreads();
  }
  for (int l = 1; l <= top; ++l) {
    const Level      &F = a.lv[l], &C = a.lv[l - 1];
    const box2d::Cls *cls = clss + 9 * l;
    double           *x = arena + F.xoff;
    const double     *b = F.boff >= 0 ? arena + F.boff : a.btop, *xc = arena + C.xoff;
    const float rfn0 = 1.0f / (float)F.n0;
    for (int t = tid; t < F.n; t += NT) { // x_f += P x_c (prolong_kernel: ascending coarse index)
      const int j = (int)(((float)t + 0.5f) * rfn0), i = t - j * F.n0;
      const int ci = (i & 1) ? 2 : 1, cj = (j & 1) ? 2 : 1, I0 = i >> 1, J0 = j >> 1;
      double    s = x[t];
      for (int bq = 0; bq < cj; ++bq) {
        const int J = J0 + bq;
        if (J >= C.n1) continue;
        for (int q = 0; q < ci; ++q) {
          const int I = I0 + q;
          if (I >= C.n0) continue;
          s = fma((ci == 2 ? 0.5 : 1.0) * (cj == 2 ? 0.5 : 1.0), xc[I + C.n0 * J], s);
        }
      }
      x[t] = s;
    }
    __syncthreads();
    for (int d = 0; d < F.ndirs; ++d, ++kn) sweep(F, cls, ft, x, b, zs, F.dirs[d], a.mode, a.seed, a.ns[kn]);
  }
  for (int t = tid; t < a.lv[top].n; t += NT) a.xtop[t] = arena[a.lv[top].xoff + t];
}

} // namespace tail2d
