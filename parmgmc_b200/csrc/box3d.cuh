// box3d.cuh -- plane kernels for the 3D 27-point Galerkin levels of the V-cycle.
//
// The eight colours (i mod 2) + 2 (j mod 2) + 4 (k mod 2) of src/mc_sor.c:257-285 on a 27-point level split by the parity of
// k: the colours of one k-parity only read planes of the other parity from outside their own plane.  box3_sweep_kernel sweeps
// the FOUR colours of one k-parity in one launch (two launches per sweep instead of eight, or four with box_pair_sweep3_kernel):
// one CTA owns a whole grid plane, updates it IN PLACE and walks down it in blocks of grid rows, so that its working set
// (the rows of three planes it is reading) stays in the SM's L1 and every value comes from HBM / L2 once per launch:
//
//   forward   block m:  colours 0, 1 on the even rows of (2Rm, 2R(m+1)]  (+ row 0 in block 0)  -- they need OLD odd rows --
//                       then colours 2, 3 on the odd rows of (2Rm, 2R(m+1))                      -- they need NEW even rows
//   backward  block m:  colours 3, 2 on the odd rows of (2Rm, 2R(m+1)), then colours 1, 0 on the even rows of [2Rm, 2R(m+1))
//
// with a block barrier between colours (a node never reads a node of its own colour, so a colour phase is race free, and
// stores of a phase are visible to the CTA's later loads through the SM's own L1).  The block's normals are generated first,
// four per Philox call, into shared memory (a node-at-a-time sweep would use one of the four values each call returns).
//
// Boundary handling without coefficient loads: the Galerkin coarsening of the constant-coefficient fine operator yields the
// same stencil at every node of a boundary CLASS (x class, y class, z class, each first / interior / last; verified bitwise
// against every node at set-up, BoxOp::detect_classes3).  Interior nodes read the interior class from kernel parameters,
// the others from a 27-entry class table in shared memory; a structurally absent neighbour has coefficient 0 in its class
// and a clamped (valid) address, so it contributes fma(0, v, s) = s.
//
// Arithmetic per node is box_sweep_kernel<3>'s / box_apply_kernel<3>'s, fma for fma (bit-identical results; tested).
#pragma once
#include <type_traits>

#include "common.hpp"
#include "philox.cuh"

namespace box3d {

// one class: NEGATED coefficients in stencil order s = (di+1) + 3 (dj+1) + 9 (dk+1) (entry 13, the centre, holds -a_ii),
// omega / a_ii, sqrt((2-omega)/omega) sqrt(a_ii)
struct __align__(16) Cls {
  double nc[27];
  double idiag, sd;
  double pad;
};
struct Tab {
  Cls c[27]; // [cx + 3 cy + 9 cz]
};

struct Args {
  int     n0, n1, n2;   // global grid
  int     slo, shi;     // owned planes
  int     R;            // even rows per block
  int     pitch4;       // n0 rounded up to 4 (generator index space, philox.cuh)
  double  omo;          // 1 - omega
  Cls     in;           // the interior class
  const Tab    *tab;    // device copy of all classes
  const double *b;      // natural layout, owned planes; may be null (b = 0)
  double       *x;      // natural layout, owned planes; updated in place
  const double *glo, *ghi; // ghost planes slo-1 and shi (slabs)
  const double *out_b;  // apply: right-hand side of the residual (null: plain product)
  double       *out;    // apply: result
};

__device__ __forceinline__ int cls1(int i, int n) { return i == 0 ? 0 : (i == n - 1 ? 2 : 1); }

// sum += sum_{s != 13} nc[s] * x_s in ascending stencil order; P[dk] are the plane bases, ro[dj] row offsets, co[di] columns
template <bool CENTRE> __device__ __forceinline__ double chain27(const double *nc, const double *const (&P)[3], const int (&ro)[3], const int (&co)[3], double acc)
{
#pragma unroll
  for (int dk = 0; dk < 3; ++dk) { // the nine loads of a plane are issued before its fmas (one memory latency per plane, not per term)
    double v[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) v[q] = P[dk][ro[q / 3] + co[q % 3]];
#pragma unroll
    for (int q = 0; q < 9; ++q) {
      if (!CENTRE && dk == 1 && q == 4) continue;
      acc = fma(nc[9 * dk + q], v[q], acc);
    }
  }
  return acc;
}

// plane bases of planes k-1, k, k+1 (absent planes alias plane k: their class coefficients are zero)
__device__ __forceinline__ void plane_ptrs(const Args &a, const double *x, int k, const double *(&P)[3])
{
  const long long unit = (long long)a.n0 * a.n1;
  P[1] = x + (long long)(k - a.slo) * unit;
  P[0] = k - 1 < 0 ? P[1] : (k - 1 < a.slo ? a.glo : P[1] - unit);
  P[2] = k + 1 >= a.n2 ? P[1] : (k + 1 >= a.shi ? a.ghi : P[1] + unit);
}

template <int NT> __global__ void __launch_bounds__(NT, 1) box3_sweep_kernel(const __grid_constant__ Args a, int kpar, int backward, NoiseArgs na)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Cls    *cls = reinterpret_cast<Cls *>(smem_raw);
  double *zs  = reinterpret_cast<double *>(cls + 27);
  const int k = a.slo + ((kpar ^ a.slo) & 1) + 2 * (int)blockIdx.x;
  if (k >= a.shi) return;
  for (int q = threadIdx.x; q < (int)(sizeof(Tab) / sizeof(double)); q += NT) reinterpret_cast<double *>(cls)[q] = reinterpret_cast<const double *>(a.tab)[q];
  const double *P[3];
  plane_ptrs(a, a.x, k, P);
  double       *xk = a.x + (long long)(k - a.slo) * a.n0 * a.n1;
  const double *bk = a.b ? a.b + (long long)(k - a.slo) * a.n0 * a.n1 : nullptr;
  const double *tk = na.mode == PMG_NOISE_INJECTED ? na.tape + (long long)(k - a.slo) * a.n0 * a.n1 : nullptr;
  const int     cz = cls1(k, a.n2), R = a.R, n0 = a.n0, n1 = a.n1;
  const int     qrow  = a.pitch4 >> 2;
  __syncthreads();

  // one colour phase: column parity ci on rows ja, ja+2, ... < jb.  The first / last column of a row belong to boundary classes:
  // they are items of their own at the END of the work list, so that the warps of the interior columns never diverge into the
  // class-table path (a phase ends with a block barrier: one slow lane per grid row would hold up every warp).
  auto phase = [&](int ci, int ja, int jb, int jz0) {
    const int nr = jb > ja ? (jb - ja + 1) / 2 : 0;
    const int i_first = ci == 0 ? 2 : 1;                          // first interior column of this parity
    const int ni      = n0 - 2 >= i_first ? (n0 - 2 - i_first) / 2 + 1 : 0; // interior columns i_first, i_first + 2, ... <= n0 - 2
    const int nb      = (ci == 0 ? 1 : 0) + ((((n0 - 1) & 1) == ci && n0 > 1) ? 1 : 0); // boundary columns 0 and / or n0 - 1
    const int   tot_int = nr * ni, total = tot_int + nr * nb;
    const float rni = 1.0f / (float)max(ni, 1); // w / ni without an integer division ((w + 0.5) / ni is never within rounding of an integer: w < 2^20, ni < 2^12)
    for (int w = threadIdx.x; w < total; w += NT) {
      int r, i;
      if (w < tot_int) {
        r = (int)(((float)w + 0.5f) * rni);
        i = i_first + 2 * (w - r * ni);
      } else {
        const int q = w - tot_int;
        r           = nb == 2 ? q >> 1 : q;
        i           = (ci == 0 && q - r * nb == 0) ? 0 : n0 - 1;
      }
      const int j  = ja + 2 * r;
      const int cx = cls1(i, n0), cy = cls1(j, n1);
      const int row = j * n0;
      double    z   = 0.0;
      if (na.mode == PMG_NOISE_PHILOX) z = zs[(j - jz0) * a.pitch4 + i];
      else if (na.mode == PMG_NOISE_INJECTED) z = tk[row + i];
      const double bv = bk ? bk[row + i] : 0.0;
      const double xo = xk[row + i];
      double       xn;
      if (cx == 1 && cy == 1 && cz == 1) {
        const int ro[3] = {row - n0, row, row + n0}, co[3] = {i - 1, i, i + 1};
        double    sum   = na.mode == PMG_NOISE_NONE ? bv : __dadd_rn(__dmul_rn(z, a.in.sd), bv); // noisy_rhs_id
        sum             = chain27<false>(a.in.nc, P, ro, co, sum);
        xn              = fma(a.in.idiag, sum, __dmul_rn(a.omo, xo));
      } else {
        const Cls &c    = cls[cx + 3 * cy + 9 * cz];
        const int  ro[3] = {j > 0 ? row - n0 : row, row, j < n1 - 1 ? row + n0 : row}, co[3] = {i > 0 ? i - 1 : i, i, i < n0 - 1 ? i + 1 : i};
        double     sum   = na.mode == PMG_NOISE_NONE ? bv : __dadd_rn(__dmul_rn(z, c.sd), bv);
        sum              = chain27<false>(c.nc, P, ro, co, sum);
        xn               = fma(c.idiag, sum, __dmul_rn(a.omo, xo));
      }
      xk[row + i] = xn;
    }
    __syncthreads();
  };
  // the block's normals: rows [jz0, jz1), four per generator call
  auto noise = [&](int jz0, int jz1) {
    if (na.mode != PMG_NOISE_PHILOX) return;
    const int total = (jz1 - jz0) * qrow;
    const float rq = 1.0f / (float)qrow;
    for (int w = threadIdx.x; w < total; w += NT) {
      const int r = (int)(((float)w + 0.5f) * rq), qi = w - r * qrow;
      double    z[4];
      philox_normal_quad(na.seed, na.call, (uint64_t)((long long)k * n1 + jz0 + r) * (uint64_t)qrow + (uint64_t)qi, z);
      double *d = zs + r * a.pitch4 + 4 * qi;
      d[0] = z[0]; d[1] = z[1]; d[2] = z[2]; d[3] = z[3];
    }
    __syncthreads();
  };

  for (int m = 0;; ++m) {
    const int lo = 2 * R * m, hi = 2 * R * (m + 1);
    if (lo >= n1 + 1) break;
    if (!backward) {
      const int aa = m == 0 ? 0 : lo + 2, ab = min(hi + 1, n1); // even rows
      const int ba = lo + 1, bb = min(hi, n1);                  // odd rows
      const int jz0 = m == 0 ? 0 : lo + 1, jz1 = min(hi + 1, n1);
      if (jz1 <= jz0) break;
      noise(jz0, jz1);
      phase(0, aa, ab, jz0);
      phase(1, aa, ab, jz0);
      phase(0, ba, bb, jz0);
      phase(1, ba, bb, jz0);
    } else {
      const int ba = lo + 1, bb = min(hi, n1); // odd rows
      const int aa = lo, ab = min(hi, n1);     // even rows of [lo, hi)
      const int jz0 = lo, jz1 = min(hi, n1);
      if (jz1 <= jz0) break;
      noise(jz0, jz1);
      phase(1, ba, bb, jz0);
      phase(0, ba, bb, jz0);
      phase(1, aa, ab, jz0);
      phase(0, aa, ab, jz0);
    }
  }
}

// ---- the same sweep with the block's rows staged in shared memory, DE-INTERLEAVED by column parity ------------------------
// box3_sweep_kernel reads its 26 neighbours with stride-2 64-bit loads: a warp's load touches 512 bytes of L1 for 256 useful
// ones, and the kernel is bound by L1 wavefronts (profiles/r2_summary.md).  Here the rows of planes k-1, k, k+1 live in a RING
// of shared-memory rows with the even columns of a grid row in one array and the odd columns in another (zeros for rows /
// planes / columns outside the grid: their class coefficients are zero, so they add fma(0, 0, s) = s).  The rows block m+1
// needs beyond those of block m are fetched with cp.async (8 bytes per copy, which de-interleaves for free) while block m is
// swept, so no load latency is exposed; the rows of plane k are never re-read from global memory (the phases write their
// results to shared AND global memory).  The noisy right-hand side w = b + sqrtdiag z of the block's rows is formed once, four
// nodes per generator call, into shared memory before the block's phases: a phase touches global memory with stores only.
// A colour phase reads unit-stride: a thread owns TWO neighbouring nodes of its colour (array indices q0, q0 + 1, q0 even)
// and fetches, per neighbouring row, the pair of its own parity and the aligned pair of the other parity with one 128-bit load
// each, plus one 64-bit load for the third column of the other parity -- 27 load instructions and 360 bytes for two nodes
// instead of 52 and 832.  The first and last column of a row (boundary classes in x) are items of their own at the end of the
// work list; boundary rows / planes run the pair code with their class's coefficients from the table.
// Arithmetic per node: box_sweep_kernel<3>'s, fma for fma.
struct SmemGeom {
  int H;  // doubles per parity array of a staged row (2 leading zeros, the columns, trailing zeros), even
  int RR; // rows of the ring per plane = 4R + 3: rows lo-1 .. hi+1 of the block being swept, and the 2R further rows of the next
};
__device__ __forceinline__ int soff(int col, int H) { return ((col & 1) ? H + 2 : 2) + (col >> 1); } // col = -1 and col = n0 fall on zero pads
__device__ __forceinline__ void cp_async8(double *dst, const double *src)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

template <int NT> __global__ void __launch_bounds__(NT, 1) box3_sweep_smem_kernel(const __grid_constant__ Args a, int kpar, int backward, NoiseArgs na, SmemGeom sg)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Cls    *cls = reinterpret_cast<Cls *>(smem_raw);
  double *zs  = reinterpret_cast<double *>(cls + 27); // noisy right-hand side of the block's rows, [2R+1][pitch4]
  const int R = a.R, n0 = a.n0, n1 = a.n1, H = sg.H, RS = 2 * sg.H, RR = sg.RR;
  double *xs  = zs + (size_t)(2 * R + 1) * a.pitch4; // [3][RR][RS]
  const int k = a.slo + ((kpar ^ a.slo) & 1) + 2 * (int)blockIdx.x;
  if (k >= a.shi) return;
  for (int q = threadIdx.x; q < (int)(sizeof(Tab) / sizeof(double)); q += NT) reinterpret_cast<double *>(cls)[q] = reinterpret_cast<const double *>(a.tab)[q];
  for (int q = threadIdx.x; q < 3 * RR * RS; q += NT) xs[q] = 0.0; // pads and absent planes stay zero for the whole launch
  const double *P[3];
  plane_ptrs(a, a.x, k, P);
  const double *P0 = k - 1 >= 0 ? P[0] : nullptr, *P1 = P[1], *P2 = k + 1 < a.n2 ? P[2] : nullptr;
  double       *xk = a.x + (long long)(k - a.slo) * n0 * n1;
  const double *bk = a.b ? a.b + (long long)(k - a.slo) * n0 * n1 : nullptr;
  const double *tk = na.mode == PMG_NOISE_INJECTED ? na.tape + (long long)(k - a.slo) * n0 * n1 : nullptr;
  const int     cz = cls1(k, a.n2);
  const int     qrow = a.pitch4 >> 2, ne = (n0 + 1) >> 1, no = n0 >> 1;
  const int     warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long PS = (long long)RR * RS; // plane stride of the ring
  __syncthreads();

  auto slot = [&](int j) { return (j + 1) % RR; }; // ring row of grid row j >= -1
  // rows [ja, jb) of the three planes into the ring: one (plane, row) per warp at a time
  auto prefetch = [&](int ja, int jb) {
    const int nrow = jb - ja;
    for (int rw = warp; rw < 3 * nrow; rw += NT / 32) {
      const int     dk = rw / nrow, j = ja + (rw - dk * nrow);
      const double *pl = dk == 0 ? P0 : (dk == 1 ? P1 : P2);
      if (pl == nullptr) continue;
      double *dst = xs + dk * PS + (long long)slot(j) * RS;
      if (j >= 0 && j < n1) {
        const double *src = pl + (long long)j * n0;
        for (int i = lane; i < n0; i += 32) cp_async8(dst + soff(i, H), src + i);
      } else {
        for (int i = lane; i < n0; i += 32) dst[soff(i, H)] = 0.0;
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // w = b + sqrtdiag z (noisy_rhs_id: two roundings) of rows [jz0, jz1), four nodes per generator call
  auto rhs_rows = [&](int jz0, int jz1) {
    const int   total = (jz1 - jz0) * qrow;
    const float rq = 1.0f / (float)qrow;
    for (int w = threadIdx.x; w < total; w += NT) {
      const int r = (int)(((float)w + 0.5f) * rq), qi = w - r * qrow, j = jz0 + r, i0 = 4 * qi;
      double    bv[4] = {0, 0, 0, 0}, z[4] = {0, 0, 0, 0};
#pragma unroll
      for (int m = 0; m < 4; ++m)
        if (bk && i0 + m < n0) bv[m] = bk[j * n0 + i0 + m];
      if (na.mode == PMG_NOISE_PHILOX) philox_normal_quad(na.seed, na.call, (uint64_t)((long long)k * n1 + j) * (uint64_t)qrow + (uint64_t)qi, z);
      else if (na.mode == PMG_NOISE_INJECTED) {
#pragma unroll
        for (int m = 0; m < 4; ++m)
          if (i0 + m < n0) z[m] = tk[j * n0 + i0 + m];
      }
      const int cy = cls1(j, n1);
      double   *d  = zs + r * a.pitch4 + i0;
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int    cx = cls1(i0 + m, n0);
        const double sd = (cx == 1 && cy == 1 && cz == 1) ? a.in.sd : cls[(i0 + m < n0 ? cx : 1) + 3 * cy + 9 * cz].sd;
        d[m]            = na.mode == PMG_NOISE_NONE ? bv[m] : __dadd_rn(__dmul_rn(z[m], sd), bv[m]);
      }
    }
  };
  // one node through the class table (first / last column of a row)
  auto node_generic = [&](int i, int j, int jz0) {
    const Cls &c   = cls[cls1(i, n0) + 3 * cls1(j, n1) + 9 * cz];
    double     sum = zs[(j - jz0) * a.pitch4 + i];
    const int  s1 = slot(j), s0 = s1 == 0 ? RR - 1 : s1 - 1, s2 = s1 + 1 == RR ? 0 : s1 + 1;
    const int  ro[3] = {s0 * RS, s1 * RS, s2 * RS};
#pragma unroll
    for (int dk = 0; dk < 3; ++dk)
#pragma unroll
      for (int dj = 0; dj < 3; ++dj) {
        const double *row = xs + dk * PS + ro[dj];
#pragma unroll
        for (int di = 0; di < 3; ++di) {
          if (dk == 1 && dj == 1 && di == 1) continue;
          sum = fma(c.nc[9 * dk + 3 * dj + di], row[soff(i + di - 1, H)], sum);
        }
      }
    double      *ctr = xs + PS + ro[1];
    const int    o   = soff(i, H);
    const double xn  = fma(c.idiag, sum, __dmul_rn(a.omo, ctr[o]));
    ctr[o]         = xn;
    xk[j * n0 + i] = xn;
  };
  // two neighbouring nodes of one colour (array indices q0, q0 + 1 of row j, both interior in x or masked out); the class of
  // the row / plane supplies the coefficients: kernel parameters for the interior class, the table otherwise
  auto pair = [&](auto itag, const Cls *c, int ci, int j, int q0, int jz0) {
    constexpr bool IN = decltype(itag)::value;
    const int      so = ci ? H + 2 : 2, to = ci ? 2 : H + 2, te_off = ci ? 2 : -1; // own-parity array, other-parity array, its third column
    const int      iA = 2 * q0 + ci, iB = iA + 2;
    const bool     vA = iA >= 1 && iA <= n0 - 2, vB = iB <= n0 - 2;
    const double   idiag = IN ? a.in.idiag : c->idiag;
    const double  *wrow = zs + (j - jz0) * a.pitch4;
    double         sA = vA ? wrow[iA] : 0.0, sB = vB ? wrow[iB] : 0.0, xoA = 0.0, xoB = 0.0;
    const int      s1 = slot(j), s0 = s1 == 0 ? RR - 1 : s1 - 1, s2 = s1 + 1 == RR ? 0 : s1 + 1;
    const int      ro[3] = {s0 * RS, s1 * RS, s2 * RS};
#pragma unroll
    for (int dk = 0; dk < 3; ++dk) {
      double2 s2v[3], t2[3];
      double  te[3];
#pragma unroll
      for (int dj = 0; dj < 3; ++dj) { // the loads of a plane are issued before its fmas
        const double *row = xs + dk * PS + ro[dj];
        s2v[dj] = *reinterpret_cast<const double2 *>(row + so + q0);
        t2[dj]  = *reinterpret_cast<const double2 *>(row + to + q0);
        te[dj]  = row[to + q0 + te_off];
      }
#pragma unroll
      for (int dj = 0; dj < 3; ++dj) {
        const double wA = ci ? t2[dj].x : te[dj], eA = ci ? t2[dj].y : t2[dj].x;
        const double wB = ci ? t2[dj].y : t2[dj].x, eB = ci ? te[dj] : t2[dj].y;
        const double n0c = IN ? a.in.nc[9 * dk + 3 * dj] : c->nc[9 * dk + 3 * dj], n1c = IN ? a.in.nc[9 * dk + 3 * dj + 1] : c->nc[9 * dk + 3 * dj + 1],
                     n2c = IN ? a.in.nc[9 * dk + 3 * dj + 2] : c->nc[9 * dk + 3 * dj + 2];
        sA = fma(n0c, wA, sA);
        sB = fma(n0c, wB, sB);
        if (dk == 1 && dj == 1) {
          xoA = s2v[dj].x;
          xoB = s2v[dj].y;
        } else {
          sA = fma(n1c, s2v[dj].x, sA);
          sB = fma(n1c, s2v[dj].y, sB);
        }
        sA = fma(n2c, eA, sA);
        sB = fma(n2c, eB, sB);
      }
    }
    const double xnA = fma(idiag, sA, __dmul_rn(a.omo, xoA)), xnB = fma(idiag, sB, __dmul_rn(a.omo, xoB));
    double      *ctr = xs + PS + ro[1] + so + q0;
    if (vA) {
      ctr[0]          = xnA;
      xk[j * n0 + iA] = xnA;
    }
    if (vB) {
      ctr[1]          = xnB;
      xk[j * n0 + iB] = xnB;
    }
  };
  // one colour phase: column parity ci on rows ja, ja+2, ... < jb.  Work list: the pairs of every row, then the first / last
  // column of the rows as items of their own
  auto phase = [&](int ci, int ja, int jb, int jz0) {
    const int   nr = jb > ja ? (jb - ja + 1) / 2 : 0;
    const int   nq = ci == 0 ? ne : no, np2 = (nq + 1) >> 1;
    const int   nb = (ci == 0 ? 1 : 0) + ((((n0 - 1) & 1) == ci && n0 > 1) ? 1 : 0); // boundary columns 0 and / or n0 - 1
    const int   tot_pairs = nr * np2, first_b = (tot_pairs + 31) & ~31, total = first_b + nr * nb; // the boundary items start a warp of their own:
    const float rnp = 1.0f / (float)max(np2, 1);                                                   // a warp that ran both paths would hold up the phase
    for (int w = threadIdx.x; w < total; w += NT) {
      if (w >= tot_pairs && w < first_b) continue;
      if (w < tot_pairs) {
        const int r = (int)(((float)w + 0.5f) * rnp), p = w - r * np2;
        const int j = ja + 2 * r, cy = cls1(j, n1);
        if (cy == 1 && cz == 1) pair(std::true_type{}, nullptr, ci, j, 2 * p, jz0);
        else pair(std::false_type{}, cls + 1 + 3 * cy + 9 * cz, ci, j, 2 * p, jz0);
      } else {
        const int q = w - first_b, r = nb == 2 ? q >> 1 : q;
        const int i = (ci == 0 && q - r * nb == 0) ? 0 : n0 - 1, j = ja + 2 * r;
        node_generic(i, j, jz0);
      }
    }
    __syncthreads();
  };

  prefetch(-1, 2 * R + 2); // rows -1 .. 2R+1 of block 0
  for (int m = 0;; ++m) {
    const int lo = 2 * R * m, hi = 2 * R * (m + 1);
    if (lo >= n1 + 1) break;
    const int jz0 = backward ? lo : (m == 0 ? 0 : lo + 1), jz1 = backward ? min(hi, n1) : min(hi + 1, n1);
    if (jz1 <= jz0) break;
    rhs_rows(jz0, jz1);
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    prefetch(hi + 2, hi + 2 * R + 2); // what block m+1 reads beyond this block's rows; the ring rows it overwrites are those of rows < lo-1
    if (!backward) {
      const int aa = m == 0 ? 0 : lo + 2, ab = min(hi + 1, n1); // even rows
      const int ba = lo + 1, bb = min(hi, n1);                  // odd rows
      phase(0, aa, ab, jz0);
      phase(1, aa, ab, jz0);
      phase(0, ba, bb, jz0);
      phase(1, ba, bb, jz0);
    } else {
      const int ba = lo + 1, bb = min(hi, n1); // odd rows
      const int aa = lo, ab = min(hi, n1);     // even rows of [lo, hi)
      phase(1, ba, bb, jz0);
      phase(0, ba, bb, jz0);
      phase(1, aa, ab, jz0);
      phase(0, aa, ab, jz0);
    }
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
}

// out = b - A x (RES) or A x on one plane per CTA; rows of the plane in blocks, unit-stride accesses
template <int NT, bool RES> __global__ void __launch_bounds__(NT, 1) box3_apply_kernel(const __grid_constant__ Args a)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Cls      *cls = reinterpret_cast<Cls *>(smem_raw);
  const int k   = a.slo + (int)blockIdx.x;
  if (k >= a.shi) return;
  for (int q = threadIdx.x; q < (int)(sizeof(Tab) / sizeof(double)); q += NT) reinterpret_cast<double *>(cls)[q] = reinterpret_cast<const double *>(a.tab)[q];
  const double *P[3];
  plane_ptrs(a, a.x, k, P);
  const long long poff = (long long)(k - a.slo) * a.n0 * a.n1;
  const int       cz = cls1(k, a.n2), n0 = a.n0, n1 = a.n1, total = n0 * n1;
  __syncthreads();
  const float rnin = 1.0f / (float)(n0 - 2);
  const int nin = n0 - 2, tot_int = nin * n1; // interior columns first, the two boundary columns of every row as items of their own
  for (int w0 = threadIdx.x; w0 < total; w0 += NT) {
    int j, i;
    if (w0 < tot_int) {
      j = (int)(((float)w0 + 0.5f) * rnin);
      i = 1 + w0 - j * nin;
    } else {
      const int q = w0 - tot_int;
      j           = q >> 1;
      i           = (q & 1) ? n0 - 1 : 0;
    }
    const int w  = j * n0 + i;
    const int cx = cls1(i, n0), cy = cls1(j, n1);
    double    acc;
    if (cx == 1 && cy == 1 && cz == 1) {
      const int ro[3] = {w - i - n0, w - i, w - i + n0}, co[3] = {i - 1, i, i + 1};
      acc             = chain27<true>(a.in.nc, P, ro, co, 0.0);
    } else {
      const Cls &c     = cls[cx + 3 * cy + 9 * cz];
      const int  row   = w - i;
      const int  ro[3] = {j > 0 ? row - n0 : row, row, j < n1 - 1 ? row + n0 : row}, co[3] = {i > 0 ? i - 1 : i, i, i < n0 - 1 ? i + 1 : i};
      acc              = chain27<true>(c.nc, P, ro, co, 0.0);
    }
    // acc = -(A x) with box_apply_kernel's roundings (negation is exact)
    a.out[poff + w] = RES ? __dadd_rn(a.out_b[poff + w], acc) : -acc;
  }
}

} // namespace box3d
