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
