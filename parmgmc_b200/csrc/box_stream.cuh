// box_stream.cuh -- fused four-colour Gibbs sweep for the 2D 9-point stencil-array levels of the Galerkin hierarchy.
//
// The launch-per-colour path (stencil_op.cu box_sweep_kernel) reads the iterate four times and the right-hand side
// sector-wise four times per sweep; this kernel does the whole sweep (colours (i mod 2) + 2 (j mod 2) in ascending order,
// descending for a backward sweep: src/mc_sor.c:257-285 on the 4-colouring of a 9-point stencil) in one pass:
//
//   rows of the first parity  r: colour a then colour b of row r need only OLD rows r-1, r+1        ("A rows")
//   rows of the second parity o: their two colours need NEW rows o-1, o+1                           ("B rows")
//
// so a warp walks down its band two rows at a time: A(e), then B(e-1), keeping four rows in registers (lane l owns columns
// c0+4l .. c0+4l+3 of a 128-column strip, lanes 0 / 31 recompute the neighbouring strips' edge columns, east / west and
// diagonal neighbours come from warp shuffles).  The result is written out of place; the band's last A row is recomputed by
// the next band.  The level vectors keep their natural layout (odd row lengths), so rows are fetched with plain loads
// (these levels are L2-resident: 2049^2 doubles = 34 MB).  Arithmetic per node is box_sweep_kernel's, fma for fma.
// Noise: one Philox call per lane and row on the padded index (philox.cuh).
#pragma once
#include "common.hpp"
#include "fastnormal.cuh"
#include "philox.cuh"

namespace boxstream {

constexpr int STRIP_OUT = 120;

struct Item {
  int strip, ja, jb; // output columns of strip `strip`, output rows [ja, jb)
};

struct Args {
  int           nx, ny;
  const Item   *items;
  int           nitems;
  int           flip; // 0: forward (colours 0,1,2,3), 1: backward (3,2,1,0)
  const double *xin, *b;
  double       *xout;
  const double *coef; // [9][nx ny] stencil arrays (nodes outside the constant interior)
  const double *idiag, *sqrtdiag;
  double        c[9], idiag_c, sd_c, omo; // the shared interior stencil and its coefficients
  int           ring, has_const;          // nodes at least `ring` away from the boundary carry the shared stencil
  int           mode;                     // PMG_NOISE_*
  const double *tape;
  PhiloxKeys    pk;
  uint32_t      call_lo, call_hi;
  int           pitch4; // nx rounded up to 4: row stride of the generator index
};

__device__ __forceinline__ double shfl_up1(double v) { return __shfl_up_sync(0xffffffffu, v, 1); }
__device__ __forceinline__ double shfl_dn1(double v) { return __shfl_down_sync(0xffffffffu, v, 1); }

template <bool INTERIOR, int TREP = 1> struct Warp {
  const Args              &a;
  const fastnormal::TablesT<TREP> ft;
  int                      lane, c;

  __device__ __forceinline__ Warp(const Args &a_, const fastnormal::TablesT<TREP> &ft_, int lane_, int c_) : a(a_), ft(ft_), lane(lane_), c(c_) {}

  __device__ __forceinline__ void load_row(const double *__restrict__ v, int j, double (&out)[4]) const
  {
    if (v == nullptr) {
      out[0] = out[1] = out[2] = out[3] = 0.0;
      return;
    }
    const double *p = v + (long long)j * a.nx + c;
    if (INTERIOR) {
#pragma unroll
      for (int m = 0; m < 4; ++m) out[m] = p[m];
    } else {
      const bool rowok = j >= 0 && j < a.ny;
#pragma unroll
      for (int m = 0; m < 4; ++m) out[m] = (rowok && c + m >= 0 && c + m < a.nx) ? p[m] : 0.0;
    }
  }

  // pull the lane's 32 bytes of row j (two lines when the row start is not 32-byte aligned) towards L1
  __device__ __forceinline__ void prefetch_row(const double *__restrict__ v, int j) const
  {
    const double *p = v + (long long)j * a.nx + c;
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p + 3));
  }

  // the four normals of row j (zero where the node does not exist)
  __device__ __forceinline__ void noise_row(int j, double (&z)[4]) const
  {
    if (a.mode == PMG_NOISE_NONE) {
      z[0] = z[1] = z[2] = z[3] = 0.0;
    } else if (a.mode == PMG_NOISE_INJECTED) {
      load_row(a.tape, j, z);
    } else {
      const long long quad = ((long long)j * a.pitch4 + c) >> 2;
      uint32_t        w0, w1, w2, w3;
      philox4x32_10_keys((uint32_t)quad, (uint32_t)((unsigned long long)quad >> 32), a.call_lo, a.call_hi, a.pk, w0, w1, w2, w3);
      fastnormal::box_muller(ft, w0, w1, z[0], z[1]);
      fastnormal::box_muller(ft, w2, w3, z[2], z[3]);
    }
  }

  // value of column M + d of a row whose columns c-1 and c+4 are xw and xe
  template <int Q> static __device__ __forceinline__ double col(const double (&X)[4], double xw, double xe) { return Q < 0 ? xw : (Q > 3 ? xe : X[Q < 0 ? 0 : (Q > 3 ? 3 : Q)]); }

  // one node (box_sweep_kernel): sum = w - sum_s c_s x_s in ascending stencil order, x = omo x + idiag sum
  template <int M> __device__ __forceinline__ void node(int j, double (&T)[4], const double (&S)[4], const double (&N)[4], double sw, double se, double tw, double te, double nw, double ne, double bval, double z) const
  {
    const double v0 = col<M - 1>(S, sw, se), v1 = col<M>(S, sw, se), v2 = col<M + 1>(S, sw, se);
    const double v3 = col<M - 1>(T, tw, te), v5 = col<M + 1>(T, tw, te);
    const double v6 = col<M - 1>(N, nw, ne), v7 = col<M>(N, nw, ne), v8 = col<M + 1>(N, nw, ne);
    if (INTERIOR) {
      double sum = a.mode == PMG_NOISE_NONE ? bval : __dadd_rn(__dmul_rn(z, a.sd_c), bval);
      sum = fma(-a.c[0], v0, sum);
      sum = fma(-a.c[1], v1, sum);
      sum = fma(-a.c[2], v2, sum);
      sum = fma(-a.c[3], v3, sum);
      sum = fma(-a.c[5], v5, sum);
      sum = fma(-a.c[6], v6, sum);
      sum = fma(-a.c[7], v7, sum);
      sum = fma(-a.c[8], v8, sum);
      const double t0 = __dmul_rn(a.omo, T[M]);
      T[M]            = fma(a.idiag_c, sum, t0);
      return;
    }
    const int i = c + M;
    if (i < 0 || i >= a.nx || j < 0 || j >= a.ny) return;
    const long long idx = (long long)j * a.nx + i, nl = (long long)a.nx * a.ny;
    const bool      shared = a.has_const && i >= a.ring && i < a.nx - a.ring && j >= a.ring && j < a.ny - a.ring;
    const double    sd = shared ? a.sd_c : a.sqrtdiag[idx], id = shared ? a.idiag_c : a.idiag[idx];
    double          sum = a.mode == PMG_NOISE_NONE ? bval : __dadd_rn(__dmul_rn(z, sd), bval);
    const double    v[9] = {v0, v1, v2, v3, 0.0, v5, v6, v7, v8};
#pragma unroll
    for (int s = 0; s < 9; ++s) {
      if (s == 4) continue;
      const int di = s % 3 - 1, dj = s / 3 - 1;
      if (i + di < 0 || i + di >= a.nx || j + dj < 0 || j + dj >= a.ny) continue; // structurally absent entry
      const double cs = shared ? a.c[s] : a.coef[(long long)s * nl + idx];
      sum             = fma(-cs, v[s], sum);
    }
    const double t0 = __dmul_rn(a.omo, T[M]);
    T[M]            = fma(id, sum, t0);
  }

  // both colours of row j: columns of parity PC first, then the others (which see the first ones updated)
  template <int PC> __device__ __forceinline__ void row_update(int j, double (&T)[4], const double (&S)[4], const double (&N)[4], const double (&bv)[4], const double (&z)[4]) const
  {
    const double sw = shfl_up1(S[3]), se = shfl_dn1(S[0]), nw = shfl_up1(N[3]), ne = shfl_dn1(N[0]);
    {
      const double tw = shfl_up1(T[3]), te = shfl_dn1(T[0]);
      node<PC>(j, T, S, N, sw, se, tw, te, nw, ne, bv[PC], z[PC]);
      node<PC + 2>(j, T, S, N, sw, se, tw, te, nw, ne, bv[PC + 2], z[PC + 2]);
    }
    {
      const double tw = shfl_up1(T[3]), te = shfl_dn1(T[0]);
      node<1 - PC>(j, T, S, N, sw, se, tw, te, nw, ne, bv[1 - PC], z[1 - PC]);
      node<3 - PC>(j, T, S, N, sw, se, tw, te, nw, ne, bv[3 - PC], z[3 - PC]);
    }
  }

  __device__ __forceinline__ void store_row(int j, const double (&T)[4], bool out_lane) const
  {
    if (!out_lane) return;
    double *p = a.xout + (long long)j * a.nx + c;
#pragma unroll
    for (int m = 0; m < 4; ++m)
      if (INTERIOR || c + m < a.nx) p[m] = T[m];
  }

  template <int PC> __device__ __forceinline__ void run(const Item it)
  {
    const int  pr = PC; // parity of the A rows == parity of the columns that go first == the sweep direction
    const bool out_lane = lane >= 1 && lane <= 30 && (INTERIOR || (c >= 0 && c < a.nx));
    int        e0 = it.ja - 1;
    if ((e0 & 1) != pr) ++e0;
    int elast = it.jb;
    if ((elast & 1) != pr) --elast;
    double Rm2[4] = {0, 0, 0, 0}, Rm1[4], R0[4];
    load_row(a.xin, e0 - 1, Rm1);
    load_row(a.xin, e0, R0);
    for (int e = e0; e <= elast; e += 2) {
      // everything this step reads is requested first, the two noise rows (independent dependency chains) are generated
      // while the loads are in flight, and the rows of the NEXT step are pulled towards L1 meanwhile
      double     Rp1[4], Rp2[4], bA[4], bB[4], zA[4], zB[4];
      const bool doB = e - 1 >= it.ja && e - 1 < it.jb; // warp-uniform
      load_row(a.xin, e + 1, Rp1);
      load_row(a.xin, e + 2, Rp2);
      load_row(a.b, e, bA);
      if (doB) load_row(a.b, e - 1, bB);
      if (INTERIOR && e + 2 <= elast) {
        prefetch_row(a.xin, e + 3);
        prefetch_row(a.xin, e + 4);
        if (a.b) {
          prefetch_row(a.b, e + 2);
          prefetch_row(a.b, e + 1);
        }
        if (a.mode == PMG_NOISE_INJECTED) {
          prefetch_row(a.tape, e + 2);
          prefetch_row(a.tape, e + 1);
        }
      }
      noise_row(e, zA);
      if (doB) noise_row(e - 1, zB);
      row_update<PC>(e, R0, Rm1, Rp1, bA, zA);
      if (e >= it.ja && e < it.jb) store_row(e, R0, out_lane);
      if (doB) {
        row_update<PC>(e - 1, Rm1, Rm2, R0, bB, zB);
        store_row(e - 1, Rm1, out_lane);
      }
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        Rm2[m] = R0[m];
        Rm1[m] = Rp1[m];
        R0[m]  = Rp2[m];
      }
    }
  }
};

// TREP: interleaved copies of the Box-Muller tables (fastnormal.cuh): fewer bank conflicts in the gathers, more to load per CTA
template <int WARPS, int MINB, int TREP = 1> __global__ void __launch_bounds__(WARPS * 32, MINB) box_stream_kernel(const Args a)
{
  __shared__ fastnormal::SharedTablesT<TREP> fts;
  const fastnormal::TablesT<TREP>            ft = fastnormal::load_tables(fts);
  __syncthreads();
  const int lane = threadIdx.x & 31, w = blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (w >= a.nitems) return;
  const Item it = a.items[w];
  const int  c0 = it.strip * STRIP_OUT - 4, c = c0 + 4 * lane;
  // rows touched: ja-2 .. jb+2; rows updated: ja-1 .. jb
  const int  r = a.ring;
  const bool interior = a.has_const && c0 >= r && c0 + 127 <= a.nx - 1 - r && it.ja - 1 >= r && it.jb <= a.ny - 1 - r && it.ja - 2 >= 0 && it.jb + 2 <= a.ny - 1;
  if (interior) {
    Warp<true, TREP> W(a, ft, lane, c);
    if (a.flip) W.template run<1>(it);
    else W.template run<0>(it);
  } else {
    Warp<false, TREP> W(a, ft, lane, c);
    if (a.flip) W.template run<1>(it);
    else W.template run<0>(it);
  }
}

} // namespace boxstream
