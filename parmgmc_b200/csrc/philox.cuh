// philox.cuh -- on-the-fly standard normals for the fused sweeps.
//
// Replaces VecSetRandomStandardNormal (reference src/parmgmc.c:70-116): instead of filling a whole
// vector from one sequential stream before each sweep, every row draws its own z from a
// counter-based generator keyed by (seed, draw#, global row), so the value does not depend on how
// rows are partitioned over threads, colours or GPUs.
//
// Definition (must stay identical to oracle/noise.c, which is only the checker):
//   pair p = global_row >> 1;  ctr = (lo32 p, hi32 p, lo32 call, hi32 call);  key = (lo32 seed, hi32 seed)
//   (w0..w3) = philox4x32-10(ctr, key)
//   u1 = (((w1:w0) >> 11) + 0.5) 2^-53,   u2 = (((w3:w2) >> 11) + 0.5) 2^-53
//   r = sqrt(-2 ln u1);   z[2p] = r cospi(2 u2);   z[2p+1] = r sinpi(2 u2)
// i.e. the reference's Box-Muller pairing (i, i+1) -> (r cos, r sin) of src/parmgmc.c:100-110.
#pragma once
#include <cstdint>

#include "common.hpp"

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t &o0, uint32_t &o1, uint32_t &o2, uint32_t &o3)
{
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    if (r) {
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
  }
  o0 = c0; o1 = c1; o2 = c2; o3 = c3;
}

// both normals of pair p
__device__ __forceinline__ void philox_normal_pair(uint64_t seed, uint64_t call, uint64_t pair, double &zc, double &zs)
{
  uint32_t w0, w1, w2, w3;
  philox4x32_10((uint32_t)pair, (uint32_t)(pair >> 32), (uint32_t)call, (uint32_t)(call >> 32), (uint32_t)seed, (uint32_t)(seed >> 32), w0, w1, w2, w3);
  const uint64_t a  = (((uint64_t)w1 << 32) | w0) >> 11;
  const uint64_t b  = (((uint64_t)w3 << 32) | w2) >> 11;
  const double   u1 = __dmul_rn(__dadd_rn((double)a, 0.5), 0x1p-53);
  const double   u2 = __dmul_rn(__dadd_rn((double)b, 0.5), 0x1p-53);
  const double   r  = sqrt(__dmul_rn(-2.0, log(u1)));
  double         s, c;
  sincospi(__dmul_rn(2.0, u2), &s, &c);
  zc = __dmul_rn(r, c);
  zs = __dmul_rn(r, s);
}

// z of one local row
__device__ __forceinline__ double noise_value(const NoiseArgs &na, int64_t local_row)
{
  if (na.mode == PMG_NOISE_INJECTED) return na.tape[local_row];
  if (na.mode == PMG_NOISE_NONE) return 0.0;
  const uint64_t g = (uint64_t)(na.row0 + local_row);
  double         zc, zs;
  philox_normal_pair(na.seed, na.call, g >> 1, zc, zs);
  return (g & 1) ? zs : zc;
}

// w = (z * sqrtdiag) + b : the two roundings of VecPointwiseMult + VecAXPY (src/pc_mcgibbs.c:124-126)
__device__ __forceinline__ double noisy_rhs(const NoiseArgs &na, int64_t local_row, double sqrtdiag, double b)
{
  if (na.mode == PMG_NOISE_NONE) return b;
  return __dadd_rn(__dmul_rn(noise_value(na, local_row), sqrtdiag), b);
}
