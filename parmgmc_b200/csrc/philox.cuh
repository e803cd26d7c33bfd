// philox.cuh -- on-the-fly standard normals for the fused sweeps.
//
// Replaces VecSetRandomStandardNormal (reference src/parmgmc.c:70-116): instead of filling a whole
// vector from one sequential stream before each sweep, every row draws its own z from a
// counter-based generator keyed by (seed, draw#, global row), so the value does not depend on how
// rows are partitioned over threads, colours or GPUs.
//
// Definition (must stay identical to oracle/noise.c, which is only the checker):
//   id = global row of the DOF; for the matrix-free grid operators (LapOp) the PADDED natural index
//        id = (k ny + j) pitch + i with pitch = nx rounded up to a multiple of 4, so that the four columns a thread of the
//        streaming kernels owns always share one generator call
//   quad q = id >> 2;  ctr = (lo32 q, hi32 q, lo32 call, hi32 call);  key = (lo32 seed, hi32 seed)
//   (w0..w3) = philox4x32-10(ctr, key)
//   rows 4q, 4q+1 use (u1,u2) = ((w0+0.5) 2^-32, (w1+0.5) 2^-32); rows 4q+2, 4q+3 use (w2, w3) likewise
//   r = sqrt(-2 ln u1);  even row: z = r cospi(2 u2);  odd row: z = r sinpi(2 u2)
// i.e. the reference's Box-Muller pairing (i, i+1) -> (r cos, r sin) of src/parmgmc.c:100-110; one generator
// call serves four consecutive rows, which is what a thread of the fused sweeps owns.
#pragma once
#include <cstdint>

#include "common.hpp"
#include "fastnormal.cuh"

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

// the same rounds with the ten round keys precomputed on the host (streaming kernels: the keys sit in the constant bank
// and fold into the LOP3s)
struct PhiloxKeys {
  uint32_t k0[10], k1[10];
};
static inline void philox_expand_keys(uint64_t seed, PhiloxKeys &pk)
{
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; ++r) {
    pk.k0[r] = k0;
    pk.k1[r] = k1;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}
__device__ __forceinline__ void philox4x32_10_keys(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKeys &pk, uint32_t &o0, uint32_t &o1, uint32_t &o2, uint32_t &o3)
{
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ pk.k0[r], n2 = hi0 ^ c3 ^ pk.k1[r];
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
  }
  o0 = c0; o1 = c1; o2 = c2; o3 = c3;
}

// Box-Muller on two 32-bit words: (r cos, r sin); the arithmetic is fastnormal.cuh's in every kernel, so all device
// paths produce bit-identical normals
__device__ __forceinline__ void box_muller_32(uint32_t wa, uint32_t wb, double &zc, double &zs)
{
  fastnormal::box_muller(fastnormal::global_tables(), wa, wb, zc, zs);
}

// the four normals of rows 4q .. 4q+3
__device__ __forceinline__ void philox_normal_quad(uint64_t seed, uint64_t call, uint64_t quad, double z[4])
{
  uint32_t w0, w1, w2, w3;
  philox4x32_10((uint32_t)quad, (uint32_t)(quad >> 32), (uint32_t)call, (uint32_t)(call >> 32), (uint32_t)seed, (uint32_t)(seed >> 32), w0, w1, w2, w3);
  box_muller_32(w0, w1, z[0], z[1]);
  box_muller_32(w2, w3, z[2], z[3]);
}

// z of one DOF (per-row kernels: only one of the four values is used); `id` is the generator index of the DOF
__device__ __forceinline__ double noise_value_id(const NoiseArgs &na, int64_t local_row, uint64_t g)
{
  if (na.mode == PMG_NOISE_INJECTED) return na.tape[local_row];
  if (na.mode == PMG_NOISE_NONE) return 0.0;
  uint32_t       w0, w1, w2, w3;
  philox4x32_10((uint32_t)(g >> 2), (uint32_t)(g >> 34), (uint32_t)na.call, (uint32_t)(na.call >> 32), (uint32_t)na.seed, (uint32_t)(na.seed >> 32), w0, w1, w2, w3);
  double zc, zs;
  if (g & 2) box_muller_32(w2, w3, zc, zs);
  else box_muller_32(w0, w1, zc, zs);
  return (g & 1) ? zs : zc;
}
__device__ __forceinline__ double noise_value(const NoiseArgs &na, int64_t local_row) { return noise_value_id(na, local_row, (uint64_t)(na.row0 + local_row)); }

// w = (z * sqrtdiag) + b : the two roundings of VecPointwiseMult + VecAXPY (src/pc_mcgibbs.c:124-126)
__device__ __forceinline__ double noisy_rhs(const NoiseArgs &na, int64_t local_row, double sqrtdiag, double b)
{
  if (na.mode == PMG_NOISE_NONE) return b;
  return __dadd_rn(__dmul_rn(noise_value(na, local_row), sqrtdiag), b);
}
__device__ __forceinline__ double noisy_rhs_id(const NoiseArgs &na, int64_t local_row, uint64_t id, double sqrtdiag, double b)
{
  if (na.mode == PMG_NOISE_NONE) return b;
  return __dadd_rn(__dmul_rn(noise_value_id(na, local_row, id), sqrtdiag), b);
}
