// chol.cu -- dense Cholesky sampler on the device: the PCCHOLSAMPLER analogue for the coarsest level.
//
// Reference: src/pc_chols.c:173-195 (dense potrf "L" once), :220-260 (trsv "L","N" / "L","T"),
// :284-287 (v = L^-1 b; v += z; y = L^-T v  =>  y ~ N(A^-1 b, A^-1)), :306-336 (cached forward solve).
// The factor is computed once on the host at set-up (not on the per-sample path) and kept on the
// device as L and L^T (both column-major, so every per-step access is a coalesced column read).
// Two per-sample paths:
//   gemv (default)  W = L^-1 is formed once at set-up; a sample is two triangular matrix-vector products
//                   v = W b + z, y = W^T v, one warp per output entry over coalesced rows (W is kept row-major and
//                   column-major).  Fully parallel: a 289-unknown coarsest level costs a few microseconds instead
//                   of the ~0.26 ms of two sequential substitutions (profiles/r1_summary.md).  Agrees with the
//                   substitution to rounding (cond(L) eps).
//   trsv            (-pc_cholsampler_b200_solve trsv) one CTA does both triangular solves from shared memory; per
//                   unknown the accumulation order is reference-BLAS dtrsv's (forward k ascending, transposed k
//                   descending), so the result is bit-identical to the sequential substitution (src/pc_chols.c:231,253).
#include "common.hpp"
#include "philox.cuh"

namespace {
constexpr int CHOL_MAX_N   = 4096;
constexpr int CHOL_THREADS = 1024;

// s[0..n) holds the rhs on entry and the solution on exit.  blk is a 32x33 staging area for the diagonal
// block, so that the 32 dependent steps of a block read shared memory instead of waiting on L2 each time.
__device__ void trsv_forward(int n, const double *__restrict__ L, double *s, double (*blk)[33])
{
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int kb = 0; kb < n; kb += 32) {
    const int kmax = min(32, n - kb);
    {
      const int r = tid & 31, c = tid >> 5; // 1024 threads: one block entry each, column-major read is coalesced
      if (r < kmax && c < kmax) blk[r][c] = L[(kb + r) + (size_t)(kb + c) * n];
    }
    __syncthreads();
    if (warp == 0) {
      const int i  = kb + lane;
      double    si = i < n ? s[i] : 0.0;
      for (int k = 0; k < kmax; ++k) {
        double xk = 0.0;
        if (lane == k) xk = si / blk[k][k];
        xk = __shfl_sync(0xffffffffu, xk, k);
        if (lane == k) si = xk;
        else if (lane > k && i < n) si = fma(-blk[lane][k], xk, si);
      }
      if (i < n) s[i] = si;
    }
    __syncthreads();
    for (int i = kb + 32 + tid; i < n; i += blockDim.x) {
      double        si = s[i];
      const double *Lp = L + i + (size_t)kb * n;
#pragma unroll 8
      for (int k = 0; k < kmax; ++k) si = fma(-Lp[(size_t)k * n], s[kb + k], si);
      s[i] = si;
    }
    __syncthreads();
  }
}

// LT[i + k n] = L[k + i n]
__device__ void trsv_backward(int n, const double *__restrict__ L, const double *__restrict__ LT, double *s, double (*blk)[33])
{
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nb = (n + 31) / 32;
  (void)L;
  for (int b = nb - 1; b >= 0; --b) {
    const int kb = b * 32, kmax = min(32, n - kb);
    {
      const int r = tid & 31, c = tid >> 5;
      if (r < kmax && c < kmax) blk[r][c] = LT[(kb + r) + (size_t)(kb + c) * n]; // = L[kb+c, kb+r]
    }
    __syncthreads();
    if (warp == 0) {
      const int i  = kb + lane;
      double    si = i < n ? s[i] : 0.0;
      for (int k = kmax - 1; k >= 0; --k) {
        double xk = 0.0;
        if (lane == k) xk = si / blk[k][k];
        xk = __shfl_sync(0xffffffffu, xk, k);
        if (lane == k) si = xk;
        else if (lane < k) si = fma(-blk[lane][k], xk, si);
      }
      if (i < n) s[i] = si;
    }
    __syncthreads();
    for (int i = tid; i < kb; i += blockDim.x) {
      double        si = s[i];
      const double *Lp = LT + i + (size_t)kb * n;
#pragma unroll 8
      for (int k = kmax - 1; k >= 0; --k) si = fma(-Lp[(size_t)k * n], s[kb + k], si);
      s[i] = si;
    }
    __syncthreads();
  }
}

// mode bit 0: forward solve of `in`; bit 1: add noise and backward solve
__global__ void __launch_bounds__(CHOL_THREADS) chol_sample_kernel(int n, const double *__restrict__ L, const double *__restrict__ LT, const double *__restrict__ in, double *__restrict__ out, NoiseArgs na, int mode)
{
  __shared__ double s[CHOL_MAX_N];
  __shared__ double blk[32][33];
  for (int i = threadIdx.x; i < n; i += blockDim.x) s[i] = in[i];
  __syncthreads();
  if (mode & 1) trsv_forward(n, L, s, blk);
  if (mode & 2) {
    if (na.mode != PMG_NOISE_NONE)
      for (int i = threadIdx.x; i < n; i += blockDim.x) s[i] = __dadd_rn(s[i], noise_value(na, i));
    __syncthreads();
    trsv_backward(n, L, LT, s, blk);
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = s[i];
}
// out[i] = sum_{k in [klo(i), khi(i))} M[i*n + k] in[k]  (+ z_i); LOWER: k <= i, else k >= i.  One warp per entry.
template <bool LOWER> __global__ void __launch_bounds__(256) tri_gemv_kernel(int n, const double *__restrict__ M, const double *__restrict__ in, double *__restrict__ out, NoiseArgs na)
{
  const int i    = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();
  if (i >= n) return;
  const int     k0 = LOWER ? 0 : i, k1 = LOWER ? i + 1 : n;
  const double *row = M + (size_t)i * n;
  // the factor is a launch constant: pull this warp's row towards L2 while the preceding kernel is still running
  for (int k = k0 + lane * 16; k < k1; k += 32 * 16) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + k));
  pdl_wait();
  double        acc = 0.0;
  for (int k = k0 + lane; k < k1; k += 32) acc = fma(row[k], in[k], acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[i] = na.mode != PMG_NOISE_NONE ? __dadd_rn(acc, noise_value(na, i)) : acc;
}

__global__ void add_noise_kernel(int n, const double *__restrict__ v, double *__restrict__ out, NoiseArgs na)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __dadd_rn(v[i], noise_value(na, i));
}
} // namespace

int CholSampler::setup(pmg_ctx c, const HostCsr &a, const LrcData *lrc)
{
  ctx = c;
  n   = a.n;
  if (n > CHOL_MAX_N) PMG_FAIL(PMG_ERR_SUP, "coarsest operator has %lld rows; the dense device Cholesky sampler handles up to %d (use more levels)", (long long)n, CHOL_MAX_N);
  std::vector<double> l((size_t)(n * n), 0.0), lt((size_t)(n * n), 0.0);
  for (int64_t r = 0; r < n; ++r)
    for (int64_t k = a.rowptr[r]; k < a.rowptr[r + 1]; ++k) l[(size_t)(r + (int64_t)a.col[k] * n)] = a.val[k];
  if (lrc) { // P = A + B diag(S) B^T, assembled densely (src/pc_chols.c:119-157 forms it as a sparse product)
    for (int j = 0; j < lrc->k; ++j) {
      const double *bj = lrc->Bh.data() + (size_t)j * n;
      const double  sj = lrc->Sh[(size_t)j];
      for (int64_t cc = 0; cc < n; ++cc) {
        const double f = sj * bj[cc];
        if (f == 0.0) continue;
        for (int64_t r = 0; r < n; ++r) l[(size_t)(r + cc * n)] += bj[r] * f;
      }
    }
  }
  const int info = host_potrf_lower(n, l);
  if (info) PMG_FAIL(PMG_ERR_NOT_SPD, "Dense Cholesky failed: leading minor of order %d is not positive definite", info); // src/pc_chols.c:192
  for (int64_t i = 0; i < n; ++i)
    for (int64_t k = 0; k < n; ++k) {
      if (k > i) l[(size_t)(i + k * n)] = 0.0; // strictly upper part of the potrf workspace is not referenced
      lt[(size_t)(i + k * n)] = k >= i ? l[(size_t)(k + i * n)] : 0.0;
    }
  if (use_gemv) { // W = L^-1 by forward substitution on the columns of the identity (column sweeps: contiguous)
    std::vector<double> w((size_t)(n * n), 0.0), wt((size_t)(n * n), 0.0);
    for (int64_t j = 0; j < n; ++j) {
      double *col = w.data() + (size_t)(j * n);
      col[j]      = 1.0;
      for (int64_t k = j; k < n; ++k) {
        const double x = col[k] / l[(size_t)(k + k * n)];
        col[k]         = x;
        const double *lk = l.data() + (size_t)(k * n);
        for (int64_t i = k + 1; i < n; ++i) col[i] = std::fma(-lk[i], x, col[i]);
      }
    }
    for (int64_t i = 0; i < n; ++i)
      for (int64_t k = 0; k <= i; ++k) wt[(size_t)(k + i * n)] = w[(size_t)(i + k * n)]; // wt = row-major W
    PMG_TRY(L.upload(wt, ctx->stream));  // "L"  slot: W row-major     (rows of W contiguous: v = W b)
    PMG_TRY(LT.upload(w, ctx->stream));  // "LT" slot: W column-major  (rows of W^T contiguous: y = W^T v)
    PMG_TRY(tmp.alloc((size_t)n));
  } else {
    PMG_TRY(L.upload(l, ctx->stream));
    PMG_TRY(LT.upload(lt, ctx->stream));
  }
  PMG_TRY(vcache.alloc((size_t)n));
  PMG_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

static int launch_chol(pmg_ctx ctx, int64_t n, const double *L, const double *LT, const double *in, double *out, const NoiseArgs &na, int mode)
{
  chol_sample_kernel<<<1, CHOL_THREADS, 0, ctx->stream>>>((int)n, L, LT, in, out, na, mode);
  PMG_CUDA(cudaGetLastError());
  ctx->launches++;
  return 0;
}

template <bool LOWER> static int launch_gemv(pmg_ctx ctx, int64_t n, const double *M, const double *in, double *out, const NoiseArgs &na)
{
  PMG_CUDA(launch_pdl(ctx->stream, tri_gemv_kernel<LOWER>, dim3((unsigned)((n * 32 + 255) / 256)), dim3(256), 0, (int)n, M, in, out, na));
  PMG_CUDA(cudaGetLastError());
  ctx->launches++;
  return 0;
}

int CholSampler::sample(const double *b, double *y, const NoiseArgs &na)
{
  if (!use_gemv) return launch_chol(ctx, n, L.p, LT.p, b, y, na, 3);
  NoiseArgs none{PMG_NOISE_NONE, nullptr, 0, 0, 0};
  PMG_TRY(launch_gemv<true>(ctx, n, L.p, b, tmp.p, na)); // v = W b + z
  return launch_gemv<false>(ctx, n, LT.p, tmp.p, y, none); // y = W^T v
}
int CholSampler::forward(const double *b, double *v)
{
  NoiseArgs none{PMG_NOISE_NONE, nullptr, 0, 0, 0};
  if (!use_gemv) return launch_chol(ctx, n, L.p, LT.p, b, v, none, 1);
  return launch_gemv<true>(ctx, n, L.p, b, v, none);
}
int CholSampler::backward_noise(const double *v, double *y, const NoiseArgs &na)
{
  if (!use_gemv) return launch_chol(ctx, n, L.p, LT.p, v, y, na, 2);
  NoiseArgs none{PMG_NOISE_NONE, nullptr, 0, 0, 0};
  if (na.mode != PMG_NOISE_NONE) {
    add_noise_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((int)n, v, tmp.p, na);
    PMG_CUDA(cudaGetLastError());
    ctx->launches++;
    v = tmp.p;
  }
  return launch_gemv<false>(ctx, n, LT.p, v, y, none);
}
