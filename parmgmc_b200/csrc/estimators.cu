// estimators.cu -- the statistics a sampling run is judged by, kept on the device.
//
// Reference: examples/benchmark/main.cc:151-175 (SaveSample: qoi = <y, meas_vec> per sample, optional Welford mean /
// variance of the whole field), src/iact.c:17-92 (Autocorrelation by FFT, IACT with Sokal's automatic window).
// The reference does this in a host callback, which forces every sample back to the host; here one kernel per
// sample reads y once (dot product partials + the Welford update) and the autocorrelation runs on cuFFT (bound at run
// time like NCCL, so the library has no link-time dependency on it).
#include <cufft.h>
#include <dlfcn.h>

#include <cmath>

#include "common.hpp"

namespace {
constexpr int Q_CHUNK = 4096, Q_THREADS = 256, Q_PER_THREAD = Q_CHUNK / Q_THREADS;

// partial[chunk] = sum over the chunk of meas * y; optionally the Welford update of (mean, M2) with sample number i
__global__ void __launch_bounds__(Q_THREADS) qoi_partial_kernel(int64_t n, const double *__restrict__ meas, const double *__restrict__ y, double *__restrict__ partial, double *__restrict__ mean, double *__restrict__ M2, double inv_i)
{
  __shared__ double red[Q_THREADS / 32];
  const int64_t     r0 = (int64_t)blockIdx.x * Q_CHUNK;
  double            acc = 0.0;
#pragma unroll
  for (int q = 0; q < Q_PER_THREAD; ++q) {
    const int64_t r = r0 + threadIdx.x + (int64_t)q * Q_THREADS;
    if (r >= n) continue;
    const double yv = y[r];
    acc             = fma(meas[r], yv, acc);
    if (mean) { // main.cc:161-170
      const double delta = yv - mean[r];
      const double m     = fma(inv_i, delta, mean[r]);
      mean[r]            = m;
      M2[r]              = fma(yv - m, delta, M2[r]);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < Q_THREADS / 32; ++w) s += red[w];
    partial[blockIdx.x] = s;
  }
}
__global__ void qoi_finish_kernel(int nchunks, const double *__restrict__ partial, double *__restrict__ out)
{
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0;
    for (int c = 0; c < nchunks; ++c) s += partial[c];
    *out = s;
  }
}

// ---- cuFFT, bound at run time ----
struct CufftApi {
  decltype(&cufftPlan1d)    Plan1d    = nullptr;
  decltype(&cufftExecZ2Z)   ExecZ2Z   = nullptr;
  decltype(&cufftDestroy)   Destroy   = nullptr;
  decltype(&cufftSetStream) SetStream = nullptr;
  bool                      ok        = false;
};
CufftApi &cufft()
{
  static CufftApi api = [] {
    CufftApi a;
    void    *h = nullptr;
    for (const char *nm : {"libcufft.so.11", "libcufft.so.12", "libcufft.so"})
      if ((h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!h) return a;
    a.Plan1d    = (decltype(a.Plan1d))dlsym(h, "cufftPlan1d");
    a.ExecZ2Z   = (decltype(a.ExecZ2Z))dlsym(h, "cufftExecZ2Z");
    a.Destroy   = (decltype(a.Destroy))dlsym(h, "cufftDestroy");
    a.SetStream = (decltype(a.SetStream))dlsym(h, "cufftSetStream");
    a.ok        = a.Plan1d && a.ExecZ2Z && a.Destroy && a.SetStream;
    return a;
  }();
  return api;
}

__global__ void acf_load_kernel(int64_t n, int64_t len, const double *__restrict__ x, double mean, cufftDoubleComplex *__restrict__ buf)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= len) return;
  buf[i].x = i < n ? x[i] - mean : 0.0; // src/iact.c:29-30: mean removed, zero padded to 2 nextpow2(n)
  buf[i].y = 0.0;
}
__global__ void acf_power_kernel(int64_t len, cufftDoubleComplex *__restrict__ buf)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= len) return;
  const double re = buf[i].x, im = buf[i].y;
  buf[i].x = re * re + im * im; // out * conj(out), src/iact.c:36
  buf[i].y = 0.0;
}
__global__ void acf_norm_kernel(int64_t n, const cufftDoubleComplex *__restrict__ buf, double *__restrict__ acf)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) acf[i] = buf[i].x / buf[0].x; // src/iact.c:42
}
} // namespace

int QoiState::init(pmg_ctx c, int64_t n_, const double *meas_host, int64_t capacity, bool with_mean_var)
{
  ctx = c;
  n   = n_;
  cap = capacity;
  count = nseen = 0;
  welford = with_mean_var;
  nchunks = (int)((n + Q_CHUNK - 1) / Q_CHUNK);
  PMG_TRY(meas.upload(meas_host, (size_t)n, ctx->stream));
  PMG_TRY(trace.alloc((size_t)std::max<int64_t>(1, cap)));
  PMG_TRY(partial.alloc((size_t)nchunks));
  if (welford) {
    PMG_TRY(mean.alloc((size_t)n));
    PMG_TRY(M2.alloc((size_t)n));
    PMG_TRY(mean.zero(ctx->stream));
    PMG_TRY(M2.zero(ctx->stream));
  } else {
    mean.release();
    M2.release();
  }
  PMG_CUDA(cudaStreamSynchronize(ctx->stream));
  on = true;
  return 0;
}

// SaveSample (examples/benchmark/main.cc:151-175) for the sample held in y (device)
int QoiState::accumulate(const double *y)
{
  if (!on) return 0;
  if (count >= cap) PMG_FAIL(PMG_ERR_ARG, "QOI trace is full (%lld samples): read it with pmg_pc_get_qoi or enlarge it", (long long)cap);
  ++nseen;
  qoi_partial_kernel<<<(unsigned)nchunks, Q_THREADS, 0, ctx->stream>>>(n, meas.p, y, partial.p, welford ? mean.p : nullptr, welford ? M2.p : nullptr, 1.0 / (double)nseen);
  qoi_finish_kernel<<<1, 32, 0, ctx->stream>>>(nchunks, partial.p, trace.p + count);
  PMG_CUDA(cudaGetLastError());
  ctx->launches += 2;
  ++count;
  return 0;
}

// Autocorrelation (src/iact.c:17-46) of a host series, on the device
int device_autocorrelation(pmg_ctx ctx, int64_t n, const double *x_host, double *acf_host)
{
  if (n < 1) PMG_FAIL(PMG_ERR_ARG, "autocorrelation of an empty series");
  if (!cufft().ok) PMG_FAIL(PMG_ERR_SUP, "libcufft could not be loaded: %s", dlerror());
  int64_t N = 1;
  while (N < n) N <<= 1;
  const int64_t len = 2 * N;
  if (len > (int64_t)1 << 30) PMG_FAIL(PMG_ERR_SUP, "series too long for one cuFFT plan");
  double mean = 0.0;
  for (int64_t i = 0; i < n; ++i) mean += 1. / (double)n * x_host[i]; // src/iact.c:28
  DevBuf<double>             x, acf;
  DevBuf<cufftDoubleComplex> buf;
  PMG_TRY(x.upload(x_host, (size_t)n, ctx->stream));
  PMG_TRY(acf.alloc((size_t)n));
  PMG_TRY(buf.alloc((size_t)len));
  const unsigned gb = (unsigned)((len + 255) / 256);
  acf_load_kernel<<<gb, 256, 0, ctx->stream>>>(n, len, x.p, mean, buf.p);
  cufftHandle plan;
  if (cufft().Plan1d(&plan, (int)len, CUFFT_Z2Z, 1) != CUFFT_SUCCESS) PMG_FAIL(PMG_ERR_CUDA, "cufftPlan1d failed");
  cufft().SetStream(plan, ctx->stream);
  bool ok = cufft().ExecZ2Z(plan, buf.p, buf.p, CUFFT_FORWARD) == CUFFT_SUCCESS;
  acf_power_kernel<<<gb, 256, 0, ctx->stream>>>(len, buf.p);
  ok = ok && cufft().ExecZ2Z(plan, buf.p, buf.p, CUFFT_INVERSE) == CUFFT_SUCCESS; // unnormalised, like FFTW (the ratio cancels it)
  acf_norm_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, buf.p, acf.p);
  ctx->launches += 3;
  cudaError_t e = cudaMemcpyAsync(acf_host, acf.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cufft().Destroy(plan);
  if (!ok || e != cudaSuccess) PMG_FAIL(PMG_ERR_CUDA, "autocorrelation: cuFFT / copy failed");
  return 0;
}

// IACT (src/iact.c:48-92): tau_i = 2 cumsum(acf)_i - 1, Sokal's window with c = 5, valid = 500 tau <= n
int device_iact(pmg_ctx ctx, int64_t n, const double *x_host, double *tau, double *acf_or_null, int *valid)
{
  if (n <= 1) PMG_FAIL(PMG_ERR_ARG, "Too few data points"); // src/iact.c:79
  std::vector<double> out((size_t)n);
  PMG_TRY(device_autocorrelation(ctx, n, x_host, out.data()));
  if (acf_or_null) std::memcpy(acf_or_null, out.data(), sizeof(double) * (size_t)n);
  for (int64_t i = 1; i < n; ++i) out[(size_t)i] = out[(size_t)i] + out[(size_t)i - 1];
  for (int64_t i = 0; i < n; ++i) out[(size_t)i] = 2 * out[(size_t)i] - 1;
  const int c    = 5;
  bool      flag = false;
  int64_t   w    = n - 1;
  for (int64_t i = 0; i < n; ++i)
    if ((double)i < c * out[(size_t)i]) {
      flag = true;
      break;
    }
  if (flag) {
    w = 0;
    for (int64_t i = 0; i < n; ++i)
      if ((double)i >= c * out[(size_t)i]) {
        w = i;
        break;
      }
  }
  *tau = out[(size_t)w];
  if (valid) *valid = 500 * (*tau) <= (double)n;
  return 0;
}
