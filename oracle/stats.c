/* stats.c -- oracle restatement of the estimators used to judge samplers statistically.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Follows src/iact.c:17-92, src/stats.c:8-117 and examples/ex7.c:61-91 of /root/reference.
 * FFTW (absent here) is replaced by an in-file radix-2 FFT; the transform length 2*nextpow2(n)
 * of iact.c:22-26 is a power of two, so the result is the same up to rounding.
 */
#include "oracle.h"
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

static void fft_inplace(double complex *a, int64_t n, int inverse)
{
  for (int64_t i = 1, j = 0; i < n; ++i) {
    int64_t bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) { double complex t = a[i]; a[i] = a[j]; a[j] = t; }
  }
  for (int64_t len = 2; len <= n; len <<= 1) {
    const double ang = 2 * M_PI / (double)len * (inverse ? 1 : -1);
    for (int64_t i = 0; i < n; i += len)
      for (int64_t k = 0; k < len / 2; ++k) {
        const double complex w = cos(ang * (double)k) + I * sin(ang * (double)k);
        const double complex u = a[i + k], v = a[i + k + len / 2] * w;
        a[i + k]           = u + v;
        a[i + k + len / 2] = u - v;
      }
  }
}

/* iact.c:17-46: mean removed, zero padded to 2*nextpow2(n), |FFT|^2, inverse FFT (unnormalised,
 * like FFTW), normalised by lag 0. */
void orc_autocorrelation(int64_t n, const double *x, double *acf)
{
  int64_t N = 1;
  while (N < n) N <<= 1;
  double complex *in = calloc((size_t)(2 * N), sizeof(double complex));
  double mean = 0;
  for (int64_t i = 0; i < n; ++i) mean += 1. / (double)n * x[i];
  for (int64_t i = 0; i < n; ++i) in[i] = x[i] - mean;
  fft_inplace(in, 2 * N, 0);
  for (int64_t i = 0; i < 2 * N; ++i) in[i] = in[i] * conj(in[i]);
  fft_inplace(in, 2 * N, 1);
  for (int64_t i = 0; i < n; ++i) acf[i] = creal(in[i]) / creal(in[0]);
  free(in);
}

/* iact.c:48-92: tau_i = 2 cumsum(acf)_i - 1; Sokal window: first i with i >= c tau_i, c = 5
 * (0 if none although some i < c tau_i exists; n-1 if no i < c tau_i); valid = 500 tau <= n. */
int orc_iact(int64_t n, const double *x, double *tau, double *acf_or_null, int *valid)
{
  if (n <= 1) return 1;
  double *out = malloc(sizeof(double) * (size_t)n);
  orc_autocorrelation(n, x, out);
  if (acf_or_null) memcpy(acf_or_null, out, sizeof(double) * (size_t)n);
  for (int64_t i = 1; i < n; ++i) out[i] = out[i] + out[i - 1];
  for (int64_t i = 0; i < n; ++i) out[i] = 2 * out[i] - 1;
  const int c = 5;
  int64_t   w;
  int       flag = 0;
  for (int64_t i = 0; i < n; ++i)
    if ((double)i < c * out[i]) { flag = 1; break; }
  if (flag) {
    flag = 0;
    w    = 0;
    for (int64_t i = 0; i < n; ++i)
      if ((double)i >= c * out[i]) { w = i; flag = 1; break; }
    if (!flag) w = 0;
  } else w = n - 1;
  *tau = out[w];
  if (valid) *valid = 500 * (*tau) <= (double)n;
  free(out);
  return 0;
}

/* stats.c:8-28 dense inverse by LU with natural ordering (no pivoting, like MATSOLVERPETSC on SPD) */
static int dense_inverse(int64_t n, const double *a_rowmajor, double *q)
{
  double *lu = malloc(sizeof(double) * (size_t)(n * n));
  memcpy(lu, a_rowmajor, sizeof(double) * (size_t)(n * n));
  for (int64_t k = 0; k < n; ++k) {
    if (lu[k * n + k] == 0) { free(lu); return 1; }
    for (int64_t i = k + 1; i < n; ++i) {
      lu[i * n + k] /= lu[k * n + k];
      for (int64_t j = k + 1; j < n; ++j) lu[i * n + j] -= lu[i * n + k] * lu[k * n + j];
    }
  }
  for (int64_t c = 0; c < n; ++c) {
    double *x = q + c * n; /* column c */
    for (int64_t i = 0; i < n; ++i) x[i] = i == c;
    for (int64_t i = 0; i < n; ++i)
      for (int64_t k = 0; k < i; ++k) x[i] -= lu[i * n + k] * x[k];
    for (int64_t i = n - 1; i >= 0; --i) {
      for (int64_t k = i + 1; k < n; ++k) x[i] -= lu[i * n + k] * x[k];
      x[i] /= lu[i * n + i];
    }
  }
  free(lu);
  return 0;
}

/* stats.c:94-117 EstimateCovarianceMatErrors.  samples: for each sample index i, `chains`
 * consecutive vectors of length n (index-major, stats.c:86-92).  errs[i] =
 * ||C_i - A^-1||_F / ||A^-1||_F with C_i the unbiased sample covariance (:63-84). */
int orc_cov_errors(int64_t n, const double *adense_rowmajor, int64_t chains, int64_t samples_per_chain, const double *samples, double *errs)
{
  double *q = malloc(sizeof(double) * (size_t)(n * n));
  double *c = malloc(sizeof(double) * (size_t)(n * n));
  double *m = malloc(sizeof(double) * (size_t)n);
  double *w = malloc(sizeof(double) * (size_t)n);
  if (dense_inverse(n, adense_rowmajor, q)) { free(q); free(c); free(m); free(w); return 1; }
  double qn = 0;
  for (int64_t i = 0; i < n * n; ++i) qn += q[i] * q[i];
  qn = sqrt(qn);
  for (int64_t s = 0; s < samples_per_chain; ++s) {
    const double *S = samples + s * chains * n;
    memset(m, 0, sizeof(double) * (size_t)n);
    memset(c, 0, sizeof(double) * (size_t)(n * n));
    for (int64_t k = 0; k < chains; ++k)
      for (int64_t i = 0; i < n; ++i) m[i] += 1. / (double)chains * S[k * n + i];
    for (int64_t k = 0; k < chains; ++k) {
      for (int64_t i = 0; i < n; ++i) w[i] = S[k * n + i] - m[i];
      for (int64_t j = 0; j < n; ++j)
        for (int64_t i = 0; i < n; ++i) c[j * n + i] += 1. / (double)(chains - 1) * (w[j] * w[i]);
    }
    double e = 0;
    for (int64_t i = 0; i < n * n; ++i) e += (c[i] - q[i]) * (c[i] - q[i]);
    errs[s] = sqrt(e) / qn;
  }
  free(q); free(c); free(m); free(w);
  return 0;
}

/* examples/ex7.c:61-91 GelmanRubin for a scalar QOI: samples[chain*len + j] */
double orc_gelman_rubin(int64_t unused_n, int64_t chains, int64_t len, const double *vals)
{
  (void)unused_n;
  double *means = calloc((size_t)chains, sizeof(double)), *vars = calloc((size_t)chains, sizeof(double));
  double  mean = 0, B = 0, W = 0;
  const double n = (double)len;
  for (int64_t i = 0; i < chains; ++i)
    for (int64_t j = 0; j < len; ++j) means[i] += 1. / n * vals[i * len + j];
  for (int64_t i = 0; i < chains; ++i) mean += 1. / (double)chains * means[i];
  for (int64_t i = 0; i < chains; ++i) B += n / ((double)chains - 1.) * (means[i] - mean) * (means[i] - mean);
  for (int64_t i = 0; i < chains; ++i)
    for (int64_t j = 0; j < len; ++j) vars[i] += 1. / (n - 1.) * (vals[i * len + j] - means[i]) * (vals[i * len + j] - means[i]);
  for (int64_t i = 0; i < chains; ++i) W += 1. / (double)chains * vars[i];
  const double gr = ((n - 1.) / n * W + 1. / n * B) / W;
  free(means); free(vars);
  return gr;
}
