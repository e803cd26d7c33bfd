/* sampler.c -- oracle restatement of the Gibbs samplers and the dense Cholesky sampler.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Follows src/pc_mcgibbs.c:119-128, :142-153, :155-188; src/pc_sorgibbs.c:76-134;
 * src/pc_chols.c:173-195, :220-291 of /root/reference.
 */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* pc_mcgibbs.c:142-153: sqrtdiag = sqrt(|a_ii|) (VecSqrtAbs), then scaled by sqrt((2-omega)/omega)
 * (VecScale).  pc_sorgibbs.c:235-236 is the omega = 1 case (scale factor exactly 1). */
void orc_sqrtdiag(int64_t n, const double *val, const int64_t *diagptr, double omega, double *sqrtdiag)
{
  const double f = sqrt((2 - omega) / omega);
  for (int64_t r = 0; r < n; ++r) {
    const double s = sqrt(fabs(val[diagptr[r]]));
    sqrtdiag[r]    = s * f;
  }
}

/* pc_mcgibbs.c:119-128 PrepareRHS_Default / pc_sorgibbs.c:81-83:
 * w = z (fresh N(0,1)); w = w .* sqrtdiag (VecPointwiseMult); w = w + 1*b (VecAXPY). */
void orc_prepare_rhs(int64_t n, const double *b, const double *sqrtdiag, const double *z, double *w)
{
  for (int64_t r = 0; r < n; ++r) {
    const double t = z[r] * sqrtdiag[r];
    w[r]           = b ? t + b[r] : t;
  }
}

/* pc_mcgibbs.c:155-188 PCApplyRichardson_MulticolorGibbs: `its` samples; forward/backward =
 * (prepare_rhs; MCSORApply); symmetric = forward sweep with fresh noise then backward sweep with
 * fresh noise (:172-182); callback after every sample (:183).  Ignores tolerances / guesszero.
 * With omega = 1, one colour, forward this is also PCApplyRichardson_SORGibbs
 * (pc_sorgibbs.c:115-134 through MatSOR). */
int orc_gibbs_richardson(int64_t n, const int64_t *rowptr, const int32_t *col, const double *val, double omega, int ncolors, const int64_t *colorptr, const int32_t *colorrows, int type, orc_noise *ns, const double *b, double *y, int64_t its, orc_sample_cb cb, void *cbctx)
{
  int64_t *diagptr  = malloc(sizeof(int64_t) * (size_t)n);
  double  *idiag    = malloc(sizeof(double) * (size_t)n);
  double  *sqrtdiag = malloc(sizeof(double) * (size_t)n);
  double  *w        = malloc(sizeof(double) * (size_t)n);
  double  *z        = malloc(sizeof(double) * (size_t)n);
  int      err      = 0;
  if (orc_diag_ptrs(n, rowptr, col, diagptr)) { err = 2; goto done; }
  orc_idiag(n, val, diagptr, omega, idiag);
  orc_sqrtdiag(n, val, diagptr, omega, sqrtdiag);
  for (int64_t it = 0; it < its && !err; ++it) {
    if (type == ORC_SOR_FORWARD || type == ORC_SOR_BACKWARD) {
      if ((err = orc_noise_fill(ns, 0, n, z))) break;
      orc_prepare_rhs(n, b, sqrtdiag, z, w);
      orc_sweep_seq(n, rowptr, col, val, diagptr, idiag, omega, ncolors, colorptr, colorrows, type, w, y);
    } else {
      if ((err = orc_noise_fill(ns, 0, n, z))) break;
      orc_prepare_rhs(n, b, sqrtdiag, z, w);
      orc_sweep_seq(n, rowptr, col, val, diagptr, idiag, omega, ncolors, colorptr, colorrows, ORC_SOR_FORWARD, w, y);
      if ((err = orc_noise_fill(ns, 0, n, z))) break;
      orc_prepare_rhs(n, b, sqrtdiag, z, w);
      orc_sweep_seq(n, rowptr, col, val, diagptr, idiag, omega, ncolors, colorptr, colorrows, ORC_SOR_BACKWARD, w, y);
    }
    if (cb) err = cb(it, y, cbctx);
  }
done:
  free(diagptr); free(idiag); free(sqrtdiag); free(w); free(z);
  return err;
}

/* ---- dense Cholesky sampler, pc_chols.c:173-195 (potrf "L", column major) ---------------- */
int orc_potrf_lower(int64_t n, double *a)
{
  for (int64_t j = 0; j < n; ++j) {
    double d = a[j + j * n];
    for (int64_t k = 0; k < j; ++k) d = fma(-a[j + k * n], a[j + k * n], d);
    if (!(d > 0)) return (int)(j + 1);
    d            = sqrt(d);
    a[j + j * n] = d;
    for (int64_t i = j + 1; i < n; ++i) {
      double s = a[i + j * n];
      for (int64_t k = 0; k < j; ++k) s = fma(-a[i + k * n], a[j + k * n], s);
      a[i + j * n] = s / d;
    }
  }
  return 0;
}

/* pc_chols.c:220-260: trsv("L","N","N") forward / trsv("L","T","N") backward, in place.
 * Accumulation order per unknown follows reference-BLAS dtrsv: forward x_i collects k = 0..i-1
 * ascending; transposed x_i collects k = n-1..i+1 descending (dtrsv 'T' loop "DO I = N,J+1,-1"). */
void orc_trsv_lower(int64_t n, const double *l, int trans, double *x)
{
  if (!trans) {
    for (int64_t i = 0; i < n; ++i) {
      double s = x[i];
      for (int64_t k = 0; k < i; ++k) s = fma(-l[i + k * n], x[k], s);
      x[i] = s / l[i + i * n];
    }
  } else {
    for (int64_t i = n - 1; i >= 0; --i) {
      double s = x[i];
      for (int64_t k = n - 1; k > i; --k) s = fma(-l[k + i * n], x[k], s);
      x[i] = s / l[i + i * n];
    }
  }
}

/* pc_chols.c:284-287: v = L^-1 b; v += z; y = L^-T v   =>  y ~ N(A^-1 b, A^-1) */
int orc_chol_sample(int64_t n, const double *l, orc_noise *ns, const double *b, double *y)
{
  double *z = malloc(sizeof(double) * (size_t)n);
  memcpy(y, b, sizeof(double) * (size_t)n);
  orc_trsv_lower(n, l, 0, y);
  const int err = orc_noise_fill(ns, 0, n, z);
  if (!err) {
    for (int64_t i = 0; i < n; ++i) y[i] = y[i] + z[i];
    orc_trsv_lower(n, l, 1, y);
  }
  free(z);
  return err;
}
