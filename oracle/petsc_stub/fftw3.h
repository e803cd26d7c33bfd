/* fftw3.h -- container-only stand-in for the three FFTW calls src/iact.c makes (TEST INFRASTRUCTURE ONLY): a plan is a
 * record of (n, in, out, sign); fftw_execute evaluates the unnormalised DFT  out[k] = sum_j in[j] exp(sign 2 pi i j k / n)
 * (FFTW's definition) with an iterative radix-2 transform (n is a power of two in iact.c) or the O(n^2) sum otherwise. */
#ifndef FFTW3_STUB_H
#define FFTW3_STUB_H
#include <complex.h>
typedef double _Complex fftw_complex;
typedef struct fftw_plan_s *fftw_plan;
#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_ESTIMATE (1U << 6)
fftw_plan fftw_plan_dft_1d(int n, fftw_complex *in, fftw_complex *out, int sign, unsigned flags);
void      fftw_execute(const fftw_plan p);
void      fftw_destroy_plan(fftw_plan p);
#endif
