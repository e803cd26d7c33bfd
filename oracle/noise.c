/* noise.c -- oracle noise sources.  TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Follows src/parmgmc.c:70-116 (VecSetRandomStandardNormal, Box-Muller branch :100-110) and
 * PETSc's default PetscRandom type "rander48" (srand48/erand48 LCG,
 * default seed 0x12345678).  The reference consumes ONE process-global stream in call order
 * (SURVEY F7); `orc_noise` is that stream.
 *
 * Three modes:
 *   tape      - injected z blocks consumed in call order (deterministic parity tests)
 *   philox    - the counter-based generator the CUDA path uses (definition below)
 *   rander48  - the reference's default generator (statistics + CPU baseline timing)
 *
 * Philox normal definition (shared with parmgmc_b200/csrc/philox.cuh):
 *   id = global row (for the matrix-free grid operators: the padded natural index, see oracle.h orc_noise.grid_*)
 *   quad q = id >> 2;  ctr = (lo32 q, hi32 q, lo32 call, hi32 call);  key = (lo32 seed, hi32 seed)
 *   (w0,w1,w2,w3) = philox4x32-10(ctr, key)
 *   rows 4q, 4q+1 use (u1,u2) = ((w0+0.5) 2^-32, (w1+0.5) 2^-32); rows 4q+2, 4q+3 use (w2, w3) likewise
 *   r = sqrt(-2 ln u1);  even row: z = r cospi(2 u2);  odd row: z = r sinpi(2 u2)
 * which is the reference's pairing (i, i+1) -> (r cos, r sin) of parmgmc.c:106-109, made
 * independent of how rows are partitioned (one generator call serves four consecutive rows).
 */
#include "oracle.h"
#include <math.h>
#include <string.h>

static inline void mulhilo(uint32_t a, uint32_t b, uint32_t *hi, uint32_t *lo)
{
  const uint64_t p = (uint64_t)a * b;
  *hi              = (uint32_t)(p >> 32);
  *lo              = (uint32_t)p;
}

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0, lo0, hi1, lo1;
    if (r) {
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
    mulhilo(0xD2511F53u, c0, &hi0, &lo0);
    mulhilo(0xCD9E8D57u, c2, &hi1, &lo1);
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* sin(pi x), cos(pi x) for x in [0,2) with exact quadrant reduction */
static void sincospi_02(double x, double *s, double *c)
{
  const double n  = nearbyint(2. * x); /* quarter turns, 0..4 */
  const double t  = x - 0.5 * n;       /* exact, |t| <= 1/4 */
  const double s0 = sin(M_PI * t), c0 = cos(M_PI * t);
  switch (((int)n) & 3) {
  case 0: *s = s0; *c = c0; break;
  case 1: *s = c0; *c = -s0; break;
  case 2: *s = -s0; *c = -c0; break;
  default: *s = -c0; *c = s0; break;
  }
}

/* the two normals of rows (2 pair, 2 pair + 1); pair = global_row >> 1 */
static void philox_pair(uint64_t seed, uint64_t call, uint64_t pair, double *zc, double *zs)
{
  const uint64_t quad   = pair >> 1;
  const uint32_t ctr[4] = {(uint32_t)quad, (uint32_t)(quad >> 32), (uint32_t)call, (uint32_t)(call >> 32)};
  const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t       w[4];
  orc_philox4x32_10(ctr, key, w);
  const int    o  = (int)(pair & 1) * 2;
  const double u1 = ((double)w[o] + 0.5) * 0x1p-32;
  const double u2 = ((double)w[o + 1] + 0.5) * 0x1p-32;
  const double r  = sqrt(-2. * log(u1));
  double       s, c;
  sincospi_02(2. * u2, &s, &c);
  *zc = r * c;
  *zs = r * s;
}

void orc_normal_philox(uint64_t seed, uint64_t call, int64_t row0, int64_t n, double *out)
{
  int64_t g = row0;
  while (g < row0 + n) {
    double zc, zs;
    philox_pair(seed, call, (uint64_t)g >> 1, &zc, &zs);
    if (g & 1) out[g++ - row0] = zs;
    else {
      out[g - row0] = zc;
      if (g + 1 < row0 + n) out[g + 1 - row0] = zs;
      g += 2;
    }
  }
}

/* rows of a grid operator: generator index = padded natural index */
void orc_normal_philox_grid(uint64_t seed, uint64_t call, int64_t row0, int64_t n, int64_t nx, int64_t pad, double *out)
{
  for (int64_t g = row0; g < row0 + n; ++g) {
    const uint64_t id = (uint64_t)(g + (g / nx) * pad);
    double         zc, zs;
    philox_pair(seed, call, id >> 1, &zc, &zs);
    out[g - row0] = (id & 1) ? zs : zc;
  }
}

/* rander48 = erand48 (PETSc rander48.c): X <- (0x5DEECE66D X + 0xB) mod 2^48, value X 2^-48 */
static double rander48_next(uint64_t *x)
{
  *x = (0x5DEECE66DULL * *x + 0xBULL) & 0xFFFFFFFFFFFFULL;
  return ldexp((double)(*x & 0xFFFF), -48) + ldexp((double)((*x >> 16) & 0xFFFF), -32) + ldexp((double)((*x >> 32) & 0xFFFF), -16);
}

void orc_noise_init_tape(orc_noise *ns, const double *tape, int64_t len)
{
  memset(ns, 0, sizeof(*ns));
  ns->mode     = 0;
  ns->tape     = tape;
  ns->tape_len = len;
}
void orc_noise_init_philox(orc_noise *ns, uint64_t seed)
{
  memset(ns, 0, sizeof(*ns));
  ns->mode = 1;
  ns->seed = seed;
}
void orc_noise_init_rander48(orc_noise *ns, uint64_t seed)
{
  memset(ns, 0, sizeof(*ns));
  ns->mode = 2;
  ns->seed = seed;
  /* PetscRandomSeed_Rander48: seed[0]=0x330e, seed[1]=low16(seed), seed[2]=next16(seed) */
  ns->x48 = 0x330EULL | ((seed & 0xFFFFULL) << 16) | (((seed >> 16) & 0xFFFFULL) << 32);
}

/* parmgmc.c:70-116: fill a whole local vector; row0 is the first global row of the block
 * (only the philox mode uses it).  Returns nonzero when an injected tape runs dry. */
int orc_noise_fill(orc_noise *ns, int64_t row0, int64_t n, double *out)
{
  if (ns->mode == 0) {
    if (ns->tape_pos + n > ns->tape_len) return 1;
    memcpy(out, ns->tape + ns->tape_pos, sizeof(double) * (size_t)n);
    ns->tape_pos += n;
  } else if (ns->mode == 1) {
    if (ns->grid_nx > 0 && ns->grid_pad > 0 && n == ns->grid_n) orc_normal_philox_grid(ns->seed, ns->call, row0, n, ns->grid_nx, ns->grid_pad, out);
    else orc_normal_philox(ns->seed, ns->call, row0, n, out);
  } else {
    for (int64_t i = 0; i < n; i += 2) { /* parmgmc.c:100-110 */
      const double u1     = rander48_next(&ns->x48);
      const double u2     = rander48_next(&ns->x48);
      const double radius = sqrt(-2.0 * log(u1));
      const double theta  = 2.0 * M_PI * u2;
      out[i]              = radius * cos(theta);
      if (i + 1 < n) out[i + 1] = radius * sin(theta);
    }
  }
  ns->call++;
  return 0;
}
