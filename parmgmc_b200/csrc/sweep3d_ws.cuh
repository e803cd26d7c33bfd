// sweep3d_ws.cuh -- the warp-specialised fused 3D red-black Gibbs sweep of sweep3d.cuh as a PERSISTENT kernel (device
// Philox noise only; K1 of SURVEY 8(d), src/pc_mcgibbs.c:119-128 + src/mc_sor.c:257-271 in one pass over memory).
//
// Same tiles, same rolling window, same arithmetic per node (update<> of sweep3d.cuh, fma for fma) -- what changes is
// everything around the arithmetic, which was two thirds of the stencil warps' instruction stream (profiles/r2_summary.md):
//  * one CTA per SM walks a work queue (atomic counter, reset by the CTA that draws the last ticket): the Box-Muller tables,
//    the coefficient table and the mbarriers are set up once, not once per tile; the mbarrier phases carry over from tile
//    to tile in a per-thread bit mask;
//  * the plane steps are unrolled in fours with the ring slot, the publish / noise buffer and the row parity as
//    compile-time constants: every shared-memory access of a step is [register + immediate], the registers (three
//    addresses per thread) are set once per tile; the output pointer and the generator counter are running values;
//  * the noise producers are decoupled from the stencil warps' plane barrier: producers and stencil warps meet on two
//    pairs of named barriers (full / empty per noise buffer, the bar.arrive / bar.sync producer-consumer pattern), so the
//    producers run up to two planes ahead and their FP64 chains fill the issue slots the stencil warps leave while they
//    wait for each other; the stencil warps synchronise among themselves on a 16-warp named barrier;
//  * a row publishes only the two columns its phase A updated (the only ones the neighbouring rows read).
// Results are bit-identical to sweep3d_kernel and to the colour-by-colour path (tests/test_gpu_parity.py).
#pragma once
#include "sweep3d.cuh"

namespace sweep3d {

constexpr int WS_NW = 16, WS_SX = 4, WS_SB = 2;
enum { BAR_STENCIL = 1, BAR_FULL = 2, BAR_EMPTY = 4 }; // named barriers: 1, 2-3 (noise buffer filled), 4-5 (noise buffer read)

__device__ __forceinline__ void nbar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void nbar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void lds128(uint32_t addr, double &v0, double &v1) { asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v0), "=d"(v1) : "r"(addr)); }

// both colour phases of one plane step for row parity P; the addresses are this thread's:
//   xa / xb : its 16-byte halves (columns 0-1 / 2-3) of its own row in the x box of plane kk (swizzle applied)
//   nr      : its 8-byte slot of its own row in the published buffer of plane kk-1
template <int P, bool INTERIOR, int LPR, int RB>
__device__ __forceinline__ void phases_ws(const Args &a, const Coef *coef, bool inner_row, const double (&xm2)[4], double (&xm1)[4], double (&x0)[4], const double (&xp1)[4], uint32_t xa, uint32_t xb, uint32_t nr, const double (&w)[4],
                                          const double (&wk)[2], const int (&ci)[4], const int (&cis)[4])
{
  { // phase A: first colour of plane kk (columns P, P+2), every neighbour still old
    const double s0 = lds64(xa - RB + 8 * P), n0 = lds64(xa + RB + 8 * P), s1 = lds64(xb - RB + 8 * P), n1 = lds64(xb + RB + 8 * P);
    const double west = P == 0 ? shfl_up1(x0[3]) : 0.0, east = P == 1 ? shfl_dn1(x0[0]) : 0.0;
    update<P, INTERIOR>(a, coef, x0, xm1[P], s0, n0, xp1[P], west, east, w[P], ci[P]);
    update<P + 2, INTERIOR>(a, coef, x0, xm1[P + 2], s1, n1, xp1[P + 2], west, east, w[P + 2], ci[P + 2]);
  }
  { // phase B: second colour of plane kk-1 (the same columns), every neighbour new; halo rows take part in the shuffles only
    const double west = P == 0 ? shfl_up1(xm1[3]) : 0.0, east = P == 1 ? shfl_dn1(xm1[0]) : 0.0;
    if (inner_row) {
      const double s0 = lds64(nr - RB), n0 = lds64(nr + RB), s1 = lds64(nr - RB + LPR * 8), n1 = lds64(nr + RB + LPR * 8);
      update<P, INTERIOR>(a, coef, xm1, xm2[P], s0, n0, x0[P], west, east, wk[0], cis[P]);
      update<P + 2, INTERIOR>(a, coef, xm1, xm2[P + 2], s1, n1, x0[P + 2], west, east, wk[1], cis[P + 2]);
    }
  }
}

// one tile.  xph / bph: current phase bit of every slot of the x / b ring (stencil threads), carried from tile to tile
template <bool PRODUCER, bool INTERIOR, bool ZCONST, int LPR, typename FT>
__device__ __forceinline__ void ws_item(const Args &a, const FT &ft, const Coef *coef, uint32_t sm, const Item it, uint32_t &xph, uint32_t &bph)
{
  constexpr int NW = WS_NW, SX = WS_SX, SB = WS_SB;
  using L = Smem<NW, SX, SB, true>;
  static_assert(SX == 4 && SB == 2, "the unrolled steps assume a 4-slot x ring and a 2-slot b ring");
  static_assert(!(INTERIOR && (ZCONST || LPR != 32)), "interior tiles are full-width and need no classes");
  constexpr int  RPW = 32 / LPR, ROWS = NW * RPW, RB = LPR * 32; // grid rows per warp, rows of the tile (2 of them halo), bytes per row
  constexpr int  NT = NW * 64;                                   // threads of the CTA
  const int      wlane = threadIdx.x & 31, lane = wlane % LPR, warp = threadIdx.x >> 5, wi = warp % NW;
  const int      w = RPW == 1 ? wi : 4 * (wi >> 1) + (wi & 1) + 2 * (wlane / LPR);
  const int      c0 = it.strip * STRIP_OUT - 4, c = c0 + 4 * lane;
  const int      y = it.ya - 1 + w;
  const int      K0 = it.ka - 1, nsteps = it.kb - it.ka + 2, nq = nsteps + 2; // plane steps K0 .. kb; x planes K0-1 .. kb+1
  const uint32_t bar_x = sm + (uint32_t)L::off_bar, bar_b = bar_x + SX * 8;

  // column / row parts of the node classification (edge tiles)
  int        colmiss[4] = {0, 0, 0, 0};
  bool       colok[4]   = {true, true, true, true};
  const bool rowok   = y >= 0 && y < a.ny;
  const int  rowmiss = (y == 0 ? 1 : 0) + (y == a.ny - 1 ? 1 : 0);
  if (!INTERIOR) {
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      colok[m]   = c + m >= 0 && c + m < a.nx;
      colmiss[m] = (c + m == 0 ? 1 : 0) + (c + m == a.nx - 1 ? 1 : 0);
    }
  }
  int ci0[4] = {0, 0, 0, 0}; // ZCONST: the classes of every plane
  if (ZCONST) {
#pragma unroll
    for (int m = 0; m < 4; ++m) ci0[m] = (rowok && colok[m]) ? 6 - colmiss[m] - rowmiss : 7;
  }
  auto classes = [&](int kk, int (&ci)[4]) {
    if (INTERIOR) return;
    if (ZCONST) {
#pragma unroll
      for (int m = 0; m < 4; ++m) ci[m] = ci0[m];
      return;
    }
    const bool kok   = kk >= 0 && kk < a.nz;
    const int  kmiss = (kk == 0 ? 1 : 0) + (kk == a.nz - 1 ? 1 : 0);
#pragma unroll
    for (int m = 0; m < 4; ++m) ci[m] = (kok && rowok && colok[m]) ? 6 - colmiss[m] - rowmiss - kmiss : 7;
  };
  const uint32_t rowslot = (uint32_t)(w * RB + lane * 8); // this thread's first 8-byte slot in a [4][LPR] row of the published / noise buffers

  if (PRODUCER) { // ---- noise producers: sqrtdiag * z of step s goes to buffer s & 1, up to two steps ahead of its use ----
    const uint32_t  zr    = sm + (uint32_t)L::off_z + rowslot;
    const long long qstep = ((long long)a.ny * a.pitch) >> 2;
    long long       quad  = (((long long)K0 * a.ny + y) * a.pitch + c) >> 2; // generator counter of (c .. c+3, y, K0): philox.cuh
    auto produce = [&](auto btag, int s) {
      constexpr int B = decltype(btag)::value;
      int           ci[4] = {0, 0, 0, 0};
      classes(K0 + s, ci);
      uint32_t w0, w1, w2, w3;
      double   z[4];
      philox4x32_10_keys((uint32_t)quad, (uint32_t)((unsigned long long)quad >> 32), a.call_lo, a.call_hi, a.pk, w0, w1, w2, w3);
      quad += qstep;
      fastnormal::box_muller(ft, w0, w1, z[0], z[1]);
      fastnormal::box_muller(ft, w2, w3, z[2], z[3]);
#pragma unroll
      for (int m = 0; m < 4; ++m) z[m] = __dmul_rn(z[m], INTERIOR ? a.sd : coef[ci[m]].sd); // first rounding of w = b + sqrtdiag z (src/pc_mcgibbs.c:124-126)
      if (s >= 2) nbar_sync(BAR_EMPTY + B, NT); // the stencil warps have read this buffer's step s-2
#pragma unroll
      for (int m = 0; m < 4; ++m) sts64(zr + B * L::NEWBUF + m * (LPR * 8), z[m]);
      nbar_arrive(BAR_FULL + B, NT);
    };
    int s = 0;
    for (; s + 1 < nsteps; s += 2) {
      produce(std::integral_constant<int, 0>{}, s);
      produce(std::integral_constant<int, 1>{}, s + 1);
    }
    if (s < nsteps) produce(std::integral_constant<int, 0>{}, s);
    return;
  }

  // ---- stencil warps ----
  const bool     inner_row  = w >= 1 && w <= ROWS - 2;
  const bool     out_thread = inner_row && lane >= 1 && lane <= LPR - 2 && (INTERIOR || (c < a.nx && y < a.ny));
  const uint32_t xbytes = (ROWS + 2) * RB, bbytes = ROWS * RB;
  const CUtensorMap *tmx = LPR == 32 ? &a.tm_x : &a.tm_x16, *tmb = LPR == 32 ? &a.tm_b : &a.tm_b16;
  // x sequence q = 0, 1, ...: plane K0 - 1 + q, slot q % 4;  b sequence r = 0, 1, ...: plane K0 + r, slot r % 2
  auto issue_x = [&](int q, int slot) {
    const uint32_t bar = bar_x + slot * 8;
    mbar_expect_tx(bar, xbytes);
    tma_load_4d(sm + (uint32_t)L::off_x + slot * L::XSTAGE, tmx, 0, c0 >> 2, it.ya - 2, K0 - 1 + q - a.tlo, bar);
  };
  auto issue_b = [&](int r, int slot) {
    const uint32_t bar = bar_b + slot * 8;
    mbar_expect_tx(bar, bbytes);
    tma_load_4d(sm + (uint32_t)L::off_b + slot * L::BSTAGE, tmb, 0, c0 >> 2, it.ya - 1, K0 + r - a.tlo, bar);
  };
  if (threadIdx.x == 0) {
#pragma unroll
    for (int q = 0; q < SX; ++q)
      if (q < nq) issue_x(q, q);
    if (a.has_b) {
#pragma unroll
      for (int r = 0; r < SB; ++r)
        if (r < nsteps) issue_b(r, r);
    }
  }
  auto wait_x = [&](int slot) {
    mbar_wait(bar_x + slot * 8, (xph >> slot) & 1u);
    xph ^= 1u << slot;
  };
  auto wait_b = [&](int slot) {
    mbar_wait(bar_b + slot * 8, (bph >> slot) & 1u);
    bph ^= 1u << slot;
  };
  // the three addresses every shared-memory access of a step is an immediate away from
  const uint32_t swzb = sweep2d::lane_swz(lane);
  const uint32_t xa = sm + (uint32_t)L::off_x + (uint32_t)((w + 1) * RB) + sweep2d::lane_seg(lane) + swzb;          // columns 0, 1 of the own row, slot 0 of the x ring
  const uint32_t xb = sm + (uint32_t)L::off_x + (uint32_t)((w + 1) * RB) + sweep2d::lane_seg(lane) + (16u - swzb);  // columns 2, 3
  const uint32_t nr = sm + (uint32_t)L::off_new + rowslot; // published buffer 0; the noise buffers are L::off_z - L::off_new further, the b ring L::off_b - L::off_x - RB from xa / xb
  constexpr uint32_t BOFF = (uint32_t)(L::off_b - L::off_x) - RB, ZOFF = (uint32_t)(L::off_z - L::off_new);

  double A[4] = {0, 0, 0, 0}, B[4], C[4], D[4], wk[2] = {0, 0}; // rolling window: planes kk-2, kk-1, kk and the incoming kk+1
  int    cis[4] = {ZCONST ? ci0[0] : 7, ZCONST ? ci0[1] : 7, ZCONST ? ci0[2] : 7, ZCONST ? ci0[3] : 7};
  double *outp = a.xout + ((long long)(K0 - 1 - a.tlo) * a.pplane + (long long)y * a.pitch + c); // plane kk-1 of step 0; only dereferenced by out_thread from step 2 on

  wait_x(0);
  lds128(xa, B[0], B[1]);
  lds128(xb, B[2], B[3]);
  wait_x(1);
  lds128(xa + L::XSTAGE, C[0], C[1]);
  lds128(xb + L::XSTAGE, C[2], C[3]);
  // step 0 reads buffer 1 as "the published plane K0-1": its phase B result is never stored, any finite values will do
  sts64(nr + L::NEWBUF, B[0]);
  sts64(nr + L::NEWBUF + LPR * 8, B[2]);
  nbar_sync(BAR_STENCIL, NW * 32);
  if (threadIdx.x == 0 && SX < nq) issue_x(SX, 0); // slot of q = 0 is free again

  // one plane step, the s-th of the tile with s % 4 == J, for a row whose first-colour columns of plane kk = K0 + s are
  // M = P, P+2; xp1 receives plane kk+1
  auto step = [&](auto ptag, auto jtag, int s, const double (&xm2)[4], double (&xm1)[4], double (&x0)[4], double (&xp1)[4]) {
    constexpr int P = decltype(ptag)::value, J = decltype(jtag)::value;
    constexpr int S0 = (J + 1) % 4, S1 = (J + 2) % 4; // ring slots of planes kk, kk+1
    double bb[4] = {0, 0, 0, 0};
    wait_x(S1);
    lds128(xa + S1 * L::XSTAGE, xp1[0], xp1[1]);
    lds128(xb + S1 * L::XSTAGE, xp1[2], xp1[3]);
    if (a.has_b) {
      wait_b(J % 2);
      lds128(xa + BOFF + (J % 2) * L::BSTAGE, bb[0], bb[1]);
      lds128(xb + BOFF + (J % 2) * L::BSTAGE, bb[2], bb[3]);
    }
    int ci[4] = {0, 0, 0, 0};
    classes(K0 + s, ci);
    // w = b + sqrtdiag z (src/pc_mcgibbs.c:124-126: two roundings; the first one is the producer's)
    double wv[4];
    nbar_sync(BAR_FULL + (J & 1), NT);
#pragma unroll
    for (int m = 0; m < 4; ++m) wv[m] = lds64(nr + ZOFF + (J & 1) * L::NEWBUF + m * (LPR * 8));
    if (s + 2 < nsteps) nbar_arrive(BAR_EMPTY + (J & 1), NT);
#pragma unroll
    for (int m = 0; m < 4; ++m) wv[m] = __dadd_rn(wv[m], bb[m]);
    phases_ws<P, INTERIOR, LPR, RB>(a, coef, inner_row, xm2, xm1, x0, xp1, xa + S0 * L::XSTAGE, xb + S0 * L::XSTAGE, nr + ((J + 1) & 1) * L::NEWBUF, wv, wk, ci, cis);
    wk[0] = wv[1 - P];
    wk[1] = wv[3 - P];
#pragma unroll
    for (int m = 0; m < 4; ++m) cis[m] = ci[m];
    if (out_thread && s >= 2) st256(outp, xm1); // plane kk-1 of this row is final; planes ka .. kb-1 are steps 2 .. nsteps-1
    outp += a.pplane;
    sts64(nr + (J & 1) * L::NEWBUF, x0[P]); // the half-updated plane kk, for the neighbouring rows
    sts64(nr + (J & 1) * L::NEWBUF + LPR * 8, x0[P + 2]);
    nbar_sync(BAR_STENCIL, NW * 32);
    if (threadIdx.x == 0) { // the boxes of plane kk (x) and of this step (b) are free
      if (s + 1 + SX < nq) issue_x(s + 1 + SX, S0);
      if (a.has_b && s + SB < nsteps) issue_b(s + SB, J % 2);
    }
  };
  // the row parity alternates from plane to plane, the rolling window has four slots and so has the x ring: steps are
  // unrolled in fours; every stencil warp of the CTA runs the same number of steps (one barrier each)
  auto run = [&](auto qtag) {
    constexpr int Q = decltype(qtag)::value;
    using T0 = std::integral_constant<int, Q>;
    using T1 = std::integral_constant<int, 1 - Q>;
    using J0 = std::integral_constant<int, 0>;
    using J1 = std::integral_constant<int, 1>;
    using J2 = std::integral_constant<int, 2>;
    using J3 = std::integral_constant<int, 3>;
    int s = 0;
    for (; s + 3 < nsteps; s += 4) {
      step(T0{}, J0{}, s, A, B, C, D);
      step(T1{}, J1{}, s + 1, B, C, D, A);
      step(T0{}, J2{}, s + 2, C, D, A, B);
      step(T1{}, J3{}, s + 3, D, A, B, C);
    }
    if (s < nsteps) {
      step(T0{}, J0{}, s, A, B, C, D);
      if (s + 1 < nsteps) {
        step(T1{}, J1{}, s + 1, B, C, D, A);
        if (s + 2 < nsteps) step(T0{}, J2{}, s + 2, C, D, A, B);
      }
    }
  };
  if (((y + K0 + a.flip) & 1) == 0) run(std::integral_constant<int, 0>{});
  else run(std::integral_constant<int, 1>{});
}

// the tile loop of one role: tickets are drawn by thread 0, one tile ahead (the atomic's latency hides behind a tile).  A
// ticket t < nitems is tile t; the nitems + gridDim.x-th draw of the launch is the last one and resets the counter.
__device__ __forceinline__ void ws_draw(const Args &a, int *slot)
{
  const unsigned t = atomicAdd(&a.queue->next, 1u);
  if (t == (unsigned)a.nitems + gridDim.x - 1u) a.queue->next = 0u;
  *slot = (int)t;
}
template <bool PRODUCER, typename FT> __device__ __forceinline__ void ws_tiles(const Args &a, const FT &ft, const Coef *coef, uint32_t sm, int *s_ticket)
{
  constexpr int NW = WS_NW;
  uint32_t      xph = 0, bph = 0;
  for (int cur = 0;; cur ^= 1) {
    const int t = s_ticket[cur];
    if (t >= a.nitems) break;
    if (threadIdx.x == 0) ws_draw(a, s_ticket + (cur ^ 1));
    const Item it = a.items[t];
    const int  c0 = it.strip * STRIP_OUT - 4;
    // zconst: every plane the tile touches has both z neighbours; interior: moreover every node the CTA updates has all six
    // neighbours, and every row / plane it reads exists and is owned
    const bool zconst   = it.ka - 1 >= 1 && it.kb <= a.nz - 2 && it.ka - 2 >= a.tlo && it.kb + 1 < a.thi;
    const bool interior = zconst && !it.narrow && c0 >= 1 && c0 + 127 <= a.nx - 2 && it.ya - 1 >= 1 && it.ya + NW - 2 <= a.ny - 2;
    if (interior) ws_item<PRODUCER, true, false, 32>(a, ft, coef, sm, it, xph, bph);
    else if (it.narrow) ws_item<PRODUCER, false, true, 16>(a, ft, coef, sm, it, xph, bph); // the host makes narrow tiles only where zconst holds
    else if (zconst) ws_item<PRODUCER, false, true, 32>(a, ft, coef, sm, it, xph, bph);
    else ws_item<PRODUCER, false, false, 32>(a, ft, coef, sm, it, xph, bph);
    nbar_sync(0, NW * 64); // both roles, from their own loops: every box, published plane and noise buffer of this tile has been read; s_ticket[cur ^ 1] is visible
  }
}

__global__ void __launch_bounds__(WS_NW * 64, 1) sweep3d_ws_kernel(const __grid_constant__ Args a)
{
  constexpr int NW = WS_NW;
  using L = Smem<NW, WS_SX, WS_SB, true>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ int s_ticket[2];
  unsigned char *base = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  auto          *fts  = reinterpret_cast<fastnormal::SharedTablesT<L::TREP> *>(base + L::off_tab);
  Coef          *coef = reinterpret_cast<Coef *>(base + L::off_coef);
  const auto     ft   = fastnormal::load_tables(*fts);
  if (threadIdx.x < 8) coef[threadIdx.x] = a.coef[threadIdx.x];
  const uint32_t sm = smem_u32(base);
  if (threadIdx.x == 0) {
    for (int s = 0; s < WS_SX + WS_SB; ++s) mbar_init(sm + (uint32_t)L::off_bar + s * 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    ws_draw(a, s_ticket);
  }
  __syncthreads();
  // the two roles never meet again in the control flow (ptxas bounds the registers of code that both could reach by the
  // smaller of the two setmaxnreg values)
  if (threadIdx.x >= NW * 32) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    ws_tiles<true>(a, ft, coef, sm, s_ticket);
    return;
  }
  asm volatile("setmaxnreg.inc.sync.aligned.u32 88;");
  ws_tiles<false>(a, ft, coef, sm, s_ticket);
}

} // namespace sweep3d
