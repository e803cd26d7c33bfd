// comm.cu -- one process per GPU.  The per-sweep ghost exchange of the slab-partitioned grid operators goes through PEER
// MEMORY over NVLink / NVSwitch: every rank owns a mailbox that its two neighbours map with CUDA IPC, and one small kernel
// per exchange pushes the boundary units into the neighbours' mailboxes, raises their flags, waits for its own flags and
// copies what arrived into the ghost units (no NCCL kernel, no proxy thread, no rendezvous: ~2 x less latency per exchange
// than a grouped ncclSend / ncclRecv).  NCCL remains for set-up collectives, the gather of replicated levels, the
// personalised exchange of row-partitioned CSR operators, and as the fallback (PMG_NO_P2P, no peer access, huge halos).
//
// Replaces the VecScatter / PetscSF plumbing of the reference (src/mc_sor.c:152-214, :318-319): the host
// program creates one context per rank, rank 0 makes a unique id, the host broadcasts it (MPI_Bcast in a
// PETSc program, torch.distributed in bench.py) and every rank joins the communicator.
#include <dlfcn.h>
#include <nccl.h> // types only: the library is bound at run time so that a process that already holds a
                  // libnccl.so.2 (e.g. PyTorch's bundled one) keeps using that one

#include "common.hpp"

namespace {
struct NcclApi {
  decltype(&ncclGetUniqueId)    GetUniqueId    = nullptr;
  decltype(&ncclCommInitRank)   CommInitRank   = nullptr;
  decltype(&ncclCommDestroy)    CommDestroy    = nullptr;
  decltype(&ncclSend)           Send           = nullptr;
  decltype(&ncclRecv)           Recv           = nullptr;
  decltype(&ncclGroupStart)     GroupStart     = nullptr;
  decltype(&ncclGroupEnd)       GroupEnd       = nullptr;
  decltype(&ncclAllGather)      AllGather      = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  bool                          ok             = false;
};
NcclApi &nccl()
{
  static NcclApi api = [] {
    NcclApi a;
    void   *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return a;
#define PMG_SYM(name) a.name = (decltype(a.name))dlsym(h, "nccl" #name)
    PMG_SYM(GetUniqueId); PMG_SYM(CommInitRank); PMG_SYM(CommDestroy); PMG_SYM(Send); PMG_SYM(Recv);
    PMG_SYM(GroupStart); PMG_SYM(GroupEnd); PMG_SYM(GetErrorString); PMG_SYM(AllGather);
#undef PMG_SYM
    a.ok = a.GetUniqueId && a.CommInitRank && a.Send && a.Recv && a.GroupStart && a.GroupEnd && a.GetErrorString && a.AllGather;
    return a;
  }();
  return api;
}
} // namespace

#define PMG_NCCL_READY()                                                                     \
  do {                                                                                       \
    if (!nccl().ok) PMG_FAIL(PMG_ERR_COMM, "libnccl.so.2 could not be loaded: %s", dlerror()); \
  } while (0)

#define PMG_NCCL(call)                                                                        \
  do {                                                                                        \
    ncclResult_t r_ = (call);                                                                 \
    if (r_ != ncclSuccess) {                                                                  \
      pmg_set_error("%s:%d: NCCL error: %s", __FILE__, __LINE__, nccl().GetErrorString(r_));  \
      return PMG_ERR_COMM;                                                                    \
    }                                                                                         \
  } while (0)

static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");

int comm_allgather_i64(pmg_ctx ctx, const int64_t *local_host, int count, int64_t *all_host);

// ---- peer-memory mailboxes ---------------------------------------------------------------------------------------------
// Mailbox of a rank: flags[channel][from][parity] (64-bit exchange numbers, written by the neighbour), one CTA counter per
// channel (local), then data[channel][from][parity][P2P_SLOT bytes].  channel 0 = exchanges queued on the compute stream,
// 1 = on the communication stream: exchanges of one channel are stream-ordered on every rank, which is what makes two slots
// per neighbour enough -- a rank pushes exchange s only after it has seen the neighbour's push of s-1, which the neighbour
// issued after it had pulled s-2 out of the slot that s is about to overwrite.  from 0 = sent by rank-1, 1 = by rank+1.
namespace p2p {
constexpr size_t SLOT = (size_t)8 << 20, HDR = 4096;
constexpr size_t TOTAL = HDR + 2 * 2 * 2 * SLOT;
__host__ __device__ inline size_t flag_off(int ch, int from, int par) { return (size_t)((ch * 2 + from) * 2 + par) * 8; }
__host__ __device__ inline size_t ctr_off(int ch) { return 256 + (size_t)ch * 8; }
__host__ __device__ inline size_t data_off(int ch, int from, int par) { return HDR + (size_t)((ch * 2 + from) * 2 + par) * SLOT; }

struct Args {
  const double       *send[2];      // my boundary units for rank-1 / rank+1
  double             *recv[2];      // my ghost units below / above
  size_t              n[2];         // doubles per side (0: no neighbour on that side)
  double             *peer_slot[2]; // where my units go in the neighbour's mailbox
  unsigned long long *peer_flag[2];
  const double       *my_slot[2];   // where the neighbour's units arrive in mine
  unsigned long long *my_flag[2];
  unsigned long long  seq[2];
  unsigned long long *counter, target; // CTAs of this channel that have pushed, and the count that completes this exchange
};

// grid-strided copies, four independent 16-byte accesses per thread and trip (the grid is sized so that one or two trips do)
template <bool FROM_MAILBOX> __device__ __forceinline__ void copy_units(double *dst, const double *src, size_t n, size_t t0, size_t nt)
{
  if (((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15) == 0) {
    const size_t   n2 = n >> 1;
    double2       *d  = reinterpret_cast<double2 *>(dst);
    const double2 *s  = reinterpret_cast<const double2 *>(src);
    for (size_t q = t0; q < n2; q += 4 * nt) {
      double2 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (q + u * nt < n2) v[u] = FROM_MAILBOX ? __ldcg(s + q + u * nt) : s[q + u * nt]; // the mailbox is written by another GPU: read it past the L1
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (q + u * nt < n2) d[q + u * nt] = v[u];
    }
    if ((n & 1) && t0 == 0) dst[n - 1] = FROM_MAILBOX ? __ldcg(src + n - 1) : src[n - 1];
  } else {
    for (size_t q = t0; q < n; q += nt) dst[q] = FROM_MAILBOX ? __ldcg(src + q) : src[q];
  }
}

__global__ void __launch_bounds__(256) halo_kernel(const Args a)
{
  const size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (size_t)gridDim.x * blockDim.x;
#pragma unroll
  for (int sd = 0; sd < 2; ++sd)
    if (a.n[sd]) copy_units<false>(a.peer_slot[sd], a.send[sd], a.n[sd], t0, nt);
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(a.counter, 1ull) == a.target - 1) { // every CTA of this exchange has pushed (and fenced): raise the neighbours' flags
      __threadfence_system();
#pragma unroll
      for (int sd = 0; sd < 2; ++sd)
        if (a.n[sd]) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.peer_flag[sd]), "l"(a.seq[sd]) : "memory");
    }
    unsigned long long t_start = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
#pragma unroll
    for (int sd = 0; sd < 2; ++sd) {
      if (!a.n[sd]) continue;
      for (unsigned spin = 0;; ++spin) {
        unsigned long long v;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(a.my_flag[sd]) : "memory");
        if (v >= a.seq[sd]) break;
        if ((spin & 0xffffu) == 0xffffu) { // a neighbour that never arrives must not hang the GPU: fail loudly after 60 s
          unsigned long long t_now;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_now));
          if (t_now - t_start > 60000000000ull) asm volatile("trap;");
        }
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int sd = 0; sd < 2; ++sd)
    if (a.n[sd]) copy_units<true>(a.recv[sd], a.my_slot[sd], a.n[sd], t0, nt);
}
} // namespace p2p

// collective over the communicator: allocate the mailbox, exchange the IPC handles, map the neighbours'
static int comm_p2p_setup(pmg_ctx ctx)
{
  ctx->p2p_ok = false;
  if (ctx->nranks < 2) return 0;
  int64_t mine[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is expected to be 64 bytes");
  if (!std::getenv("PMG_NO_P2P")) {
    void *base = nullptr;
    if (cudaMalloc(&base, p2p::TOTAL) == cudaSuccess && cudaMemset(base, 0, p2p::TOTAL) == cudaSuccess && cudaDeviceSynchronize() == cudaSuccess) {
      cudaIpcMemHandle_t h;
      if (cudaIpcGetMemHandle(&h, base) == cudaSuccess) {
        mine[0] = 1;
        std::memcpy(mine + 1, &h, 64);
        ctx->p2p_base = base;
      } else cudaFree(base);
    }
    cudaGetLastError();
  }
  std::vector<int64_t> all((size_t)9 * ctx->nranks);
  PMG_TRY(comm_allgather_i64(ctx, mine, 9, all.data()));
  bool every = true;
  for (int r = 0; r < ctx->nranks; ++r) every = every && all[(size_t)9 * r] == 1;
  int64_t failed = 0;
  if (every) {
    for (int sd = 0; sd < 2; ++sd) {
      const int nb = ctx->rank + (sd == 0 ? -1 : 1);
      if (nb < 0 || nb >= ctx->nranks) continue;
      cudaIpcMemHandle_t h;
      std::memcpy(&h, &all[(size_t)9 * nb + 1], 64);
      if (cudaIpcOpenMemHandle(&ctx->p2p_peer[sd], h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        ctx->p2p_peer[sd] = nullptr;
        failed            = 1;
      }
    }
  }
  std::vector<int64_t> fails((size_t)ctx->nranks);
  PMG_TRY(comm_allgather_i64(ctx, &failed, 1, fails.data())); // also the barrier after which every mailbox is zeroed and mapped
  for (int r = 0; r < ctx->nranks; ++r) every = every && fails[(size_t)r] == 0;
  ctx->p2p_ok = every;
  if (!every) comm_p2p_teardown(ctx);
  return 0;
}
void comm_p2p_teardown(pmg_ctx ctx)
{
  for (int sd = 0; sd < 2; ++sd) {
    if (ctx->p2p_peer[sd]) cudaIpcCloseMemHandle(ctx->p2p_peer[sd]);
    ctx->p2p_peer[sd] = nullptr;
  }
  if (ctx->p2p_base) cudaFree(ctx->p2p_base);
  ctx->p2p_base = nullptr;
  ctx->p2p_ok   = false;
}

extern "C" {

int pmg_comm_unique_id(unsigned char id[128])
{
  PMG_NCCL_READY();
  ncclUniqueId u;
  PMG_NCCL(nccl().GetUniqueId(&u));
  std::memcpy(id, &u, 128);
  return PMG_OK;
}

int pmg_ctx_comm_init(pmg_ctx ctx, int rank, int nranks, const unsigned char id[128])
{
  if (!ctx || rank < 0 || rank >= nranks) PMG_FAIL(PMG_ERR_ARG, "pmg_ctx_comm_init: bad arguments");
  PMG_NCCL_READY();
  PMG_CUDA(cudaSetDevice(ctx->device));
  ncclUniqueId u;
  std::memcpy(&u, id, 128);
  ncclComm_t comm;
  PMG_NCCL(nccl().CommInitRank(&comm, nranks, u, rank));
  ctx->nccl_comm = comm;
  ctx->rank      = rank;
  ctx->nranks    = nranks;
  if (!ctx->comm_stream) { // highest priority: a halo exchange queued beside a sweep's interior launch must get its CTAs first
    int lo = 0, hi = 0;
    PMG_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    PMG_CUDA(cudaStreamCreateWithPriority(&ctx->comm_stream, cudaStreamNonBlocking, hi));
  }
  PMG_TRY(comm_p2p_setup(ctx));
  return PMG_OK;
}

int pmg_ctx_comm_p2p(pmg_ctx ctx, int *enabled) // 1: the slab halo exchange runs over peer memory; 0: over NCCL
{
  if (!ctx || !enabled) PMG_FAIL(PMG_ERR_ARG, "pmg_ctx_comm_p2p: bad arguments");
  *enabled = ctx->p2p_ok ? 1 : 0;
  return PMG_OK;
}

int pmg_ctx_comm_rank(pmg_ctx ctx, int *rank, int *nranks)
{
  if (rank) *rank = ctx->rank;
  if (nranks) *nranks = ctx->nranks;
  return PMG_OK;
}
}

// exchange `count` doubles with the lower (rank-1) and upper (rank+1) neighbour in one grouped call
int comm_halo_exchange(pmg_ctx ctx, const double *send_lo, double *recv_lo, const double *send_hi, double *recv_hi, size_t count_lo, size_t count_hi, cudaStream_t stream)
{
  if (ctx->nranks == 1) return 0;
  if (ctx->p2p_ok && count_lo * sizeof(double) <= p2p::SLOT && count_hi * sizeof(double) <= p2p::SLOT) { // the same decision on both sides of a pair: their counts agree
    const int   ch = stream == ctx->comm_stream ? 1 : 0;
    p2p::Args   a;
    std::memset(&a, 0, sizeof a);
    const double *send[2] = {send_lo, send_hi};
    double       *recv[2] = {recv_lo, recv_hi};
    const size_t  cnt[2]  = {ctx->rank > 0 ? count_lo : 0, ctx->rank < ctx->nranks - 1 ? count_hi : 0};
    size_t        most = 0;
    for (int sd = 0; sd < 2; ++sd) {
      if (!cnt[sd]) continue;
      const unsigned long long seq = ++ctx->p2p_seq[ch][sd];
      const int                par = (int)(seq & 1);
      char *peer = (char *)ctx->p2p_peer[sd], *mybox = (char *)ctx->p2p_base;
      a.send[sd] = send[sd]; a.recv[sd] = recv[sd]; a.n[sd] = cnt[sd]; a.seq[sd] = seq;
      a.peer_slot[sd] = (double *)(peer + p2p::data_off(ch, 1 - sd, par)); // for rank-1 I am its upper neighbour (from = 1), for rank+1 its lower one
      a.peer_flag[sd] = (unsigned long long *)(peer + p2p::flag_off(ch, 1 - sd, par));
      a.my_slot[sd]   = (const double *)(mybox + p2p::data_off(ch, sd, par));
      a.my_flag[sd]   = (unsigned long long *)(mybox + p2p::flag_off(ch, sd, par));
      most            = std::max(most, cnt[sd]);
    }
    if (!most) return 0;
    // 256 threads x 4 x 16 bytes = 16 KB per CTA and trip; at most 48 CTAs (all of them must be resident: every CTA pushes, then waits)
    static const int cta_kb = std::getenv("PMG_P2P_CTA_KB") ? std::max(1, std::atoi(std::getenv("PMG_P2P_CTA_KB"))) : 16;
    const size_t     per    = (size_t)cta_kb << 10;
    const unsigned   grid   = (unsigned)std::min<size_t>(48, std::max<size_t>(1, (most * sizeof(double) + per - 1) / per));
    a.counter = (unsigned long long *)((char *)ctx->p2p_base + p2p::ctr_off(ch));
    a.target  = (ctx->p2p_ctas[ch] += grid);
    p2p::halo_kernel<<<grid, 256, 0, stream>>>(a);
    PMG_CUDA(cudaGetLastError());
    return 0;
  }
  ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
  if (!comm) PMG_FAIL(PMG_ERR_COMM, "communicator not initialised");
  PMG_NCCL(nccl().GroupStart());
  if (ctx->rank > 0 && count_lo) {
    PMG_NCCL(nccl().Send(send_lo, count_lo, ncclDouble, ctx->rank - 1, comm, stream));
    PMG_NCCL(nccl().Recv(recv_lo, count_lo, ncclDouble, ctx->rank - 1, comm, stream));
  }
  if (ctx->rank < ctx->nranks - 1 && count_hi) {
    PMG_NCCL(nccl().Send(send_hi, count_hi, ncclDouble, ctx->rank + 1, comm, stream));
    PMG_NCCL(nccl().Recv(recv_hi, count_hi, ncclDouble, ctx->rank + 1, comm, stream));
  }
  PMG_NCCL(nccl().GroupEnd());
  return 0;
}

// every rank contributes `count` int64 values; all_host receives nranks*count values in rank order (set-up only: blocking)
int comm_allgather_i64(pmg_ctx ctx, const int64_t *local_host, int count, int64_t *all_host)
{
  if (ctx->nranks == 1) {
    std::memcpy(all_host, local_host, sizeof(int64_t) * (size_t)count);
    return 0;
  }
  ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
  if (!comm) PMG_FAIL(PMG_ERR_COMM, "communicator not initialised");
  DevBuf<int64_t> in, out;
  PMG_TRY(in.upload(local_host, (size_t)count, ctx->stream));
  PMG_TRY(out.alloc((size_t)count * (size_t)ctx->nranks));
  PMG_NCCL(nccl().AllGather(in.p, out.p, (size_t)count, ncclInt64, comm, ctx->stream));
  PMG_CUDA(cudaMemcpyAsync(all_host, out.p, sizeof(int64_t) * (size_t)count * (size_t)ctx->nranks, cudaMemcpyDeviceToHost, ctx->stream));
  PMG_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// variable-size all-gather: rank r contributes counts[r] doubles (its own block is `send`), placed at recv + displs[r]
// on every rank.  One grouped send/recv round; the own block is a device copy.
int comm_allgatherv(pmg_ctx ctx, const double *send, double *recv, const int64_t *counts, const int64_t *displs, cudaStream_t stream)
{
  const int me = ctx->rank;
  if (counts[me] && recv + displs[me] != send) PMG_CUDA(cudaMemcpyAsync(recv + displs[me], send, sizeof(double) * (size_t)counts[me], cudaMemcpyDeviceToDevice, stream));
  if (ctx->nranks == 1) return 0;
  ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
  if (!comm) PMG_FAIL(PMG_ERR_COMM, "communicator not initialised");
  PMG_NCCL(nccl().GroupStart());
  for (int r = 0; r < ctx->nranks; ++r) {
    if (r == me) continue;
    if (counts[me]) PMG_NCCL(nccl().Send(send, (size_t)counts[me], ncclDouble, r, comm, stream));
    if (counts[r]) PMG_NCCL(nccl().Recv(recv + displs[r], (size_t)counts[r], ncclDouble, r, comm, stream));
  }
  PMG_NCCL(nccl().GroupEnd());
  return 0;
}

// personalised exchange: send + send_off[r] .. send_off[r+1] goes to rank r, recv + recv_off[r] .. recv_off[r+1] comes from
// rank r (the per-colour ghost gather of MCSORApply_MPIAIJ, src/mc_sor.c:318-319, for row-partitioned CSR operators)
int comm_exchange_v(pmg_ctx ctx, const double *send, const int64_t *send_off, double *recv, const int64_t *recv_off, cudaStream_t stream)
{
  if (ctx->nranks == 1) return 0;
  ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
  if (!comm) PMG_FAIL(PMG_ERR_COMM, "communicator not initialised");
  PMG_NCCL(nccl().GroupStart());
  for (int r = 0; r < ctx->nranks; ++r) {
    if (r == ctx->rank) continue;
    const int64_t ns = send_off[r + 1] - send_off[r], nr = recv_off[r + 1] - recv_off[r];
    if (ns) PMG_NCCL(nccl().Send(send + send_off[r], (size_t)ns, ncclDouble, r, comm, stream));
    if (nr) PMG_NCCL(nccl().Recv(recv + recv_off[r], (size_t)nr, ncclDouble, r, comm, stream));
  }
  PMG_NCCL(nccl().GroupEnd());
  return 0;
}
