// comm.cu -- one process per GPU; NCCL over NVLink for the per-sweep ghost exchange.
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
