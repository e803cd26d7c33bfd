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
    PMG_SYM(GroupStart); PMG_SYM(GroupEnd); PMG_SYM(GetErrorString);
#undef PMG_SYM
    a.ok = a.GetUniqueId && a.CommInitRank && a.Send && a.Recv && a.GroupStart && a.GroupEnd && a.GetErrorString;
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
  if (!ctx->comm_stream) PMG_CUDA(cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
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
