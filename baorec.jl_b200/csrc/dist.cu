// Multi-GPU plumbing: one process per GPU, NCCL communicator created from a unique id that the
// host side (torch.distributed / MPI.jl) broadcasts.  Slabs along z.
#include <nccl.h>
#include <string.h>

#include "internal.cuh"

using namespace baorec;

#define BR_NCCL(expr)                                                                              \
  do {                                                                                             \
    ncclResult_t _e = (expr);                                                                      \
    if (_e != ncclSuccess) {                                                                       \
      baorec::set_error("%s:%d NCCL error %s in %s", __FILE__, __LINE__, ncclGetErrorString(_e), #expr); \
      return BAOREC_ERR_NCCL;                                                                      \
    }                                                                                              \
  } while (0)

extern "C" {

int baorec_comm_unique_id(void* out128) {
  BR_REQUIRE(out128 != nullptr, "out128 is NULL");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
  ncclUniqueId id;
  BR_NCCL(ncclGetUniqueId(&id));
  memcpy(out128, &id, sizeof(id));
  return BAOREC_OK;
}

int baorec_comm_init(baorec_ctx* ctx, int rank, int nranks, const void* unique_id128) {
  BR_REQUIRE(ctx != nullptr && unique_id128 != nullptr, "NULL argument");
  BR_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "rank / nranks");
  BR_CUDA(cudaSetDevice(ctx->device));
  if (ctx->comm) {
    ncclCommDestroy((ncclComm_t)ctx->comm);
    ctx->comm = nullptr;
  }
  ncclUniqueId id;
  memcpy(&id, unique_id128, sizeof(id));
  ncclComm_t comm;
  BR_NCCL(ncclCommInitRank(&comm, nranks, id, rank));
  ctx->comm = comm;
  ctx->rank = rank;
  ctx->nranks = nranks;
  return BAOREC_OK;
}

int baorec_comm_destroy_internal(baorec_ctx* ctx) {
  if (ctx && ctx->comm) {
    ncclCommDestroy((ncclComm_t)ctx->comm);
    ctx->comm = nullptr;
  }
  return BAOREC_OK;
}

int baorec_plan_dist(baorec_ctx* ctx, int nx, int ny, int nz, const float box_size[3], const float box_min[3]) {
  (void)nx; (void)ny; (void)nz; (void)box_size; (void)box_min;
  BR_REQUIRE(ctx != nullptr, "ctx is NULL");
  set_error("baorec_plan_dist: slab-decomposed plan is not available in this build");
  return BAOREC_ERR_INVALID;
}

}  // extern "C"
