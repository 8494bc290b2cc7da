// NCCL plumbing: one process per GPU, communicator owned by the b2p context. Replaces the
// MPI_Comm the reference threads through its solvers (utils/communication.hpp:188-361).
#include <nccl.h>

#include "b2p_internal.hpp"

using namespace b2p;

#define B2P_NCCL(ctx, call)                                                                             \
  do                                                                                                    \
  {                                                                                                     \
    ncclResult_t r__ = (call);                                                                          \
    if (r__ != ncclSuccess)                                                                             \
    {                                                                                                   \
      b2p::set_error(ctx, "%s:%d NCCL error %s: %s", __FILE__, __LINE__, #call, ncclGetErrorString(r__)); \
      return B2P_ERR_NCCL;                                                                              \
    }                                                                                                   \
  } while (0)

extern "C"
{

int b2p_nccl_unique_id(void *out128)
{
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  if (!out128) return B2P_ERR_ARG;
  ncclUniqueId id;
  B2P_NCCL(nullptr, ncclGetUniqueId(&id));
  memcpy(out128, &id, sizeof(id));
  return B2P_SUCCESS;
}

int b2p_ctx_create_dist(int cuda_device, const void *nccl_unique_id, int rank, int nranks, b2p_ctx **out)
{
  int rc = b2p_ctx_create(cuda_device, out);
  if (rc) return rc;
  b2p_ctx *ctx = *out;
  ctx->rank = rank;
  ctx->nranks = nranks;
  if (nranks > 1)
  {
    B2P_CHECK(ctx, nccl_unique_id, B2P_ERR_ARG, "b2p_ctx_create_dist: unique id required for nranks > 1");
    ncclUniqueId id;
    memcpy(&id, nccl_unique_id, sizeof(id));
    ncclComm_t comm;
    B2P_NCCL(ctx, ncclCommInitRank(&comm, nranks, id, rank));
    ctx->comm = comm;
  }
  return B2P_SUCCESS;
}

}  // extern "C"
