// Shared-dof exchange between mesh partitions: the P (owner -> ghost copies) and P^T (ghost
// contributions summed into the owner) of ParOperator (/root/reference/palace/linalg/rap.cpp:212-222),
// which the reference delegates to MFEM's conforming prolongation over MPI. Here every rank keeps its
// L-vector as [owned | ghosts grouped by owner rank], so
//   forward: pack owned values per neighbour -> grouped ncclSend/ncclRecv straight into the ghost segment
//   reverse: ghost segments sent as they lie -> received buffers added into the owners (deterministic order)
// All peers are one NVSwitch hop apart, so one ncclGroup per exchange, on the solver stream.
#include <nccl.h>

#include "b2p_linalg.hpp"

namespace b2p
{


namespace
{
__global__ void pack_kernel(const double *__restrict__ x, const int32_t *__restrict__ idx, int64_t n, double *__restrict__ buf)
{
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) buf[i] = x[idx[i]];
}
// One thread per received value; values for the same owner dof from different neighbours are
// added with atomics (at most a handful per dof; the per-neighbour buffers are disjoint in idx).
__global__ void unpack_add_kernel(double *__restrict__ y, const int32_t *__restrict__ idx, int64_t n, const double *__restrict__ buf)
{
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) atomicAdd(y + idx[i], buf[i]);
}
}  // namespace

int halo_forward(Halo *h, double *lx)
{
  if (!h || h->nbr.empty()) return B2P_SUCCESS;
  b2p_ctx *c = h->ctx;
  const int64_t ns = h->send_off.back();
  if (ns > 0) pack_kernel<<<(int)std::min<int64_t>((ns + 255) / 256, 1024), 256, 0, c->stream>>>(lx, h->d_send_idx, ns, h->d_buf);
  ncclGroupStart();
  for (size_t k = 0; k < h->nbr.size(); k++)
  {
    const int64_t sc = h->send_off[k + 1] - h->send_off[k], rc = h->recv_off[k + 1] - h->recv_off[k];
    if (sc > 0) ncclSend(h->d_buf + h->send_off[k], sc, ncclDouble, h->nbr[k], (ncclComm_t)c->comm, c->stream);
    if (rc > 0) ncclRecv(lx + h->n_true + h->recv_off[k], rc, ncclDouble, h->nbr[k], (ncclComm_t)c->comm, c->stream);
  }
  ncclResult_t r = ncclGroupEnd();
  B2P_CHECK(c, r == ncclSuccess, B2P_ERR_NCCL, "halo_forward: %s", ncclGetErrorString(r));
  return B2P_SUCCESS;
}

// Split variants: the owned part of the L-vector is the T-vector itself, ghosts live in h->d_xg / d_yg.
int halo_forward_split(Halo *h, const double *x, cudaStream_t s)
{
  if (!h || h->nbr.empty()) return B2P_SUCCESS;
  b2p_ctx *c = h->ctx;
  const int64_t ns = h->send_off.back();
  if (ns > 0) pack_kernel<<<(int)std::min<int64_t>((ns + 255) / 256, 1024), 256, 0, s>>>(x, h->d_send_idx, ns, h->d_buf);
  ncclGroupStart();
  for (size_t k = 0; k < h->nbr.size(); k++)
  {
    const int64_t sc = h->send_off[k + 1] - h->send_off[k], rc = h->recv_off[k + 1] - h->recv_off[k];
    if (sc > 0) ncclSend(h->d_buf + h->send_off[k], sc, ncclDouble, h->nbr[k], (ncclComm_t)c->comm, s);
    if (rc > 0) ncclRecv(h->d_xg + h->recv_off[k], rc, ncclDouble, h->nbr[k], (ncclComm_t)c->comm, s);
  }
  ncclResult_t r = ncclGroupEnd();
  B2P_CHECK(c, r == ncclSuccess, B2P_ERR_NCCL, "halo_forward_split: %s", ncclGetErrorString(r));
  return B2P_SUCCESS;
}

int halo_reverse_split(Halo *h, double *y, cudaStream_t s)
{
  if (!h || h->nbr.empty()) return B2P_SUCCESS;
  b2p_ctx *c = h->ctx;
  const int64_t ns = h->send_off.back();
  ncclGroupStart();
  for (size_t k = 0; k < h->nbr.size(); k++)
  {
    const int64_t sc = h->send_off[k + 1] - h->send_off[k], rc = h->recv_off[k + 1] - h->recv_off[k];
    if (rc > 0) ncclSend(h->d_yg + h->recv_off[k], rc, ncclDouble, h->nbr[k], (ncclComm_t)c->comm, s);
    if (sc > 0) ncclRecv(h->d_buf + h->send_off[k], sc, ncclDouble, h->nbr[k], (ncclComm_t)c->comm, s);
  }
  ncclResult_t r = ncclGroupEnd();
  B2P_CHECK(c, r == ncclSuccess, B2P_ERR_NCCL, "halo_reverse_split: %s", ncclGetErrorString(r));
  if (ns > 0) unpack_add_kernel<<<(int)std::min<int64_t>((ns + 255) / 256, 1024), 256, 0, s>>>(y, h->d_send_idx, ns, h->d_buf);
  return B2P_SUCCESS;
}

int halo_reverse(Halo *h, double *ly)
{
  if (!h || h->nbr.empty()) return B2P_SUCCESS;
  b2p_ctx *c = h->ctx;
  const int64_t ns = h->send_off.back();
  ncclGroupStart();
  for (size_t k = 0; k < h->nbr.size(); k++)
  {
    const int64_t sc = h->send_off[k + 1] - h->send_off[k], rc = h->recv_off[k + 1] - h->recv_off[k];
    if (rc > 0) ncclSend(ly + h->n_true + h->recv_off[k], rc, ncclDouble, h->nbr[k], (ncclComm_t)c->comm, c->stream);
    if (sc > 0) ncclRecv(h->d_buf + h->send_off[k], sc, ncclDouble, h->nbr[k], (ncclComm_t)c->comm, c->stream);
  }
  ncclResult_t r = ncclGroupEnd();
  B2P_CHECK(c, r == ncclSuccess, B2P_ERR_NCCL, "halo_reverse: %s", ncclGetErrorString(r));
  if (ns > 0)
    unpack_add_kernel<<<(int)std::min<int64_t>((ns + 255) / 256, 1024), 256, 0, c->stream>>>(ly, h->d_send_idx, ns, h->d_buf);
  return B2P_SUCCESS;
}

}  // namespace b2p

using namespace b2p;

struct b2p_halo
{
  Halo h;
};
namespace b2p
{
Halo *halo_of(b2p_halo *h) { return h ? &h->h : nullptr; }
}

extern "C"
{

int b2p_halo_create(b2p_ctx *ctx, int64_t n_true, int64_t n_ghost, int n_nbr, const int32_t *nbr_ranks, const int64_t *send_counts,
                    const int32_t *send_idx, const int64_t *recv_counts, b2p_halo **out)
{
  B2P_CHECK(ctx, ctx && out && n_true >= 0 && n_ghost >= 0 && n_nbr >= 0, B2P_ERR_ARG, "b2p_halo_create: bad argument");
  b2p_halo *hh = new b2p_halo;
  Halo &h = hh->h;
  h.ctx = ctx;
  h.n_true = n_true;
  h.n_ghost = n_ghost;
  h.send_off.assign(1, 0);
  h.recv_off.assign(1, 0);
  for (int k = 0; k < n_nbr; k++)
  {
    B2P_CHECK(ctx, nbr_ranks[k] >= 0 && nbr_ranks[k] < ctx->nranks && nbr_ranks[k] != ctx->rank, B2P_ERR_ARG,
              "b2p_halo_create: bad neighbour rank %d", nbr_ranks[k]);
    h.nbr.push_back(nbr_ranks[k]);
    h.send_off.push_back(h.send_off.back() + send_counts[k]);
    h.recv_off.push_back(h.recv_off.back() + recv_counts[k]);
  }
  B2P_CHECK(ctx, h.recv_off.back() == n_ghost, B2P_ERR_ARG, "b2p_halo_create: receive counts (%lld) != ghosts (%lld)",
            (long long)h.recv_off.back(), (long long)n_ghost);
  const int64_t ns = h.send_off.back();
  for (int64_t i = 0; i < ns; i++)
    B2P_CHECK(ctx, send_idx[i] >= 0 && send_idx[i] < n_true, B2P_ERR_ARG, "b2p_halo_create: send index outside the owned range");
  int rc;
  if ((rc = upload(ctx, send_idx, (size_t)ns, &h.d_send_idx))) return rc;
  if (ns > 0) B2P_CUDA(ctx, cudaMalloc((void **)&h.d_buf, sizeof(double) * ns));
  if (n_ghost > 0)
  {
    B2P_CUDA(ctx, cudaMalloc((void **)&h.d_xg, sizeof(double) * n_ghost));
    B2P_CUDA(ctx, cudaMalloc((void **)&h.d_yg, sizeof(double) * n_ghost));
  }
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  B2P_CUDA(ctx, cudaStreamCreateWithPriority(&h.comm_stream, cudaStreamNonBlocking, hi));
  B2P_CUDA(ctx, cudaEventCreateWithFlags(&h.ev_in, cudaEventDisableTiming));
  B2P_CUDA(ctx, cudaEventCreateWithFlags(&h.ev_fwd, cudaEventDisableTiming));
  B2P_CHECK(ctx, n_nbr == 0 || ctx->comm, B2P_ERR_NCCL, "b2p_halo_create: context has no NCCL communicator");
  *out = hh;
  return B2P_SUCCESS;
}

int b2p_halo_forward(b2p_halo *h, double *lvec)
{
  if (!h || !lvec) return B2P_ERR_ARG;
  return halo_forward(&h->h, lvec);
}
int b2p_halo_reverse(b2p_halo *h, double *lvec)
{
  if (!h || !lvec) return B2P_ERR_ARG;
  return halo_reverse(&h->h, lvec);
}
void b2p_halo_destroy(b2p_halo *h)
{
  if (!h) return;
  cudaFree(h->h.d_send_idx);
  cudaFree(h->h.d_buf);
  cudaFree(h->h.d_xg);
  cudaFree(h->h.d_yg);
  if (h->h.comm_stream) cudaStreamDestroy(h->h.comm_stream);
  if (h->h.ev_in) cudaEventDestroy(h->h.ev_in);
  if (h->h.ev_fwd) cudaEventDestroy(h->h.ev_fwd);
  delete h;
}

}  // extern "C"
