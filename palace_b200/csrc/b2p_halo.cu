// Shared-dof exchange between mesh partitions: the P (owner -> ghost copies) and P^T (ghost
// contributions summed into the owner) of ParOperator (/root/reference/palace/linalg/rap.cpp:212-222),
// which the reference delegates to MFEM's conforming prolongation over MPI. Here every rank keeps its
// L-vector as [owned | ghosts grouped by owner rank], so
//   forward: pack owned values per neighbour -> grouped ncclSend/ncclRecv straight into the ghost segment
//   reverse: ghost segments sent as they lie -> received buffers added into the owners (deterministic order)
// All peers are one NVSwitch hop apart, so one ncclGroup per exchange, on the solver stream.
#include <nccl.h>

#include "b2p_linalg.hpp"
#include "b2p_pipe.cuh"

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

// ---- peer-memory exchange (NVLink stores + flags) ------------------------------------------------
// Mailbox of a rank (one cudaMalloc, IPC-shared with its neighbours):
//   [ fwd_recv : n_ghost doubles ][ rev_recv : n_send doubles ][ flag_fwd : n_nbr u64 ][ flag_rev : n_nbr u64 ]
// A sender packs straight into the peer's mailbox and then raises the peer's flag to its own running
// epoch; the receiver spins on its local flag. Epochs live in device memory, so the sequence can be
// replayed from a CUDA graph. Used only by the paired forward/reverse exchange of ParOperator::Mult.
namespace
{
// grid = (GX, n_nbr): blocks of column k copy segment k into the peer's mailbox; the last block of the
// column to finish (device counter) publishes the data: fence, then release-store of the new epoch into
// the peer's flag.
__global__ void p2p_push_kernel(const double *__restrict__ src, const int32_t *__restrict__ idx, const long long *seg_off,
                                double *const *dst_ptr, unsigned long long *const *flag_ptr, unsigned long long *epoch,
                                unsigned int *done, unsigned long long *bump_expect, const int *recv_has = nullptr,
                                double *zero_buf = nullptr, long long zero_n = 0)
{
  // (optionally) clear the ghost accumulation buffer of this step in the same launch
  for (long long i = (long long)(blockIdx.y * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x; i < zero_n;
       i += (long long)gridDim.x * gridDim.y * blockDim.x)
    zero_buf[i] = 0.0;
  const int k = blockIdx.y;
  const long long b = seg_off[k], e = seg_off[k + 1];
  double *dst = dst_ptr[k];
  for (long long i = b + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < e; i += (long long)gridDim.x * blockDim.x)
    dst[i - b] = idx ? src[idx[i]] : src[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0)
  {
    const unsigned int prev = atomicAdd(done + k, 1u);
    if (prev == gridDim.x - 1)
    {
      done[k] = 0;
      // the consumer kernel (next in stream order) waits for this epoch itself; only neighbours that
      // actually send me ghosts (recv_has[k]) ever raise my flag
      if (bump_expect && recv_has[k]) ++bump_expect[k];
      const unsigned long long ep = ++epoch[k];
      if (e > b)
      {
        __threadfence_system();
        st_release_sys_u64(flag_ptr[k], ep);
      }
    }
  }
}
// Reverse exchange, receiving side: block k waits for neighbour k's flag, then adds its segment into y.
__global__ void p2p_wait_add_kernel(const long long *seg_off, const unsigned long long *flags, unsigned long long *expect,
                                    double *__restrict__ y, const int32_t *__restrict__ idx, const double *__restrict__ buf)
{
  const int k = blockIdx.x;
  const long long b = seg_off[k], e = seg_off[k + 1];
  __shared__ unsigned long long want;
  if (threadIdx.x == 0) want = ++expect[k];
  __syncthreads();
  if (e <= b) return;
  if (threadIdx.x == 0)
  {
    unsigned long long v;
    do
    {
      v = ld_acquire_sys_u64(flags + k);
    } while (v < want);
  }
  __syncthreads();
  for (long long i = b + threadIdx.x; i < e; i += blockDim.x) atomicAdd(y + idx[i], __ldcv(buf + i));
}
__global__ void p2p_wait_kernel(int nseg, const long long *seg_off, const unsigned long long *flags, unsigned long long *expect)
{
  // Same epoch semantics as the in-kernel wait (p2p_push_kernel bumps expect[k] only where recv_has[k]): a neighbour that
  // sends this rank nothing never raises flag k, so its expected epoch must not advance either -- otherwise a later Mult
  // that waits inside the element kernel would spin on it forever.
  const int k = threadIdx.x;
  if (k >= nseg) return;
  if (seg_off[k + 1] > seg_off[k])
  {
    const unsigned long long ep = ++expect[k];
    unsigned long long v;
    do
    {
      v = ld_acquire_sys_u64(flags + k);
    } while (v < ep);
  }
}

// ---- fused variants used by ParOperator::Mult (one launch before and one after the element kernel) ----
// PRE: forward exchange + every zero-fill of the step. Blocks [0, nn*GX) copy column k = block / GX of the send list into
// the peer's mailbox exactly as p2p_push_kernel does (the last block of a column publishes the epoch); then ALL blocks
// zero y (the output T-vector) and the ghost accumulation buffer. Thread k of block 0 also advances the expected epoch of
// this step's reverse exchange, so that every block of the POST kernel can read it without a race.
__global__ void p2p_pre_kernel(const double *__restrict__ src, const int32_t *__restrict__ idx, const long long *seg_off, int nn, int GX,
                               double *const *dst_ptr, unsigned long long *const *flag_ptr, unsigned long long *epoch,
                               unsigned int *done, unsigned long long *bump_expect, const int *recv_has, unsigned long long *expect_rev,
                               double *__restrict__ y, long long ny, double *__restrict__ yg, long long nyg)
{
  griddep_launch_dependents();  // the element kernel may start its prologue now (it waits on this grid before its first scatter / ghost read)
  if (blockIdx.x == 0 && (int)threadIdx.x < nn) ++expect_rev[threadIdx.x];
  if ((int)blockIdx.x < nn * GX)
  {
    const int k = blockIdx.x / GX, bx = blockIdx.x % GX;
    const long long b = seg_off[k], e = seg_off[k + 1];
    double *__restrict__ dst = dst_ptr[k];
    // (GX is sized so that this is a single pass: a short kernel of dependent idx -> x -> remote store chains is bound by
    // their latency, 8 us for 15k values on 16 blocks -- measured with B2P_HALO_TIMING)
    for (long long i = b + (long long)bx * blockDim.x + threadIdx.x; i < e; i += (long long)GX * blockDim.x) dst[i - b] = __ldg(src + __ldg(idx + i));
    // Device-scope fence + block barrier + device-scope counter: the ONE system-scope release by the publishing thread below is
    // cumulative over everything ordered before it (a system-scope fence per thread costs ~10 us per kernel on NVSwitch
    // systems -- measured with B2P_HALO_TIMING: 12.6 us for a PRE kernel that pushes nothing).
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0)
    {
      const unsigned int prev = atomicAdd(done + k, 1u);
      if (prev == (unsigned)GX - 1)
      {
        __threadfence();  // acquire side of the counter: every block's stores are ordered before the release below
        done[k] = 0;
        if (bump_expect && recv_has[k]) ++bump_expect[k];
        const unsigned long long ep = ++epoch[k];
        if (e > b)
        {
          __threadfence_system();
          st_release_sys_u64(flag_ptr[k], ep);
        }
      }
    }
  }
  const long long stride = (long long)gridDim.x * blockDim.x, t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (long long i = t0; i < nyg; i += stride) yg[i] = 0.0;
  if ((reinterpret_cast<uintptr_t>(y) & 15) == 0)
  {
    double2 *y2 = reinterpret_cast<double2 *>(y);
    const long long n2 = ny / 2;
    for (long long i = t0; i < n2; i += stride) y2[i] = make_double2(0.0, 0.0);
    if (t0 == 0 && (ny & 1)) y[ny - 1] = 0.0;
  }
  else
    for (long long i = t0; i < ny; i += stride) y[i] = 0.0;
}
// POST: reverse exchange in one launch. Block (k, bx) copies its slice of ghost segment k into the peer's receive block
// (last block of the column publishes the epoch), then waits for the peer's flag and adds its slice of the values received
// from neighbour k into y. Waiting blocks only depend on the PEER's copies, never on blocks of this grid that have not
// started, so the grid need not be co-resident.
__global__ void p2p_post_kernel(const double *__restrict__ yg, const long long *recv_off, const long long *send_off, int GX,
                                double *const *dst_ptr, unsigned long long *const *flag_ptr, unsigned long long *epoch, unsigned int *done,
                                const unsigned long long *flags, const unsigned long long *expect, double *__restrict__ y,
                                const int32_t *__restrict__ idx, const double *__restrict__ buf)
{
  const int k = blockIdx.x / GX, bx = blockIdx.x % GX;
  griddep_wait();  // launched with programmatic stream serialisation: the element kernel has completed from here on
  {
    const long long b = recv_off[k], e = recv_off[k + 1];
    double *__restrict__ dst = dst_ptr[k];
    for (long long i = b + (long long)bx * blockDim.x + threadIdx.x; i < e; i += (long long)GX * blockDim.x) dst[i - b] = __ldcv(yg + i);
    // Device-scope fence + block barrier + device-scope counter: the ONE system-scope release by the publishing thread below is
    // cumulative over everything ordered before it (a system-scope fence per thread costs ~10 us per kernel on NVSwitch
    // systems -- measured with B2P_HALO_TIMING: 12.6 us for a PRE kernel that pushes nothing).
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0)
    {
      const unsigned int prev = atomicAdd(done + k, 1u);
      if (prev == (unsigned)GX - 1)
      {
        __threadfence();  // acquire side of the counter: every block's stores are ordered before the release below
        done[k] = 0;
        const unsigned long long ep = ++epoch[k];
        if (e > b)
        {
          __threadfence_system();
          st_release_sys_u64(flag_ptr[k], ep);
        }
      }
    }
  }
  const long long b = send_off[k], e = send_off[k + 1];
  if (e <= b) return;
  if (threadIdx.x == 0)
  {
    const unsigned long long want = expect[k];
    unsigned long long v;
    do
    {
      v = ld_acquire_sys_u64(flags + k);
    } while (v < want);
  }
  __syncthreads();
  for (long long i = b + (long long)bx * blockDim.x + threadIdx.x; i < e; i += (long long)GX * blockDim.x) atomicAdd(y + idx[i], __ldcv(buf + i));
}
}  // namespace

// Forward exchange + zero-fill of y and of the ghost accumulators in one launch (ParOperator::Mult). `in_kernel_wait`: the
// element kernel polls the flags itself; otherwise a separate one-warp wait kernel follows.
int halo_pre_p2p(Halo *h, const double *x, double *y, long long ny, bool in_kernel_wait, cudaStream_t s)
{
  const int nn = (int)h->nbr.size();
  const long long ns = nn ? h->send_off.back() : 0;
  long long seg = 0;  // largest segment: one value per thread up to 16k values per neighbour
  for (int k = 0; k < nn; k++) seg = std::max<long long>(seg, h->send_off[k + 1] - h->send_off[k]);
  const int GX = nn ? (int)std::min<long long>((seg + 255) / 256 + 1, 64) : 1;
  int grid = h->ctx->sm_count * 4;
  if (grid < nn * GX) grid = nn * GX;
  B2P_LAUNCH(p2p_pre_kernel, grid, 256, 0, s, x, h->d_send_idx, h->d_send_off, nn, GX, h->d_peer_fwd, h->d_peer_flag_fwd, h->d_epoch, h->d_done,
             in_kernel_wait ? h->d_epoch + 2 * 32 : nullptr, h->d_recv_has, h->d_epoch + 3 * 32, y, ny, h->d_yg, (long long)h->n_ghost);
  if (!in_kernel_wait && nn > 0) B2P_LAUNCH(p2p_wait_kernel, 1, 32, 0, s, nn, h->d_recv_off, h->d_flags, h->d_epoch + 2 * 32);
  B2P_CUDA(h->ctx, cudaGetLastError());
  return B2P_SUCCESS;
}
// Reverse exchange of the step opened by halo_pre_p2p: ghost contributions -> owners, summed into y, in one launch.
int halo_post_p2p(Halo *h, double *y, cudaStream_t s, bool pdl)
{
  const int nn = (int)h->nbr.size();
  if (nn == 0) return B2P_SUCCESS;
  const long long nr = h->recv_off.back(), ns = h->send_off.back();
  long long seg = 0;
  for (int k = 0; k < nn; k++)
    seg = std::max<long long>(seg, std::max<long long>(h->send_off[k + 1] - h->send_off[k], h->recv_off[k + 1] - h->recv_off[k]));
  (void)nr;
  (void)ns;
  const int GX = (int)std::min<long long>((seg + 255) / 256 + 1, 64);
  if (pdl)
    B2P_LAUNCH_PDL(p2p_post_kernel, nn * GX, 256, 0, s, (const double *)h->d_yg, (const long long *)h->d_recv_off, (const long long *)h->d_send_off, GX,
                   (double *const *)h->d_peer_rev, (unsigned long long *const *)h->d_peer_flag_rev, h->d_epoch + 32, h->d_done + 32,
                   (const unsigned long long *)(h->d_flags + 32), (const unsigned long long *)(h->d_epoch + 3 * 32), y,
                   (const int32_t *)h->d_send_idx, (const double *)h->d_mail_rev);
  else
    B2P_LAUNCH(p2p_post_kernel, nn * GX, 256, 0, s, h->d_yg, h->d_recv_off, h->d_send_off, GX, h->d_peer_rev, h->d_peer_flag_rev, h->d_epoch + 32,
               h->d_done + 32, h->d_flags + 32, h->d_epoch + 3 * 32, y, h->d_send_idx, h->d_mail_rev);
  B2P_CUDA(h->ctx, cudaGetLastError());
  return B2P_SUCCESS;
}

int halo_forward_p2p(Halo *h, const double *x, bool in_kernel_wait, cudaStream_t s)
{
  const int nn = (int)h->nbr.size();
  if (nn == 0) return B2P_SUCCESS;
  const long long ns = h->send_off.back();
  dim3 grid((unsigned)std::min<long long>((ns / nn + 255) / 256 + 1, 32), nn);
  B2P_LAUNCH(p2p_push_kernel, grid, 256, 0, s, x, h->d_send_idx, h->d_send_off, h->d_peer_fwd, h->d_peer_flag_fwd, h->d_epoch, h->d_done,
                                       in_kernel_wait ? h->d_epoch + 2 * 32 : nullptr, h->d_recv_has, h->d_yg, h->n_ghost);
  if (!in_kernel_wait) B2P_LAUNCH(p2p_wait_kernel, 1, 32, 0, s, nn, h->d_recv_off, h->d_flags, h->d_epoch + 2 * 32);
  B2P_CUDA(h->ctx, cudaGetLastError());
  return B2P_SUCCESS;
}

int halo_reverse_p2p(Halo *h, double *y, cudaStream_t s)
{
  const int nn = (int)h->nbr.size();
  if (nn == 0) return B2P_SUCCESS;
  const long long nr = h->recv_off.back(), ns = h->send_off.back();
  dim3 grid((unsigned)std::min<long long>((nr / nn + 255) / 256 + 1, 32), nn);
  B2P_LAUNCH(p2p_push_kernel, grid, 256, 0, s, h->d_yg ? h->d_yg : h->d_mail, nullptr, h->d_recv_off, h->d_peer_rev, h->d_peer_flag_rev, h->d_epoch + 32,
                                       h->d_done + 32, nullptr);
  (void)ns;
  B2P_LAUNCH(p2p_wait_add_kernel, nn, 1024, 0, s, h->d_send_off, h->d_flags + 32, h->d_epoch + 3 * 32, y, h->d_send_idx, h->d_mail_rev);
  B2P_CUDA(h->ctx, cudaGetLastError());
  return B2P_SUCCESS;
}

int halo_forward(Halo *h, double *lx)
{
  if (!h || h->nbr.empty()) return B2P_SUCCESS;
  b2p_ctx *c = h->ctx;
  const int64_t ns = h->send_off.back();
  if (ns > 0) B2P_LAUNCH(pack_kernel, (int)std::min<int64_t>((ns + 255) / 256, 1024), 256, 0, c->stream, lx, h->d_send_idx, ns, h->d_buf);
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
  if (ns > 0) B2P_LAUNCH(pack_kernel, (int)std::min<int64_t>((ns + 255) / 256, 1024), 256, 0, s, x, h->d_send_idx, ns, h->d_buf);
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
  if (ns > 0) B2P_LAUNCH(unpack_add_kernel, (int)std::min<int64_t>((ns + 255) / 256, 1024), 256, 0, s, y, h->d_send_idx, ns, h->d_buf);
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
    B2P_LAUNCH(unpack_add_kernel, (int)std::min<int64_t>((ns + 255) / 256, 1024), 256, 0, c->stream, ly, h->d_send_idx, ns, h->d_buf);
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

// Blob exchanged between ranks at set-up (caller all-gathers it): IPC handle + exchange offsets.
struct P2PBlob
{
  cudaIpcMemHandle_t handle;
  int rank, n_nbr;
  long long n_ghost, n_send;
  int nbr[26];
  long long send_off[27], recv_off[27];
};

int b2p_halo_p2p_export(b2p_halo *hh, void *blob, size_t *bytes)
{
  if (!hh || !bytes) return B2P_ERR_ARG;
  *bytes = sizeof(P2PBlob);
  if (!blob) return B2P_SUCCESS;
  Halo &h = hh->h;
  b2p_ctx *ctx = h.ctx;
  B2P_CHECK(ctx, h.nbr.size() <= 26, B2P_ERR_UNSUPPORTED, "b2p_halo_p2p_export: more than 26 neighbours");
  const long long ns = h.send_off.back();
  if (!h.d_mail)
  {
    const size_t words = (size_t)h.n_ghost + (size_t)ns + 64;
    B2P_CUDA(ctx, cudaMalloc((void **)&h.d_mail, words * 8));
    B2P_CUDA(ctx, cudaMemset(h.d_mail, 0, words * 8));
    h.d_mail_rev = h.d_mail + h.n_ghost;
    h.d_flags = (unsigned long long *)(h.d_mail + h.n_ghost + ns);
  }
  P2PBlob b;
  memset(&b, 0, sizeof(b));
  B2P_CUDA(ctx, cudaIpcGetMemHandle(&b.handle, h.d_mail));
  b.rank = ctx->rank;
  b.n_nbr = (int)h.nbr.size();
  b.n_ghost = h.n_ghost;
  b.n_send = ns;
  for (size_t k = 0; k < h.nbr.size(); k++) b.nbr[k] = h.nbr[k];
  for (size_t k = 0; k <= h.nbr.size(); k++)
  {
    b.send_off[k] = h.send_off[k];
    b.recv_off[k] = h.recv_off[k];
  }
  memcpy(blob, &b, sizeof(b));
  return B2P_SUCCESS;
}

int b2p_halo_p2p_import(b2p_halo *hh, const void *blobs, size_t stride, int nranks)
{
  if (!hh || !blobs) return B2P_ERR_ARG;
  Halo &h = hh->h;
  b2p_ctx *ctx = h.ctx;
  B2P_CHECK(ctx, h.d_mail && stride >= sizeof(P2PBlob) && nranks == ctx->nranks, B2P_ERR_ARG, "b2p_halo_p2p_import: export first");
  const int nn = (int)h.nbr.size();
  std::vector<double *> pf(nn), pr(nn);
  std::vector<unsigned long long *> ff(nn), fr(nn);
  for (int k = 0; k < nn; k++)
  {
    P2PBlob pb;
    memcpy(&pb, (const char *)blobs + (size_t)h.nbr[k] * stride, sizeof(pb));
    B2P_CHECK(ctx, pb.rank == h.nbr[k], B2P_ERR_ARG, "b2p_halo_p2p_import: blob %d is from rank %d", h.nbr[k], pb.rank);
    int kk = -1;
    for (int j = 0; j < pb.n_nbr; j++)
      if (pb.nbr[j] == ctx->rank) kk = j;
    B2P_CHECK(ctx, kk >= 0, B2P_ERR_ARG, "b2p_halo_p2p_import: rank %d does not list me as a neighbour", h.nbr[k]);
    B2P_CHECK(ctx, pb.recv_off[kk + 1] - pb.recv_off[kk] == h.send_off[k + 1] - h.send_off[k] &&
                       pb.send_off[kk + 1] - pb.send_off[kk] == h.recv_off[k + 1] - h.recv_off[k],
              B2P_ERR_ARG, "b2p_halo_p2p_import: exchange sizes with rank %d do not match", h.nbr[k]);
    void *base = nullptr;
    B2P_CUDA(ctx, cudaIpcOpenMemHandle(&base, pb.handle, cudaIpcMemLazyEnablePeerAccess));
    h.peer_maps.push_back(base);
    double *mail = (double *)base;
    unsigned long long *flags = (unsigned long long *)(mail + pb.n_ghost + pb.n_send);
    pf[k] = mail + pb.recv_off[kk];               // my forward data -> the peer's ghost segment for me
    pr[k] = mail + pb.n_ghost + pb.send_off[kk];  // my reverse data -> the peer's receive block for me
    ff[k] = flags + kk;
    fr[k] = flags + 32 + kk;
  }
  int rc;
  std::vector<long long> so(h.send_off.begin(), h.send_off.end()), ro(h.recv_off.begin(), h.recv_off.end());
  if ((rc = upload(ctx, (const int64_t *)so.data(), so.size(), (int64_t **)&h.d_send_off))) return rc;
  if ((rc = upload(ctx, (const int64_t *)ro.data(), ro.size(), (int64_t **)&h.d_recv_off))) return rc;
  if ((rc = upload(ctx, (const int64_t *)pf.data(), pf.size(), (int64_t **)&h.d_peer_fwd))) return rc;
  if ((rc = upload(ctx, (const int64_t *)pr.data(), pr.size(), (int64_t **)&h.d_peer_rev))) return rc;
  if ((rc = upload(ctx, (const int64_t *)ff.data(), ff.size(), (int64_t **)&h.d_peer_flag_fwd))) return rc;
  if ((rc = upload(ctx, (const int64_t *)fr.data(), fr.size(), (int64_t **)&h.d_peer_flag_rev))) return rc;
  B2P_CUDA(ctx, cudaMalloc((void **)&h.d_epoch, 128 * 8));
  B2P_CUDA(ctx, cudaMemset(h.d_epoch, 0, 128 * 8));
  {
    std::vector<int32_t> has(32, 0);
    for (int k = 0; k < nn; k++) has[k] = h.recv_off[k + 1] > h.recv_off[k] ? 1 : 0;
    if ((rc = upload(ctx, has.data(), has.size(), (int32_t **)&h.d_recv_has))) return rc;
  }
  B2P_CUDA(ctx, cudaMalloc((void **)&h.d_done, 64 * 4));
  B2P_CUDA(ctx, cudaMemset(h.d_done, 0, 64 * 4));
  // the forward data lands directly in the ghost piece of the input L-vector
  cudaFree(h.d_xg);
  h.d_xg = h.d_mail;
  h.p2p = true;
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
  if (!h->h.p2p) cudaFree(h->h.d_xg);
  cudaFree(h->h.d_yg);
  for (void *m : h->h.peer_maps) cudaIpcCloseMemHandle(m);
  cudaFree(h->h.d_mail);
  cudaFree(h->h.d_epoch);
  cudaFree(h->h.d_done);
  cudaFree(h->h.d_recv_has);
  cudaFree(h->h.d_send_off);
  cudaFree(h->h.d_recv_off);
  cudaFree(h->h.d_peer_fwd);
  cudaFree(h->h.d_peer_rev);
  cudaFree(h->h.d_peer_flag_fwd);
  cudaFree(h->h.d_peer_flag_rev);
  if (h->h.comm_stream) cudaStreamDestroy(h->h.comm_stream);
  if (h->h.ev_in) cudaEventDestroy(h->h.ev_in);
  if (h->h.ev_fwd) cudaEventDestroy(h->h.ev_fwd);
  delete h;
}

}  // extern "C"
