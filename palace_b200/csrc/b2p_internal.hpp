// Internal declarations shared by the b2p translation units (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/b2p.h"

struct ncclComm;

namespace b2p
{

void set_error(b2p_ctx *ctx, const char *fmt, ...);

#define B2P_CUDA(ctx, call)                                                                              \
  do                                                                                                     \
  {                                                                                                      \
    cudaError_t e__ = (call);                                                                            \
    if (e__ != cudaSuccess)                                                                              \
    {                                                                                                    \
      b2p::set_error(ctx, "%s:%d CUDA error %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return B2P_ERR_CUDA;                                                                               \
    }                                                                                                    \
  } while (0)

#define B2P_CHECK(ctx, cond, code, ...)     \
  do                                        \
  {                                         \
    if (!(cond))                            \
    {                                       \
      b2p::set_error(ctx, __VA_ARGS__);     \
      return code;                          \
    }                                       \
  } while (0)

template <typename T>
int upload(b2p_ctx *ctx, const T *host, size_t n, T **dptr);

// Kernel launch and dynamic shared memory go through these two macros so that the kernel SOURCES can also be
// compiled for the host-thread SIMT emulation of tests/emu (test infrastructure: B2P_EMU is defined only by
// tests/emu/Makefile, never for libb2p.so). In the product build they expand to the plain CUDA constructs.
#ifdef B2P_EMU
#define B2P_LAUNCH(kern, grid, block, shmem, stream, ...) \
  ::cuda_emu::launch(dim3(grid), dim3(block), (size_t)(shmem), [=]() { kern(__VA_ARGS__); })
#define B2P_LAUNCH_PDL(kern, grid, block, shmem, stream, ...) B2P_LAUNCH(kern, grid, block, shmem, stream, __VA_ARGS__)
#define B2P_DYN_SMEM(type, name) type *name = reinterpret_cast<type *>(::cuda_emu::dyn_smem())
#define B2P_DYN_SMEM_ALIGNED16(type, name) B2P_DYN_SMEM(type, name)
#else
// launch that may overlap the tail of the previous kernel in the stream (the kernel must call griddep_wait() before it
// touches anything the predecessor writes)
#define B2P_LAUNCH_PDL(kern, grid, block, shmem, strm__, ...)                          \
  do                                                                                   \
  {                                                                                    \
    cudaLaunchConfig_t cfg__ = {};                                                     \
    cfg__.gridDim = dim3(grid);                                                        \
    cfg__.blockDim = dim3(block);                                                      \
    cfg__.dynamicSmemBytes = (shmem);                                                  \
    cfg__.stream = (strm__);                                                           \
    cudaLaunchAttribute attr__[1];                                                     \
    attr__[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                 \
    attr__[0].val.programmaticStreamSerializationAllowed = 1;                          \
    cfg__.attrs = attr__;                                                              \
    cfg__.numAttrs = 1;                                                                \
    cudaLaunchKernelEx(&cfg__, kern, __VA_ARGS__);                                     \
  } while (0)
#define B2P_LAUNCH(kern, grid, block, shmem, stream, ...) kern<<<(grid), (block), (shmem), (stream)>>>(__VA_ARGS__)
#define B2P_DYN_SMEM(type, name) extern __shared__ type name[]
#define B2P_DYN_SMEM_ALIGNED16(type, name) extern __shared__ __align__(16) type name[]
#endif

}  // namespace b2p

struct b2p_ctx
{
  int device = 0;
  int sm_count = 0;
  int rank = 0, nranks = 1;
  ncclComm *comm = nullptr;
  cudaStream_t stream = 0;  // stream of the linear algebra / solver layer (b2p_ctx_set_stream)
  cudaStream_t graph_stream = nullptr;  // internal BLOCKING stream for graph capture when `stream` is the legacy stream
  std::string last_error;
  // scratch for reductions (dot products): device partials + pinned host result
  double *d_red = nullptr;
  double *h_red = nullptr;
  size_t red_cap = 0;
  // sink of the branch-free scatter (nd_hex_apply6_kernel): masked restriction entries add 0.0 to a per-thread slot here
  double *d_sink = nullptr;
  static constexpr int SINK_SLOTS = 1 << 16;
};

// Geometry q-data of one element block, device resident.
//   qd[ne][10][Q]: {w detJ, (adjJ^T/detJ)[9] column-major}, points in qslot() order; attr[ne] int32 (1-based).
// (The reference stores attr as an 11th double per point, mesh.cpp:188-195; it is constant per
// element, so it is kept once per element here.)
// Storage slot of quadrature point (qx,qy,qz) inside one component row of q-data: x SLOWEST, so that
// a thread owning a (qy,qz) line reads consecutive words across the warp (conflict-free LDS, see
// b2p_hex_nd2.cu). Logical point order everywhere else is x fastest (basis.cpp:18-19).
__host__ __device__ inline int qslot(int q, int qx, int qy, int qz) { return qy + q * (qz + q * qx); }
__host__ __device__ inline int qslot_of(int q, int iq) { return qslot(q, iq % q, (iq / q) % q, iq / (q * q)); }

// Sentinel in the signed lexicographic restriction: dof masked out (essential BC) -> gathers 0, no scatter.
#define B2P_SKIP_IDX (-2147483647 - 1)

struct b2p_geom
{
  b2p_ctx *ctx = nullptr;
  int ne = 0, q1d = 0, Q = 0;  // q1d == 0: general (non-tensor) point set in plain order
  double *qd = nullptr;
  int32_t *attr = nullptr;
  int refcount = 1;
};

struct b2p_op
{
  b2p_ctx *ctx = nullptr;
  b2p_geom *geom = nullptr;  // shared (refcounted)
  int kind = 0, p = 0, q1d = 0, ne = 0, P = 0;
  int PS = 0;  // restriction row stride: P padded to a multiple of 4 (16-byte rows for TMA bulk copies)
  int64_t lsize = 0;
  int assembled = 0;
  // restriction in LEXICOGRAPHIC element order with the sign folded in:
  // lidx[e][l] >= 0 -> +x[lidx], < 0 -> -x[-1-lidx]   (layout [ne][P])
  int32_t *lidx = nullptr;
  int32_t *lidx_bc = nullptr;  // same with essential dofs replaced by B2P_SKIP_IDX (optional)
  int64_t ess_n = -1;          // fingerprint of the essential set behind lidx_bc: number of distinct dofs, order-free hash
  uint64_t ess_hash = 0;
  std::vector<double> h_tab;   // host copy of the packed 1-D tables (kernel parameter block)
  // 1-D tables (device): Bo[q1d][p], Bc[q1d][p+1], Gc[q1d][p+1]
  double *tab = nullptr;  // packed Bo | Bc | Gc
  // coefficients: material tables + per-element material index for the two parts
  double *mat = nullptr;       // [n_mat_total][9] column-major (mass part first, then curl part)
  int32_t *emat = nullptr;     // [ne][2] indices into mat (value part, derivative part)
  int n_mat = 0;
  bool iso = false;            // every material matrix is c * I
  int tab_sym = -1;            // 1-D tables mirror-symmetric (nd_hex_apply5_kernel): -1 unknown, 0 no, 1 yes
  int tab_sym6 = -1;           // same to round-off (nd_hex_apply6_kernel stores half the rows)
  int tab_sym7 = -1;           // same within 8e-15 of the largest entry (nd_hex_apply7_kernel, p = 4, 5, 6)
  double *ecoef = nullptr;     // [ne][18] per-element coefficient matrices (TMA-friendly copy of mat[emat])
  // assembled q-data (optional): aq[ne][ncomp][Q], symmetric 6 per part
  double *aq = nullptr;
  int aq_ncomp = 0;
  int64_t aq_estride = 0;  // doubles per element (even, so every element block is 16-byte aligned for TMA)
  // dense-basis (non-tensor) operators (b2p_dense.cu)
  bool dense = false;
  double *dense_T = nullptr;   // [Rpad][Ppad] stacked interp/deriv tables, zero padded
  int dense_Ppad = 0, dense_Rpad = 0, dense_row_u = -1, dense_row_c = -1;
  int8_t *curl_orient = nullptr;  // [ne][P][3] tridiagonal orientation (ND tets/prisms p >= 2) or null
  double *pair_zcoef = nullptr;  // [ne][36] {mass, 0, curl, 0}: coefficient block of the two-vector apply (built on first use)
  int coeff_version = 0, pair_version = -1;  // b2p_op_set_coeff bumps coeff_version (coarsened operators follow their parent's)
  bool owns_coeff = true;  // coarsened operators share the fine operator's coefficient arrays
  b2p_op *parent = nullptr;
  int refcount = 1;
};

struct b2p_interp
{
  b2p_ctx *ctx = nullptr;
  int ne = 0, in_P = 0, out_P = 0, in_PS = 0, out_PS = 0, ncomp = 0;
  int64_t in_lsize = 0, out_lsize = 0;
  int32_t *in_lidx = nullptr, *out_lidx = nullptr;  // signed lexicographic restrictions [ne][PS]
  int32_t *out_owner_lidx = nullptr;  // out_lidx with every dof kept in ONE element only (owner-computes Mult), or null
  double *inv_mult = nullptr;                       // [out_lsize] 1 / (local elements touching the dof)
  // per component: dims and matrix offsets into `mats`
  int in_off[3], in_n[3][3], out_off[3], out_n[3][3], mat_off[3][3];
  int ident[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  double *mats = nullptr;
  int n_mats = 0;
  // element-dense variant (any element type, b2p_interp_create_dense): one [out_P][in_P] matrix, native dof order,
  // optional tridiagonal transformations of the two restrictions
  bool dense = false;
  double *dmat = nullptr;
  int8_t *in_co = nullptr, *out_co = nullptr;  // [ne][P][3] rows or null
};

namespace b2p
{
// Element sub-range and owned/ghost split of one apply (defaults: all elements, no ghosts).
struct ApplyRange
{
  int e_off = 0, e_cnt = -1;
  long long n_owned = -1;      // dofs below this index live in x / y, the rest in xg / yg
  const double *xg = nullptr;
  double *yg = nullptr;
  // peer-memory halo: elements >= wait_from_elem need the ghost values; before touching them a warp
  // waits until flags[k] >= expect[k] for k < wait_n (acquire at system scope)
  const unsigned long long *wait_flags = nullptr, *wait_expect = nullptr;
  int wait_n = 0, wait_from_elem = 0;
  bool pdl = false;  // launch with programmatic stream serialisation (kernels that call griddep_wait() only)
};
// Kernel launchers (defined in the .cu files).
int launch_nd_hex_apply(b2p_op *op, const int32_t *lidx, double alpha, const double *x, double *y, const ApplyRange &rg, cudaStream_t s);
int launch_nd_hex_apply4(b2p_op *op, const int32_t *lidx, double alpha, const double *x, double *y, const ApplyRange &rg, cudaStream_t s);
int launch_nd_hex_apply5(b2p_op *op, const int32_t *lidx, double alpha, const double *x, double *y, const ApplyRange &rg, cudaStream_t s);
bool nd_hex_apply5_eligible(b2p_op *op);
int launch_nd_hex_apply6(b2p_op *op, const int32_t *lidx, double alpha, const double *x, double *y, const ApplyRange &rg, cudaStream_t s);
bool nd_hex_apply6_eligible(b2p_op *op);
bool nd_tables_symmetric(const double *tab, int p, int q, double rel_tol);
int launch_nd_hex_apply7(b2p_op *op, const int32_t *lidx, double alpha, const double *x, double *y, const ApplyRange &rg, cudaStream_t s);
bool nd_hex_apply7_eligible(b2p_op *op);
// fused complex apply (b2p_hex_nd4.cu): both parts of a split complex vector in one pass over the geometry
bool nd_hex_apply4z_eligible(const b2p_op *op);
int launch_nd_hex_apply4z(b2p_op *op, int kind, const int32_t *lidx, const double *zcoef, int has_imag, double alpha, const double *xr,
                          const double *xi, double *yr, double *yi, cudaStream_t s);
// Does the op's masked restriction (b2p_op_set_essential) eliminate exactly this set of L dofs?  (No mask <=> empty set.)
bool op_essential_matches(const b2p_op *op, const int32_t *ess_ldofs, int64_t n);
int apply_range(b2p_op *op, const int32_t *lidx, double alpha, const double *x, double *y, const ApplyRange &rg, int flags,
                cudaStream_t s);
int launch_nd_hex_diag(b2p_op *op, double *diag, cudaStream_t s);
int launch_h1_hex_apply(b2p_op *op, const int32_t *lidx, double alpha, const double *x, double *y, const ApplyRange &rg, cudaStream_t s);
int launch_h1_hex_apply3(b2p_op *op, const int32_t *lidx, double alpha, const double *x, double *y, const ApplyRange &rg,
                         cudaStream_t s);
int launch_h1_hex_diag(b2p_op *op, double *diag, cudaStream_t s);
int launch_assemble_qdata(b2p_op *op, cudaStream_t s);
int launch_dense_apply(b2p_op *op, const int32_t *lidx, double alpha, const double *x, double *y, const ApplyRange &rg,
                       cudaStream_t s, bool transpose = false);
int launch_dense_diag(b2p_op *op, double *diag, cudaStream_t s);
int launch_geom_hex(b2p_ctx *ctx, int ne, int k, int q1d, const double *d_xe, const double *d_B, const double *d_G,
                    const double *d_qw, double *d_qd, cudaStream_t s);
}  // namespace b2p
