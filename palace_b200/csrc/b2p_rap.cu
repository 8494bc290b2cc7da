// ParOperator with a GENERAL conforming prolongation P (non-conforming / AMR meshes): the reference applies
//   y = P^T A P x     with the essential true dofs masked before P and overwritten after P^T
// for every finite element space (/root/reference/palace/linalg/rap.cpp:195-234); on conforming meshes P is the
// identity / ghost broadcast that b2p_operator_par builds into its element kernels and halo, on non-conforming meshes it is
// a sparse matrix whose slave rows interpolate from master dofs (mfem::ParFiniteElementSpace::GetProlongationMatrix()).
// Here: the local operator stays what it is (a b2p_operator on the L-vector, no essential dofs), P arrives as a host CSR
// matrix [lsize x tsize] and lives on the device together with its transpose (built once on the host), and
//   Mult              rap.cpp:195-234     tx = x, tx[ess] = 0; lx = P tx; ly = A lx; y = P^T ly; y[ess] = x[ess] | 0
//   MultTranspose     rap.cpp:236-275     the same with A^T
//   AssembleDiagonal  rap.cpp:154-193     d = |P|^T d_L ("a convergent diagonal is assembled with |P|^T d_l": entry-wise absolute
//                                         values, HypreParMatrix::AbsMultTranspose), essential entries 1 | 0
// Single partition. SpMV rows carry 1-30 entries (identity rows for true dofs, a face's worth of masters for a slave), so the
// kernel uses 4 lanes per row; these products are a small fraction of the element kernel's traffic.
#include <algorithm>
#include <memory>
#include <vector>

#include "b2p_linalg.hpp"

struct b2p_operator;
namespace b2p
{
Operator *operator_of(b2p_operator *A);
b2p_operator *wrap_operator(std::unique_ptr<Operator> &&op);
}

struct b2p_spmat
{
  b2p_ctx *ctx = nullptr;
  int64_t rows = 0, cols = 0, nnz = 0;
  int32_t *d_rowptr = nullptr, *d_col = nullptr;    // A   [rows x cols]
  double *d_val = nullptr;
  int32_t *d_rowptr_t = nullptr, *d_col_t = nullptr;  // A^T [cols x rows]
  double *d_val_t = nullptr;
  int refcount = 1;
};

namespace b2p
{
namespace
{
constexpr int SPM_LANES = 4;
// y = op(A) x with op = identity or entry-wise absolute value; lanes of a row fold with shuffles inside their group
template <bool ABS>
__global__ void spmat_mult_kernel(int64_t n, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                                  const double *__restrict__ val, const double *__restrict__ x, double *__restrict__ y)
{
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t r = t / SPM_LANES;
  const int sub = (int)(t % SPM_LANES);
  double s = 0.0;
  if (r < n)
    for (int32_t k = rowptr[r] + sub; k < rowptr[r + 1]; k += SPM_LANES) s += (ABS ? fabs(val[k]) : val[k]) * x[col[k]];
#pragma unroll
  for (int o = SPM_LANES / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (r < n && sub == 0) y[r] = s;
}

int spmat_apply(const b2p_spmat *m, bool transpose, bool absval, const double *x, double *y, cudaStream_t s)
{
  const int64_t n = transpose ? m->cols : m->rows;
  if (n == 0) return B2P_SUCCESS;
  const int32_t *rp = transpose ? m->d_rowptr_t : m->d_rowptr, *cl = transpose ? m->d_col_t : m->d_col;
  const double *vl = transpose ? m->d_val_t : m->d_val;
  const int nt = 256;
  const unsigned grid = (unsigned)((n * SPM_LANES + nt - 1) / nt);
  if (absval)
    B2P_LAUNCH(spmat_mult_kernel<true>, grid, nt, 0, s, n, rp, cl, vl, x, y);
  else
    B2P_LAUNCH(spmat_mult_kernel<false>, grid, nt, 0, s, n, rp, cl, vl, x, y);
  B2P_CUDA(m->ctx, cudaGetLastError());
  return B2P_SUCCESS;
}

void spmat_release(b2p_spmat *m)
{
  if (!m || --m->refcount > 0) return;
  cudaFree(m->d_rowptr);
  cudaFree(m->d_col);
  cudaFree(m->d_val);
  cudaFree(m->d_rowptr_t);
  cudaFree(m->d_col_t);
  cudaFree(m->d_val_t);
  delete m;
}

class RapOperator : public Operator
{
public:
  RapOperator(b2p_ctx *c, Operator *A_, b2p_spmat *P_, const int32_t *ess_tdofs, int64_t n_ess_, int diag_policy_)
    : Operator(c, P_->cols, P_->cols), A(A_), P(P_), n_ess(n_ess_), diag_policy(diag_policy_), lx(c, P_->rows), ly(c, P_->rows)
  {
    P->refcount++;
    if (n_ess > 0)
    {
      ok = upload(c, ess_tdofs, (size_t)n_ess, &d_ess) == B2P_SUCCESS;
      tx.resize(c, P->cols);
    }
    ok = ok && (P->rows == 0 || (lx.p && ly.p)) && (n_ess == 0 || P->cols == 0 || tx.p);
  }
  bool ok = true;  // device allocations of the constructor succeeded
  ~RapOperator() override
  {
    cudaFree(d_ess);
    spmat_release(P);
  }
  void Mult(const double *x, double *y) const override { Apply(x, y, false); }
  void MultTranspose(const double *x, double *y) const override { Apply(x, y, true); }
  const int32_t *EssentialTrueDofs() const override { return d_ess; }
  int64_t NumEssential() const override { return n_ess; }
  void AssembleDiagonal(double *d) const override
  {
    A->AssembleDiagonal(lx.p);
    spmat_apply(P, true, true, lx.p, d, ctx->stream);
    if (n_ess > 0) vec::set_sub(ctx, d, d_ess, n_ess, diag_policy == 1 ? 1.0 : 0.0);
  }

private:
  void Apply(const double *x, double *y, bool transpose) const
  {
    cudaStream_t s = ctx->stream;
    const double *src = x;
    if (n_ess > 0)
    {
      vec::copy(ctx, tx.p, x, width);
      vec::set_sub(ctx, tx.p, d_ess, n_ess, 0.0);
      src = tx.p;
    }
    spmat_apply(P, false, false, src, lx.p, s);
    if (transpose)
      A->MultTranspose(lx.p, ly.p);
    else
      A->Mult(lx.p, ly.p);
    spmat_apply(P, true, false, ly.p, y, s);
    if (n_ess > 0)
    {
      if (diag_policy == 1)
        vec::set_sub_from(ctx, y, d_ess, n_ess, x);
      else
        vec::set_sub(ctx, y, d_ess, n_ess, 0.0);
    }
  }
  Operator *A;  // not owned
  b2p_spmat *P;
  int32_t *d_ess = nullptr;
  int64_t n_ess;
  int diag_policy;
  mutable DVec lx, ly, tx;
};

// y = L A R x with sparse L, R (either may be absent): level prolongations of non-conforming spaces are
// ParOperator(interpolator, trial space, test space, use_R = true) = R_fine I P_coarse (rap.cpp:195-234 with RestrictionMatrixMult,
// rap.cpp:320-345), the transpose P_coarse^T I^T R_fine^T (rap.cpp:236-275).
class TripleOperator : public Operator
{
public:
  TripleOperator(b2p_ctx *c, b2p_spmat *L_, Operator *A_, b2p_spmat *R_)
    : Operator(c, L_ ? L_->rows : A_->height, R_ ? R_->cols : A_->width), L(L_), A(A_), R(R_)
  {
    if (L) L->refcount++;
    if (R) R->refcount++;
    if (R) in.resize(c, A->width);
    if (L) out.resize(c, A->height);
  }
  ~TripleOperator() override
  {
    spmat_release(L);
    spmat_release(R);
  }
  void Mult(const double *x, double *y) const override
  {
    const double *src = x;
    if (R)
    {
      spmat_apply(R, false, false, x, in.p, ctx->stream);
      src = in.p;
    }
    A->Mult(src, L ? out.p : y);
    if (L) spmat_apply(L, false, false, out.p, y, ctx->stream);
  }
  void MultTranspose(const double *x, double *y) const override
  {
    const double *src = x;
    if (L)
    {
      spmat_apply(L, true, false, x, out.p, ctx->stream);
      src = out.p;
    }
    A->MultTranspose(src, R ? in.p : y);
    if (R) spmat_apply(R, true, false, in.p, y, ctx->stream);
  }

private:
  b2p_spmat *L;
  Operator *A;  // not owned
  b2p_spmat *R;
  mutable DVec in, out;
};
}  // namespace
}  // namespace b2p

using namespace b2p;

extern "C"
{

int b2p_spmat_create(b2p_ctx *ctx, int64_t rows, int64_t cols, const int32_t *rowptr, const int32_t *col, const double *val,
                     b2p_spmat **out)
{
  B2P_CHECK(ctx, ctx && rowptr && out && rows >= 0 && cols >= 0, B2P_ERR_ARG, "b2p_spmat_create: bad argument");
  const int64_t nnz = rowptr[rows];
  B2P_CHECK(ctx, rowptr[0] == 0 && nnz >= 0 && (nnz == 0 || (col && val)), B2P_ERR_ARG, "b2p_spmat_create: malformed CSR arrays");
  for (int64_t r = 0; r < rows; r++)
    B2P_CHECK(ctx, rowptr[r + 1] >= rowptr[r], B2P_ERR_ARG, "b2p_spmat_create: rowptr decreases at row %lld", (long long)r);
  for (int64_t k = 0; k < nnz; k++)
    B2P_CHECK(ctx, col[k] >= 0 && col[k] < cols, B2P_ERR_ARG, "b2p_spmat_create: column %d outside [0, %lld)", col[k], (long long)cols);
  // transpose on the host (counting sort by column; rows of the transpose come out sorted by original row)
  std::vector<int32_t> rpt((size_t)cols + 1, 0), clt((size_t)nnz);
  std::vector<double> vlt((size_t)nnz);
  for (int64_t k = 0; k < nnz; k++) rpt[(size_t)col[k] + 1]++;
  for (int64_t c = 0; c < cols; c++) rpt[(size_t)c + 1] += rpt[(size_t)c];
  {
    std::vector<int32_t> fill(rpt.begin(), rpt.end() - 1);
    for (int64_t r = 0; r < rows; r++)
      for (int32_t k = rowptr[r]; k < rowptr[r + 1]; k++)
      {
        const int32_t q = fill[(size_t)col[k]]++;
        clt[(size_t)q] = (int32_t)r;
        vlt[(size_t)q] = val[k];
      }
  }
  auto *m = new b2p_spmat;
  m->ctx = ctx;
  m->rows = rows;
  m->cols = cols;
  m->nnz = nnz;
  int rc;
  if ((rc = upload(ctx, rowptr, (size_t)rows + 1, &m->d_rowptr)) || (rc = upload(ctx, rpt.data(), rpt.size(), &m->d_rowptr_t)) ||
      (nnz > 0 && ((rc = upload(ctx, col, (size_t)nnz, &m->d_col)) || (rc = upload(ctx, val, (size_t)nnz, &m->d_val)) ||
                   (rc = upload(ctx, clt.data(), clt.size(), &m->d_col_t)) || (rc = upload(ctx, vlt.data(), vlt.size(), &m->d_val_t)))))
  {
    spmat_release(m);
    return rc;
  }
  *out = m;
  return B2P_SUCCESS;
}

int b2p_spmat_mult(b2p_spmat *m, int transpose, const double *x, double *y)
{
  if (!m || !x || !y) return B2P_ERR_ARG;
  return spmat_apply(m, transpose != 0, false, x, y, m->ctx->stream);
}

void b2p_spmat_destroy(b2p_spmat *m) { spmat_release(m); }

int b2p_operator_rap(b2p_ctx *ctx, b2p_operator *A_local, b2p_spmat *P, const int32_t *ess_tdofs, int64_t n_ess, int diag_policy,
                     b2p_operator **out)
{
  B2P_CHECK(ctx, ctx && A_local && P && out && n_ess >= 0 && (n_ess == 0 || ess_tdofs), B2P_ERR_ARG, "b2p_operator_rap: bad argument");
  B2P_CHECK(ctx, ctx->nranks == 1, B2P_ERR_UNSUPPORTED, "b2p_operator_rap: partitioned spaces are not supported");
  Operator *A = operator_of(A_local);
  B2P_CHECK(ctx, A->height == P->rows && A->width == P->rows, B2P_ERR_ARG,
            "b2p_operator_rap: the local operator is %lld x %lld, the prolongation has %lld rows", (long long)A->height, (long long)A->width,
            (long long)P->rows);
  for (int64_t i = 0; i < n_ess; i++)
    B2P_CHECK(ctx, ess_tdofs[i] >= 0 && ess_tdofs[i] < P->cols, B2P_ERR_ARG, "b2p_operator_rap: essential true dof %d outside [0, %lld)",
              ess_tdofs[i], (long long)P->cols);
  auto op = std::make_unique<RapOperator>(ctx, A, P, ess_tdofs, n_ess, diag_policy);
  B2P_CHECK(ctx, op->ok, B2P_ERR_CUDA, "b2p_operator_rap: device allocation failed");
  *out = wrap_operator(std::move(op));
  return B2P_SUCCESS;
}

int b2p_operator_triple(b2p_ctx *ctx, b2p_spmat *L, b2p_operator *A_mid, b2p_spmat *R, b2p_operator **out)
{
  B2P_CHECK(ctx, ctx && A_mid && out, B2P_ERR_ARG, "b2p_operator_triple: bad argument");
  Operator *A = operator_of(A_mid);
  B2P_CHECK(ctx, (!L || L->cols == A->height) && (!R || R->rows == A->width), B2P_ERR_ARG,
            "b2p_operator_triple: sizes do not chain (L %lld x %lld, A %lld x %lld, R %lld x %lld)", (long long)(L ? L->rows : 0),
            (long long)(L ? L->cols : 0), (long long)A->height, (long long)A->width, (long long)(R ? R->rows : 0), (long long)(R ? R->cols : 0));
  *out = wrap_operator(std::make_unique<TripleOperator>(ctx, L, A, R));
  return B2P_SUCCESS;
}
}
