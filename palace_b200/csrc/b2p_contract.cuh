// Shared-memory tensor-contraction helpers for the sum-factorised element kernels.
#pragma once

namespace b2p
{

__device__ __forceinline__ double gather1(const double *__restrict__ x, int32_t gi)
{
  if (gi == B2P_SKIP_IDX) return 0.0;
  return (gi >= 0) ? __ldg(x + gi) : -__ldg(x + (-1 - gi));
}
__device__ __forceinline__ void scatter1(double *y, int32_t gi, double v)
{
  if (gi == B2P_SKIP_IDX) return;
  if (gi >= 0)
    atomicAdd(y + gi, v);
  else
    atomicAdd(y + (-1 - gi), -v);
}

// One pencil of a 3-D tensor contraction along axis AX.
//   in  has dims (D0, D1, D2), x fastest; the contracted axis has NIN entries and becomes NOUT.
//   out[o] (+)= sgn * sum_i M[o*mso + i*msi] * in[i]
// L-vector stored in two pieces: owned dofs [0, n_owned) in the true-dof vector itself, ghost dofs
// in a separate small buffer (so a ParOperator never copies T-vectors into L-vectors).
struct VSplit
{
  long long n_owned;
  const double *xg;
  double *yg;
};
__device__ __forceinline__ const double *split_src(const double *x, const VSplit &sp, int32_t a)
{
  return a < sp.n_owned ? x + a : sp.xg + (a - sp.n_owned);
}
__device__ __forceinline__ double gather2(const double *__restrict__ x, const VSplit &sp, int32_t gi)
{
  if (gi == B2P_SKIP_IDX) return 0.0;
  return (gi >= 0) ? __ldg(split_src(x, sp, gi)) : -__ldg(split_src(x, sp, -1 - gi));
}
__device__ __forceinline__ void scatter2(double *y, const VSplit &sp, int32_t gi, double v)
{
  if (gi == B2P_SKIP_IDX) return;
  const int32_t a = gi >= 0 ? gi : -1 - gi;
  atomicAdd(a < sp.n_owned ? y + a : sp.yg + (a - sp.n_owned), gi >= 0 ? v : -v);
}

template <int AX, int D0, int D1, int D2, int NIN, int NOUT, bool ACC>
__device__ __forceinline__ void pencil(int r, const double *__restrict__ in, double *__restrict__ out,
                                       const double *__restrict__ M, int mso, int msi, double sgn)
{
  static_assert((AX == 0 ? D0 : AX == 1 ? D1 : D2) == NIN, "axis size mismatch");
  int ibase, obase, istr, ostr;
  if (AX == 0)
  {
    ibase = r * D0;  // r = j + k*D1
    obase = r * NOUT;
    istr = ostr = 1;
  }
  else if (AX == 1)
  {
    const int i = r % D0, k = r / D0;
    ibase = i + k * D0 * D1;
    obase = i + k * D0 * NOUT;
    istr = ostr = D0;
  }
  else
  {
    ibase = obase = r;  // r = i + j*D0
    istr = ostr = D0 * D1;
  }
  double v[NIN];
#pragma unroll
  for (int i = 0; i < NIN; i++) v[i] = in[ibase + i * istr];
#pragma unroll
  for (int o = 0; o < NOUT; o++)
  {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NIN; i++) s += M[o * mso + i * msi] * v[i];
    if (ACC)
      out[obase + o * ostr] += sgn * s;
    else
      out[obase + o * ostr] = sgn * s;
  }
}

template <int AX, int D0, int D1, int D2>
struct NPencil
{
  static constexpr int value = (AX == 0 ? D1 * D2 : AX == 1 ? D0 * D2 : D0 * D1);
};

// Block-wide contraction over NEB elements; `in`/`out` are per-element arrays with strides.
template <int AX, int D0, int D1, int D2, int NIN, int NOUT, bool ACC, int NEB, int NT>
__device__ __forceinline__ void contract(const double *in, int in_estride, double *out, int out_estride,
                                         const double *M, int mso, int msi, double sgn)
{
  constexpr int NP = NPencil<AX, D0, D1, D2>::value;
  for (int w = threadIdx.x; w < NEB * NP; w += NT)
  {
    const int e = w / NP, r = w % NP;
    pencil<AX, D0, D1, D2, NIN, NOUT, ACC>(r, in + e * in_estride, out + e * out_estride, M, mso, msi, sgn);
  }
}

}  // namespace b2p
