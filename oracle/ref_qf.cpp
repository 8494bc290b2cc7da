// TEST INFRASTRUCTURE ONLY. Builds oracle/_ref/libpalace_qf_ref.so from the reference's OWN
// QFunction headers where they lie under /root/reference (never copied into this repo):
//   palace/fem/qfunctions/33/{geom,hdiv,hcurl,hdivmass,hdiv_build,hcurl_build,hdivmass_build}_33_qf.h
//   palace/fem/qfunctions/apply/apply_33_qf.h
//   palace/fem/qfunctions/32/{geom,hcurl}_32_qf.h   (boundary integrators)
//   palace/fem/qfunctions/31/{geom,hcurl}_31_qf.h   (line elements in 3-D: boundaries of the wave ports' 2-D submeshes)
//   palace/fem/qfunctions/33/{hcurlhdiv,hcurlhdiv_error}_33_qf.h   (mixed curl / weak curl integrators, flux error estimators)
// behind the minimal libCEED macro shim below. Used only to pin oracle.cpp's restatement of the
// pointwise arithmetic (tests/test_oracle_golden.py, tests/test_bdr_cpu.py). The libCEED operator/basis/restriction layer and
// MFEM are un-vendored, so this is the only part of the reference path that compiles here.
#define CEED_QFUNCTION(name) static inline int name
#define CEED_QFUNCTION_HELPER static inline
#define CeedPragmaSIMD
typedef double CeedScalar;
typedef int CeedInt;

#include "fem/qfunctions/33/geom_33_qf.h"
#include "fem/qfunctions/33/hcurl_33_qf.h"
#include "fem/qfunctions/33/hdiv_33_qf.h"
#include "fem/qfunctions/33/hdivmass_33_qf.h"
#include "fem/qfunctions/33/hcurl_build_33_qf.h"
#include "fem/qfunctions/33/hdiv_build_33_qf.h"
#include "fem/qfunctions/33/hdivmass_build_33_qf.h"
#include "fem/qfunctions/apply/apply_33_qf.h"
// boundary (2-D elements embedded in 3-D) geometry factors and the H(curl) mass QFunction
#include "fem/qfunctions/32/geom_32_qf.h"
#include "fem/qfunctions/32/hcurl_32_qf.h"
// 1-D elements embedded in 3-D
#include "fem/qfunctions/31/geom_31_qf.h"
#include "fem/qfunctions/31/hcurl_31_qf.h"
// scalar curl of a 2-D element (boundary curl-curl: integ/curlcurl.cpp:54-60, case 32 -> f_apply_l2_1)
#include "fem/qfunctions/1/l2_1_qf.h"
// mixed H(curl) / H(div) mass (MixedVectorCurl / MixedVectorWeakCurl integrators, FluxProjector) and the element error integrands
#include "fem/qfunctions/33/hcurlhdiv_33_qf.h"
#include "fem/qfunctions/33/hcurlhdiv_error_33_qf.h"

extern "C"
{
// in0 = {attr[Q], qw[Q], J[9][Q]} split as the reference passes them (geom_33_qf.h:12).
int ref_build_geom_factor_33(int Q, const double *attr, const double *qw, const double *J, double *qdata)
{
  const double *in[3] = {attr, qw, J};
  double *out[1] = {qdata};
  return f_build_geom_factor_33(nullptr, Q, in, out);
}
int ref_apply_hcurl_33(void *ctx, int Q, const double *qdata, const double *u, double *v)
{
  const double *in[2] = {qdata, u};
  double *out[1] = {v};
  return f_apply_hcurl_33(ctx, Q, in, out);
}
int ref_apply_hdiv_33(void *ctx, int Q, const double *qdata, const double *u, double *v)
{
  const double *in[2] = {qdata, u};
  double *out[1] = {v};
  return f_apply_hdiv_33(ctx, Q, in, out);
}
int ref_apply_hdivmass_33(void *ctx, int Q, const double *qdata, const double *u, const double *curlu, double *v,
                          double *curlv)
{
  const double *in[3] = {qdata, u, curlu};
  double *out[2] = {v, curlv};
  return f_apply_hdivmass_33(ctx, Q, in, out);
}
// in = {attr[Q], qw[Q], J[6][Q]} (geom_32_qf.h:12): qdata[8][Q] = {attr, w |J|, (adj(J)^T / |J|)[6]}
int ref_build_geom_factor_32(int Q, const double *attr, const double *qw, const double *J, double *qdata)
{
  const double *in[3] = {attr, qw, J};
  double *out[1] = {qdata};
  return f_build_geom_factor_32(nullptr, Q, in, out);
}
int ref_apply_hcurl_32(void *ctx, int Q, const double *qdata, const double *u, double *v)
{
  const double *in[2] = {qdata, u};
  double *out[1] = {v};
  return f_apply_hcurl_32(ctx, Q, in, out);
}
// in = {attr[Q], qw[Q], J[3][Q]} (geom_31_qf.h:12): qdata[5][Q] = {attr, w |J|, (adj(J)^T / |J|)[3]}
int ref_build_geom_factor_31(int Q, const double *attr, const double *qw, const double *J, double *qdata)
{
  const double *in[3] = {attr, qw, J};
  double *out[1] = {qdata};
  return f_build_geom_factor_31(nullptr, Q, in, out);
}
int ref_apply_hcurl_31(void *ctx, int Q, const double *qdata, const double *u, double *v)
{
  const double *in[2] = {qdata, u};
  double *out[1] = {v};
  return f_apply_hcurl_31(ctx, Q, in, out);
}
int ref_build_hcurl_33(void *ctx, int Q, const double *qdata, double *qd)
{
  const double *in[1] = {qdata};
  double *out[1] = {qd};
  return f_build_hcurl_33(ctx, Q, in, out);
}
int ref_build_hdiv_33(void *ctx, int Q, const double *qdata, double *qd)
{
  const double *in[1] = {qdata};
  double *out[1] = {qd};
  return f_build_hdiv_33(ctx, Q, in, out);
}
int ref_build_hdivmass_33(void *ctx, int Q, const double *qdata, double *qd)
{
  const double *in[1] = {qdata};
  double *out[1] = {qd};
  return f_build_hdivmass_33(ctx, Q, in, out);
}
// v = w detJ (J / detJ)^T C (J^-T) u: H(curl) trial values, H(div) test values (hcurlhdiv_33_qf.h:10-31)
int ref_apply_hcurlhdiv_33(void *ctx, int Q, const double *qdata, const double *u, double *v)
{
  const double *in[2] = {qdata, u};
  double *out[1] = {v};
  return f_apply_hcurlhdiv_33(ctx, Q, in, out);
}
// v = w detJ (J^-T)^T C (J / detJ) u: H(div) trial values, H(curl) test values (hcurlhdiv_33_qf.h:33-55)
int ref_apply_hdivhcurl_33(void *ctx, int Q, const double *qdata, const double *u, double *v)
{
  const double *in[2] = {qdata, u};
  double *out[1] = {v};
  return f_apply_hdivhcurl_33(ctx, Q, in, out);
}
// v[Q] = w detJ |C2 P2 u2 - C1 P1 u1|^2 with a pair context (hcurlhdiv_error_33_qf.h): u1 H(curl) / u2 H(div) and the reverse
int ref_apply_hcurlhdiv_error_33(void *ctx, int Q, const double *qdata, const double *u1, const double *u2, double *v)
{
  const double *in[3] = {qdata, u1, u2};
  double *out[1] = {v};
  return f_apply_hcurlhdiv_error_33(ctx, Q, in, out);
}
int ref_apply_hdivhcurl_error_33(void *ctx, int Q, const double *qdata, const double *u1, const double *u2, double *v)
{
  const double *in[3] = {qdata, u1, u2};
  double *out[1] = {v};
  return f_apply_hdivhcurl_error_33(ctx, Q, in, out);
}
// v[Q] = coeff qw^2 / (w |J|) u: in = {qdata (attr, w |J|, ...), qw, u} (l2_1_qf.h:9-22; EvalMode::Weight is an active input)
int ref_apply_l2_1(void *ctx, int Q, const double *qdata, const double *qw, const double *u, double *v)
{
  const double *in[3] = {qdata, qw, u};
  double *out[1] = {v};
  return f_apply_l2_1(ctx, Q, in, out);
}
}
