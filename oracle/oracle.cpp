// TEST INFRASTRUCTURE ONLY (CPU oracle). Nothing under oracle/ is shipped or timed as product;
// only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg call it.
//
// CPU restatement of the reference's partial-assembly operator path for H(curl) / H1 hexahedra:
//   y_L (+)= sum_e  E_e^T [ interp^T v + deriv^T w ],  (v, w) = D_q( interp E_e x,  deriv E_e x )
// exactly as Palace hands it to libCEED: a NON-tensor dense basis for vector elements
// (/root/reference/palace/fem/libceed/basis.cpp:40-85,169-186), native-order element restriction
// with sign flips (/root/reference/palace/fem/libceed/restriction.cpp:207-297), stored geometry
// q-data {attr, w detJ, adj(J)^T/detJ} (/root/reference/palace/fem/mesh.cpp:146-209,
// /root/reference/palace/fem/qfunctions/33/geom_33_qf.h:9-34) and the pointwise QFunctions
// (/root/reference/palace/fem/qfunctions/33/{hdiv,hcurl,hdivmass}_33_qf.h).
//
// PARITY PINNING: the pointwise D and geometry-factor arithmetic is pinned against the
// reference's own QFunction headers compiled from /root/reference (oracle/_ref, see ref_qf.cpp and
// tests/test_oracle_golden.py, tests/test_bdr_cpu.py). The element tables (MFEM ND_HexahedronElement, un-vendored, pinned
// tag d9d6526c) are a restatement from the published element definition: "parity unpinned" at
// the MFEM boundary; pinned only through operator identities and analytic cavity eigenvalues.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#include <pthread.h>
#include <sched.h>
#include "basis1d.hpp"

using orc::ld;

namespace
{

// ---- MFEM ND_HexahedronElement native ordering (restated; MFEM fem/fe/fe_nd.cpp at d9d6526c) ----
// dof_map[lex] = native index, or -1-native when the native basis function is the negated
// lexicographic one. Lexicographic layout: x-directed block (i<p open, j,k<=p closed, i fastest),
// then y-directed (i<=p, j<p, k<=p), then z-directed (i,j<=p, k<p).
void nd_hex_dofmap(int p, std::vector<int> &dof_map)
{
  const int dof3 = p * (p + 1) * (p + 1);
  dof_map.assign(3 * dof3, 0);
  int o = 0;
  // edges: (0,1) (1,2) (3,2) (0,3) (4,5) (5,6) (7,6) (4,7) (0,4) (1,5) (2,6) (3,7)
  for (int i = 0; i < p; i++) dof_map[0 * dof3 + i + (0 + 0 * (p + 1)) * p] = o++;
  for (int i = 0; i < p; i++) dof_map[1 * dof3 + p + (i + 0 * p) * (p + 1)] = o++;
  for (int i = 0; i < p; i++) dof_map[0 * dof3 + i + (p + 0 * (p + 1)) * p] = o++;
  for (int i = 0; i < p; i++) dof_map[1 * dof3 + 0 + (i + 0 * p) * (p + 1)] = o++;
  for (int i = 0; i < p; i++) dof_map[0 * dof3 + i + (0 + p * (p + 1)) * p] = o++;
  for (int i = 0; i < p; i++) dof_map[1 * dof3 + p + (i + p * p) * (p + 1)] = o++;
  for (int i = 0; i < p; i++) dof_map[0 * dof3 + i + (p + p * (p + 1)) * p] = o++;
  for (int i = 0; i < p; i++) dof_map[1 * dof3 + 0 + (i + p * p) * (p + 1)] = o++;
  for (int i = 0; i < p; i++) dof_map[2 * dof3 + 0 + (0 + i * (p + 1)) * (p + 1)] = o++;
  for (int i = 0; i < p; i++) dof_map[2 * dof3 + p + (0 + i * (p + 1)) * (p + 1)] = o++;
  for (int i = 0; i < p; i++) dof_map[2 * dof3 + p + (p + i * (p + 1)) * (p + 1)] = o++;
  for (int i = 0; i < p; i++) dof_map[2 * dof3 + 0 + (p + i * (p + 1)) * (p + 1)] = o++;
  // faces
  // (3,2,1,0) bottom
  for (int j = 1; j < p; j++)
    for (int i = 0; i < p; i++) dof_map[0 * dof3 + i + ((p - j) + 0 * (p + 1)) * p] = o++;
  for (int j = 0; j < p; j++)
    for (int i = 1; i < p; i++) dof_map[1 * dof3 + i + ((p - 1 - j) + 0 * p) * (p + 1)] = -1 - (o++);
  // (0,1,5,4) front
  for (int k = 1; k < p; k++)
    for (int i = 0; i < p; i++) dof_map[0 * dof3 + i + (0 + k * (p + 1)) * p] = o++;
  for (int k = 0; k < p; k++)
    for (int i = 1; i < p; i++) dof_map[2 * dof3 + i + (0 + k * (p + 1)) * (p + 1)] = o++;
  // (1,2,6,5) right
  for (int k = 1; k < p; k++)
    for (int j = 0; j < p; j++) dof_map[1 * dof3 + p + (j + k * p) * (p + 1)] = o++;
  for (int k = 0; k < p; k++)
    for (int j = 1; j < p; j++) dof_map[2 * dof3 + p + (j + k * (p + 1)) * (p + 1)] = o++;
  // (2,3,7,6) back
  for (int k = 1; k < p; k++)
    for (int i = 0; i < p; i++) dof_map[0 * dof3 + (p - 1 - i) + (p + k * (p + 1)) * p] = -1 - (o++);
  for (int k = 0; k < p; k++)
    for (int i = 1; i < p; i++) dof_map[2 * dof3 + (p - i) + (p + k * (p + 1)) * (p + 1)] = o++;
  // (3,0,4,7) left
  for (int k = 1; k < p; k++)
    for (int j = 0; j < p; j++) dof_map[1 * dof3 + 0 + ((p - 1 - j) + k * p) * (p + 1)] = -1 - (o++);
  for (int k = 0; k < p; k++)
    for (int j = 1; j < p; j++) dof_map[2 * dof3 + 0 + ((p - j) + k * (p + 1)) * (p + 1)] = o++;
  // (4,5,6,7) top
  for (int j = 1; j < p; j++)
    for (int i = 0; i < p; i++) dof_map[0 * dof3 + i + (j + p * (p + 1)) * p] = o++;
  for (int j = 0; j < p; j++)
    for (int i = 1; i < p; i++) dof_map[1 * dof3 + i + (j + p * p) * (p + 1)] = o++;
  // interior
  for (int k = 1; k < p; k++)
    for (int j = 1; j < p; j++)
      for (int i = 0; i < p; i++) dof_map[0 * dof3 + i + (j + k * (p + 1)) * p] = o++;
  for (int k = 1; k < p; k++)
    for (int j = 0; j < p; j++)
      for (int i = 1; i < p; i++) dof_map[1 * dof3 + i + (j + k * p) * (p + 1)] = o++;
  for (int k = 0; k < p; k++)
    for (int j = 1; j < p; j++)
      for (int i = 1; i < p; i++) dof_map[2 * dof3 + i + (j + k * (p + 1)) * (p + 1)] = o++;
}

struct Hex1D
{
  orc::Table1D open, closed;  // at quadrature points
  std::vector<ld> qx, qw;
};

Hex1D make_1d(int p, int q1d)
{
  Hex1D h;
  std::vector<ld> cp, op, ow;
  orc::gauss_lobatto(p + 1, cp);
  orc::gauss_legendre(p, op, ow);
  orc::gauss_legendre(q1d, h.qx, h.qw);
  h.open = orc::make_table(op, h.qx);
  h.closed = orc::make_table(cp, h.qx);
  return h;
}

// ---- pointwise 3x3 helpers: restatement of qfunctions/33/utils_33_qf.h:20-84 (column-major) ----
inline void adjJt33(const double J[9], double A[9])
{
  A[0] = J[4] * J[8] - J[7] * J[5];
  A[3] = J[7] * J[2] - J[1] * J[8];
  A[6] = J[1] * J[5] - J[4] * J[2];
  A[1] = J[6] * J[5] - J[3] * J[8];
  A[4] = J[0] * J[8] - J[6] * J[2];
  A[7] = J[3] * J[2] - J[0] * J[5];
  A[2] = J[3] * J[7] - J[6] * J[4];
  A[5] = J[6] * J[1] - J[0] * J[7];
  A[8] = J[0] * J[4] - J[3] * J[1];
}
// y = A^T B C x
inline void multAtBCx(const double A[9], const double B[9], const double C[9], const double x[3], double y[3])
{
  double t[3], z[3];
  for (int r = 0; r < 3; r++) t[r] = C[r] * x[0] + C[r + 3] * x[1] + C[r + 6] * x[2];
  for (int r = 0; r < 3; r++) z[r] = B[r] * t[0] + B[r + 3] * t[1] + B[r + 6] * t[2];
  for (int r = 0; r < 3; r++) y[r] = A[3 * r] * z[0] + A[3 * r + 1] * z[1] + A[3 * r + 2] * z[2];
}

// Coefficient context: restatement of qfunctions/coeff/coeff_qf.h:7-43 + coeff_3_qf.h:9-23.
union IntScalar
{
  int first;
  double second;
};
static_assert(sizeof(IntScalar) == 8, "ctx entries are 8 bytes");
inline int num_attr(const IntScalar *c) { return c[0].first; }
inline int num_mat(const IntScalar *c) { return c[1 + num_attr(c)].first; }
inline const IntScalar *mat_coeff(const IntScalar *c) { return c + 2 + num_attr(c); }
inline const IntScalar *pair_second(const IntScalar *c) { return c + 2 + num_attr(c) + 9 * num_mat(c); }
inline void coeff_unpack3(const IntScalar *c, int attr, double coeff[9])
{
  const int k = (num_attr(c) > 0) ? c[1 + (attr - 1)].first : 0;
  for (int i = 0; i < 9; i++) coeff[i] = mat_coeff(c)[9 * k + i].second;
}

enum Kind
{
  CURLCURL = 0,       // hdiv_33 on curl           (integ/curlcurl.cpp:47-52)
  ND_MASS = 1,        // hcurl_33 on interp        (integ/vecfemass.cpp:72-105)
  CURLCURL_MASS = 2,  // hdivmass_33 (pair ctx)    (integ/curlcurlmass.cpp:37-44)
  H1_DIFFUSION = 3,   // hcurl_33 on grad          (integ/diffusion.cpp:37-42)
  ND_WEAKCURL = 4,    // hcurlhdiv_33: interp -> curl test values   (integ/mixedveccurl.cpp:68-117, MixedVectorWeakCurlIntegrator)
  ND_MIXEDCURL = 5    // hdivhcurl_33: curl -> interp test values   (integ/mixedveccurl.cpp:23-66, MixedVectorCurlIntegrator)
};
// which reference-space fields a kind evaluates (trial) and tests with: value part u / v, derivative part c / w
inline bool needs_u(int kind) { return kind == ND_MASS || kind == CURLCURL_MASS || kind == ND_WEAKCURL || kind == ND_MIXEDCURL; }
inline bool needs_c(int kind) { return kind != ND_MASS; }

// One quadrature point of D. qd = {attr, wdetJ, adjJt[9]}; u = value part, c = derivative part.
inline void apply_D(int kind, const IntScalar *ctx, const double qd[11], const double u[3], const double c[3],
                    double v[3], double w[3])
{
  const int attr = (int)qd[0];
  const double wdetJ = qd[1];
  const double *adjJt = qd + 2;
  double coeff[9];
  v[0] = v[1] = v[2] = w[0] = w[1] = w[2] = 0.0;
  if (kind == ND_MASS || kind == CURLCURL_MASS)
  {
    coeff_unpack3(ctx, attr, coeff);
    multAtBCx(adjJt, coeff, adjJt, u, v);  // hcurl_33_qf.h:24 / hdivmass_33_qf.h:24
    for (int d = 0; d < 3; d++) v[d] *= wdetJ;
  }
  if (kind == H1_DIFFUSION)
  {
    coeff_unpack3(ctx, attr, coeff);
    multAtBCx(adjJt, coeff, adjJt, c, w);
    for (int d = 0; d < 3; d++) w[d] *= wdetJ;
  }
  if (kind == CURLCURL || kind == CURLCURL_MASS)
  {
    const IntScalar *c2 = (kind == CURLCURL_MASS) ? pair_second(ctx) : ctx;
    double J[9];
    coeff_unpack3(c2, attr, coeff);
    adjJt33(adjJt, J);                     // J/detJ = adj(adjJt/detJ)^T   (hdiv_33_qf.h:24)
    multAtBCx(J, coeff, J, c, w);
    for (int d = 0; d < 3; d++) w[d] *= wdetJ;
  }
  if (kind == ND_WEAKCURL || kind == ND_MIXEDCURL)
  {
    // the two spaces meet at the point: an H(curl) value maps with adjJt / detJ = J^-T, a curl (H(div)) with J / detJ
    double J[9];
    coeff_unpack3(ctx, attr, coeff);
    adjJt33(adjJt, J);
    if (kind == ND_WEAKCURL)
      multAtBCx(J, coeff, adjJt, u, w);  // hcurlhdiv_33_qf.h:24 (f_apply_hcurlhdiv_33): trial value -> test curl
    else
      multAtBCx(adjJt, coeff, J, c, v);  // hcurlhdiv_33_qf.h:48 (f_apply_hdivhcurl_33): trial curl -> test value
    for (int d = 0; d < 3; d++)
    {
      v[d] *= wdetJ;
      w[d] *= wdetJ;
    }
  }
}

}  // namespace

extern "C"
{

int orc_nd_hex_ndof(int p) { return 3 * p * (p + 1) * (p + 1); }

void orc_gauss_legendre(int n, double *x, double *w)
{
  std::vector<ld> xx, ww;
  orc::gauss_legendre(n, xx, ww);
  for (int i = 0; i < n; i++)
  {
    x[i] = (double)xx[i];
    w[i] = (double)ww[i];
  }
}

void orc_gauss_lobatto(int n, double *x)
{
  std::vector<ld> xx;
  orc::gauss_lobatto(n, xx);
  for (int i = 0; i < n; i++) x[i] = (double)xx[i];
}

void orc_nd_hex_dofmap(int p, int *dof_map)
{
  std::vector<int> m;
  nd_hex_dofmap(p, m);
  std::memcpy(dof_map, m.data(), m.size() * sizeof(int));
}

// 1-D tables at the q1d Gauss-Legendre points: Bo[q1d][p], Bc[q1d][p+1], Gc[q1d][p+1], qw[q1d].
void orc_nd_hex_1d(int p, int q1d, double *Bo, double *Bc, double *Gc, double *qw)
{
  Hex1D h = make_1d(p, q1d);
  for (int q = 0; q < q1d; q++)
  {
    for (int i = 0; i < p; i++) Bo[q * p + i] = (double)h.open.b(q, i);
    for (int j = 0; j <= p; j++)
    {
      Bc[q * (p + 1) + j] = (double)h.closed.b(q, j);
      Gc[q * (p + 1) + j] = (double)h.closed.g(q, j);
    }
    qw[q] = (double)h.qw[q];
  }
}

// Dense tables as GetDofToQuad(ir, FULL) supplies them to CeedBasisCreateHcurl
// (basis.cpp:43,71-76): interp[d][q][i], curl[d][q][i] in NATIVE dof order, reference-element
// values; quadrature points x-fastest (basis.cpp:18-19); qw[Q] unnormalised (basis.cpp:51-62).
void orc_nd_hex_tables(int p, int q1d, double *interp, double *curl, double *qw)
{
  Hex1D h = make_1d(p, q1d);
  std::vector<int> dof_map;
  nd_hex_dofmap(p, dof_map);
  const int P = 3 * p * (p + 1) * (p + 1), Q = q1d * q1d * q1d, dof3 = P / 3;
  std::memset(interp, 0, sizeof(double) * 3 * Q * P);
  std::memset(curl, 0, sizeof(double) * 3 * Q * P);
  auto put = [&](double *T, int d, int q, int lex, ld val)
  {
    int n = dof_map[lex];
    ld s = 1;
    if (n < 0)
    {
      n = -1 - n;
      s = -1;
    }
    T[((size_t)d * Q + q) * P + n] = (double)(s * val);
  };
  for (int qz = 0; qz < q1d; qz++)
    for (int qy = 0; qy < q1d; qy++)
      for (int qx = 0; qx < q1d; qx++)
      {
        const int q = qx + q1d * (qy + q1d * qz);
        qw[q] = (double)(h.qw[qx] * h.qw[qy] * h.qw[qz]);
        // x-directed: (o_i(x) c_j(y) c_k(z), 0, 0); curl = (0, d/dz, -d/dy)
        for (int k = 0; k <= p; k++)
          for (int j = 0; j <= p; j++)
            for (int i = 0; i < p; i++)
            {
              const int lex = 0 * dof3 + i + (j + k * (p + 1)) * p;
              put(interp, 0, q, lex, h.open.b(qx, i) * h.closed.b(qy, j) * h.closed.b(qz, k));
              put(curl, 1, q, lex, h.open.b(qx, i) * h.closed.b(qy, j) * h.closed.g(qz, k));
              put(curl, 2, q, lex, -h.open.b(qx, i) * h.closed.g(qy, j) * h.closed.b(qz, k));
            }
        // y-directed: (0, c_i o_j c_k, 0); curl = (-d/dz, 0, d/dx)
        for (int k = 0; k <= p; k++)
          for (int j = 0; j < p; j++)
            for (int i = 0; i <= p; i++)
            {
              const int lex = 1 * dof3 + i + (j + k * p) * (p + 1);
              put(interp, 1, q, lex, h.closed.b(qx, i) * h.open.b(qy, j) * h.closed.b(qz, k));
              put(curl, 0, q, lex, -h.closed.b(qx, i) * h.open.b(qy, j) * h.closed.g(qz, k));
              put(curl, 2, q, lex, h.closed.g(qx, i) * h.open.b(qy, j) * h.closed.b(qz, k));
            }
        // z-directed: (0, 0, c_i c_j o_k); curl = (d/dy, -d/dx, 0)
        for (int k = 0; k < p; k++)
          for (int j = 0; j <= p; j++)
            for (int i = 0; i <= p; i++)
            {
              const int lex = 2 * dof3 + i + (j + k * (p + 1)) * (p + 1);
              put(interp, 2, q, lex, h.closed.b(qx, i) * h.closed.b(qy, j) * h.open.b(qz, k));
              put(curl, 0, q, lex, h.closed.b(qx, i) * h.closed.g(qy, j) * h.open.b(qz, k));
              put(curl, 1, q, lex, -h.closed.g(qx, i) * h.closed.b(qy, j) * h.open.b(qz, k));
            }
      }
}

// H1 hex (GaussLobatto nodes), LEXICOGRAPHIC dof order (Palace uses the lexicographic
// restriction for scalar tensor elements, restriction.cpp:413-426). interp[Q][P], grad[d][Q][P].
void orc_h1_hex_tables(int p, int q1d, double *interp, double *grad, double *qw)
{
  std::vector<ld> cp, qx, qwv;
  orc::gauss_lobatto(p + 1, cp);
  orc::gauss_legendre(q1d, qx, qwv);
  orc::Table1D t = orc::make_table(cp, qx);
  const int n = p + 1, P = n * n * n, Q = q1d * q1d * q1d;
  for (int qz = 0; qz < q1d; qz++)
    for (int qy = 0; qy < q1d; qy++)
      for (int qxx = 0; qxx < q1d; qxx++)
      {
        const int q = qxx + q1d * (qy + q1d * qz);
        qw[q] = (double)(qwv[qxx] * qwv[qy] * qwv[qz]);
        for (int k = 0; k < n; k++)
          for (int j = 0; j < n; j++)
            for (int i = 0; i < n; i++)
            {
              const int l = i + n * (j + n * k);
              interp[(size_t)q * P + l] = (double)(t.b(qxx, i) * t.b(qy, j) * t.b(qz, k));
              grad[((size_t)0 * Q + q) * P + l] = (double)(t.g(qxx, i) * t.b(qy, j) * t.b(qz, k));
              grad[((size_t)1 * Q + q) * P + l] = (double)(t.b(qxx, i) * t.g(qy, j) * t.b(qz, k));
              grad[((size_t)2 * Q + q) * P + l] = (double)(t.b(qxx, i) * t.b(qy, j) * t.g(qz, k));
            }
      }
}

// Geometry q-data for order-k hexes. xe[ne][3][(k+1)^3]: element node coordinates, lexicographic
// GaussLobatto nodes (component-major as mfem::Ordering::byNODES, mesh.hpp:31-33).
// qdata[ne][11][Q] = {attr, w detJ, (adjJ^T/detJ)[9] column-major}: restatement of
// qfunctions/33/geom_33_qf.h:9-34 with J(q) = sum_n x_n grad phi_n(q) (column-major sdim x dim).
void orc_geom_hex_qdata(int ne, int k, int q1d, const double *xe, const int *attr, double *qdata)
{
  const int n = k + 1, Nn = n * n * n, Q = q1d * q1d * q1d;
  std::vector<double> interp((size_t)Q * Nn), grad((size_t)3 * Q * Nn), qw(Q);
  orc_h1_hex_tables(k, q1d, interp.data(), grad.data(), qw.data());
  for (int e = 0; e < ne; e++)
    for (int q = 0; q < Q; q++)
    {
      double J[9], A[9];
      for (int c = 0; c < 3; c++)      // physical component (row)
        for (int d = 0; d < 3; d++)    // reference direction (column)
        {
          double s = 0;
          for (int m = 0; m < Nn; m++) s += xe[((size_t)e * 3 + c) * Nn + m] * grad[((size_t)d * Q + q) * Nn + m];
          J[c + 3 * d] = s;
        }
      adjJt33(J, A);
      const double detJ = J[0] * A[0] + J[1] * A[1] + J[2] * A[2];
      double *qd = qdata + (size_t)e * 11 * Q;
      qd[0 * Q + q] = (double)attr[e];
      qd[1 * Q + q] = qw[q] * detJ;
      for (int i = 0; i < 9; i++) qd[(2 + i) * Q + q] = A[i] / detJ;
    }
}

// Pointwise D for one batch of Q points (same argument layout as the reference QFunctions:
// qdata[11][Q], u[3][Q], c[3][Q] -> v[3][Q], w[3][Q]). Used to pin against oracle/_ref.
void orc_apply_D(int kind, const void *ctx, int Q, const double *qdata, const double *u, const double *c, double *v,
                 double *w)
{
  for (int q = 0; q < Q; q++)
  {
    double qd[11], ul[3] = {0, 0, 0}, cl[3] = {0, 0, 0}, vl[3], wl[3];
    for (int i = 0; i < 11; i++) qd[i] = qdata[i * Q + q];
    if (u)
      for (int d = 0; d < 3; d++) ul[d] = u[d * Q + q];
    if (c)
      for (int d = 0; d < 3; d++) cl[d] = c[d * Q + q];
    apply_D(kind, (const IntScalar *)ctx, qd, ul, cl, vl, wl);
    if (v)
      for (int d = 0; d < 3; d++) v[d * Q + q] = vl[d];
    if (w)
      for (int d = 0; d < 3; d++) w[d * Q + q] = wl[d];
  }
}

// y_L += A x_L  (ceed::Operator::AddMult semantics, libceed/operator.cpp:148-178,192-212).
// interp[3][Q][P] (ND) or [1][Q][P] replicated as needed; for H1 kinds interp may be null.
// idx[ne][P]; orient[ne][P] in {+1,-1} or null (restriction.cpp:290-297).
void orc_apply_add(int kind, int ne, int P, int Q, const double *interp, const double *deriv, const int *idx,
                   const signed char *orient, const double *qdata, const void *ctx, const double *x, double *y)
{
  const bool need_u = needs_u(kind);
  const bool need_c = needs_c(kind);
  std::vector<double> ue(P), ye(P), u(3 * Q), c(3 * Q), v(3 * Q), w(3 * Q);
  for (int e = 0; e < ne; e++)
  {
    const int *ie = idx + (size_t)e * P;
    const signed char *oe = orient ? orient + (size_t)e * P : nullptr;
    for (int i = 0; i < P; i++) ue[i] = (oe ? (double)oe[i] : 1.0) * x[ie[i]];
    for (int d = 0; d < 3; d++)
      for (int q = 0; q < Q; q++)
      {
        double su = 0, sc = 0;
        if (need_u)
        {
          const double *row = interp + ((size_t)d * Q + q) * P;
          for (int i = 0; i < P; i++) su += row[i] * ue[i];
        }
        if (need_c)
        {
          const double *row = deriv + ((size_t)d * Q + q) * P;
          for (int i = 0; i < P; i++) sc += row[i] * ue[i];
        }
        u[d * Q + q] = su;
        c[d * Q + q] = sc;
      }
    orc_apply_D(kind, ctx, Q, qdata + (size_t)e * 11 * Q, u.data(), c.data(), v.data(), w.data());
    std::fill(ye.begin(), ye.end(), 0.0);
    for (int d = 0; d < 3; d++)
      for (int q = 0; q < Q; q++)
      {
        if (need_u)
        {
          const double *row = interp + ((size_t)d * Q + q) * P;
          const double s = v[d * Q + q];
          for (int i = 0; i < P; i++) ye[i] += row[i] * s;
        }
        if (need_c)
        {
          const double *row = deriv + ((size_t)d * Q + q) * P;
          const double s = w[d * Q + q];
          for (int i = 0; i < P; i++) ye[i] += row[i] * s;
        }
      }
    for (int i = 0; i < P; i++) y[ie[i]] += (oe ? (double)oe[i] : 1.0) * ye[i];
  }
}

// y_L += sum_e E_e^T T_e^T (B^T D B) T_e E_e x_L with the "curl-oriented" restriction of ND tets/prisms at
// p >= 2: T_e is the row-major tridiagonal int8 matrix co[ne][P][3] = {T(i,i-1), T(i,i), T(i,i+1)} that Palace
// fills column by column from the element's DofTransformation (restriction.cpp:301-329) and libCEED applies as
// x_e = T_e x[idx_e] (CeedElemRestrictionCreateCurlOriented), transposed on the way back.
void orc_apply_add_co(int kind, int ne, int P, int Q, const double *interp, const double *deriv, const int *idx,
                      const signed char *co, const double *qdata, const void *ctx, const double *x, double *y)
{
  const bool need_u = needs_u(kind);
  const bool need_c = needs_c(kind);
  std::vector<double> xe(P), ue(P), ye(P), u(3 * Q), c(3 * Q), v(3 * Q), w(3 * Q);
  for (int e = 0; e < ne; e++)
  {
    const int *ie = idx + (size_t)e * P;
    const signed char *ce = co + (size_t)e * P * 3;
    for (int i = 0; i < P; i++) xe[i] = x[ie[i]];
    for (int i = 0; i < P; i++)
    {
      double s = (double)ce[3 * i + 1] * xe[i];
      if (i > 0) s += (double)ce[3 * i + 0] * xe[i - 1];
      if (i < P - 1) s += (double)ce[3 * i + 2] * xe[i + 1];
      ue[i] = s;
    }
    for (int d = 0; d < 3; d++)
      for (int q = 0; q < Q; q++)
      {
        double su = 0, sc = 0;
        if (need_u)
        {
          const double *row = interp + ((size_t)d * Q + q) * P;
          for (int i = 0; i < P; i++) su += row[i] * ue[i];
        }
        if (need_c)
        {
          const double *row = deriv + ((size_t)d * Q + q) * P;
          for (int i = 0; i < P; i++) sc += row[i] * ue[i];
        }
        u[d * Q + q] = su;
        c[d * Q + q] = sc;
      }
    orc_apply_D(kind, ctx, Q, qdata + (size_t)e * 11 * Q, u.data(), c.data(), v.data(), w.data());
    std::fill(ye.begin(), ye.end(), 0.0);
    for (int d = 0; d < 3; d++)
      for (int q = 0; q < Q; q++)
      {
        if (need_u)
        {
          const double *row = interp + ((size_t)d * Q + q) * P;
          const double s = v[d * Q + q];
          for (int i = 0; i < P; i++) ye[i] += row[i] * s;
        }
        if (need_c)
        {
          const double *row = deriv + ((size_t)d * Q + q) * P;
          const double s = w[d * Q + q];
          for (int i = 0; i < P; i++) ye[i] += row[i] * s;
        }
      }
    for (int i = 0; i < P; i++)
    {
      // (T^T y_e)_i = T(i,i) y_i + T(i-1,i) y_{i-1} + T(i+1,i) y_{i+1}
      double s = (double)ce[3 * i + 1] * ye[i];
      if (i > 0) s += (double)ce[3 * (i - 1) + 2] * ye[i - 1];
      if (i < P - 1) s += (double)ce[3 * (i + 1) + 0] * ye[i + 1];
      y[ie[i]] += s;
    }
  }
}

// Element matrices Ae[ne][P][P] (row-major, in the restricted/oriented basis, i.e. including
// the sign flips), for assembling the reference-equivalent sparse matrix in tests.
void orc_element_matrices(int kind, int ne, int P, int Q, const double *interp, const double *deriv,
                          const signed char *orient, const double *qdata, const void *ctx, double *Ae)
{
  const bool need_u = needs_u(kind);
  const bool need_c = needs_c(kind);
  std::vector<double> u(3 * Q), c(3 * Q), v(3 * Q), w(3 * Q);
  for (int e = 0; e < ne; e++)
  {
    const signed char *oe = orient ? orient + (size_t)e * P : nullptr;
    for (int j = 0; j < P; j++)
    {
      for (int d = 0; d < 3; d++)
        for (int q = 0; q < Q; q++)
        {
          u[d * Q + q] = need_u ? interp[((size_t)d * Q + q) * P + j] : 0.0;
          c[d * Q + q] = need_c ? deriv[((size_t)d * Q + q) * P + j] : 0.0;
        }
      orc_apply_D(kind, ctx, Q, qdata + (size_t)e * 11 * Q, u.data(), c.data(), v.data(), w.data());
      for (int i = 0; i < P; i++)
      {
        double s = 0;
        for (int d = 0; d < 3; d++)
          for (int q = 0; q < Q; q++)
          {
            if (need_u) s += interp[((size_t)d * Q + q) * P + i] * v[d * Q + q];
            if (need_c) s += deriv[((size_t)d * Q + q) * P + i] * w[d * Q + q];
          }
        const double sg = oe ? (double)(oe[i] * oe[j]) : 1.0;
        Ae[((size_t)e * P + i) * P + j] = sg * s;
      }
    }
  }
}

// diag_L += diag(E^T B^T D B E)  (ceed::Operator::AssembleDiagonal, libceed/operator.cpp:116-143).
void orc_diag_add(int kind, int ne, int P, int Q, const double *interp, const double *deriv, const int *idx,
                  const double *qdata, const void *ctx, double *diag)
{
  const bool need_u = needs_u(kind);
  const bool need_c = needs_c(kind);
  for (int e = 0; e < ne; e++)
    for (int i = 0; i < P; i++)
    {
      double s = 0;
      for (int q = 0; q < Q; q++)
      {
        double qd[11], ul[3] = {0, 0, 0}, cl[3] = {0, 0, 0}, vl[3], wl[3];
        for (int m = 0; m < 11; m++) qd[m] = qdata[((size_t)e * 11 + m) * Q + q];
        for (int d = 0; d < 3; d++)
        {
          if (need_u) ul[d] = interp[((size_t)d * Q + q) * P + i];
          if (need_c) cl[d] = deriv[((size_t)d * Q + q) * P + i];
        }
        apply_D(kind, (const IntScalar *)ctx, qd, ul, cl, vl, wl);
        for (int d = 0; d < 3; d++) s += ul[d] * vl[d] + cl[d] * wl[d];
      }
      diag[idx[(size_t)e * P + i]] += s;
    }
}

// Multi-threaded variant for the CPU baseline timing (std::thread over contiguous element chunks,
// like the reference's per-thread Ceed contexts, /root/reference/palace/fem/libceed/ceed.cpp:29-40,
// operator.cpp:163-177): threads compute element vectors (E-vector) in parallel, then the
// scatter-add into y is done over disjoint dof ranges in parallel.
void orc_apply_add_mt(int nthreads, int kind, int ne, int P, int Q, const double *interp, const double *deriv,
                      const int *idx, const signed char *orient, const double *qdata, const void *ctx, const double *x,
                      double *y, long long lsize)
{
  if (nthreads <= 1)
  {
    orc_apply_add(kind, ne, P, Q, interp, deriv, idx, orient, qdata, ctx, x, y);
    return;
  }
  const bool need_u = needs_u(kind);
  const bool need_c = needs_c(kind);
  std::vector<double> Ye((size_t)ne * P);
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; t++)
  {
    th.emplace_back(
        [&, t]()
        {
          const int e0 = (int)((long long)ne * t / nthreads), e1 = (int)((long long)ne * (t + 1) / nthreads);
          std::vector<double> ue(P), u(3 * Q), c(3 * Q), v(3 * Q), w(3 * Q);
          for (int e = e0; e < e1; e++)
          {
            const int *ie = idx + (size_t)e * P;
            const signed char *oe = orient ? orient + (size_t)e * P : nullptr;
            double *ye = Ye.data() + (size_t)e * P;
            for (int i = 0; i < P; i++) ue[i] = (oe ? (double)oe[i] : 1.0) * x[ie[i]];
            for (int d = 0; d < 3; d++)
              for (int q = 0; q < Q; q++)
              {
                double su = 0, sc = 0;
                if (need_u)
                {
                  const double *row = interp + ((size_t)d * Q + q) * P;
                  for (int i = 0; i < P; i++) su += row[i] * ue[i];
                }
                if (need_c)
                {
                  const double *row = deriv + ((size_t)d * Q + q) * P;
                  for (int i = 0; i < P; i++) sc += row[i] * ue[i];
                }
                u[d * Q + q] = su;
                c[d * Q + q] = sc;
              }
            orc_apply_D(kind, ctx, Q, qdata + (size_t)e * 11 * Q, u.data(), c.data(), v.data(), w.data());
            for (int i = 0; i < P; i++) ye[i] = 0.0;
            for (int d = 0; d < 3; d++)
              for (int q = 0; q < Q; q++)
              {
                if (need_u)
                {
                  const double *row = interp + ((size_t)d * Q + q) * P;
                  const double sv = v[d * Q + q];
                  for (int i = 0; i < P; i++) ye[i] += row[i] * sv;
                }
                if (need_c)
                {
                  const double *row = deriv + ((size_t)d * Q + q) * P;
                  const double sw = w[d * Q + q];
                  for (int i = 0; i < P; i++) ye[i] += row[i] * sw;
                }
              }
            if (oe)
              for (int i = 0; i < P; i++) ye[i] *= (double)oe[i];
          }
        });
  }
  for (auto &t : th) t.join();
  th.clear();
  // scatter: thread t owns dofs in [lsize*t/nt, lsize*(t+1)/nt)
  const int nts = nthreads > 16 ? 16 : nthreads;
  for (int t = 0; t < nts; t++)
  {
    th.emplace_back(
        [&, t]()
        {
          const long long lo = lsize * t / nts, hi = lsize * (t + 1) / nts;
          const size_t tot = (size_t)ne * P;
          for (size_t k = 0; k < tot; k++)
          {
            const long long g = idx[k];
            if (g >= lo && g < hi) y[g] += Ye[k];
          }
        });
  }
  for (auto &t : th) t.join();
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------
// Reference arm of bench.py (`--impl reference`, `cpu_baseline`): the same dense-basis algorithm as orc_apply_add,
// organised the way libCEED's /cpu/self/opt/blocked backend runs it on the host -- blocks of 8 elements with the
// element index innermost (one SIMD lane per element), a PERSISTENT pool of threads pinned one per physical core,
// and a precomputed transposed restriction (dof -> E-vector entries, libCEED's t_offsets) so that the scatter-add is a
// race-free parallel loop over dofs. Same arithmetic per element as orc_apply_add (verified against it in
// tests/test_oracle_identities.py); only the loop order and the threading differ.
// ---------------------------------------------------------------------------------------------------------------
namespace
{
constexpr int BLK = 8;
typedef double vblk __attribute__((vector_size(BLK * sizeof(double)), aligned(8)));  // one element lane per SIMD lane

struct OrcPool
{
  int nt = 0;
  std::vector<std::thread> th;
  std::mutex mu;
  std::condition_variable cv_go, cv_done;
  std::function<void(int)> job;
  long long epoch = 0;
  int pending = 0;
  bool stop = false;
  void run(const std::function<void(int)> &f)
  {
    std::unique_lock<std::mutex> lk(mu);
    job = f;
    pending = nt;
    epoch++;
    cv_go.notify_all();
    cv_done.wait(lk, [&] { return pending == 0; });
  }
};

std::vector<int> physical_core_cpus()
{
  // one logical CPU per (package, core) pair, among the CPUs this process may run on
  std::vector<int> out;
  cpu_set_t allowed;
  CPU_ZERO(&allowed);
  if (sched_getaffinity(0, sizeof(allowed), &allowed) != 0)
  {
    const unsigned n = std::thread::hardware_concurrency();
    for (unsigned c = 0; c < (n ? n : 1); c++) out.push_back((int)c);
    return out;
  }
  std::vector<std::pair<int, int>> seen;
  for (int c = 0; c < CPU_SETSIZE; c++)
  {
    if (!CPU_ISSET(c, &allowed)) continue;
    int pkg = 0, core = c;
    char path[128];
    snprintf(path, sizeof(path), "/sys/devices/system/cpu/cpu%d/topology/physical_package_id", c);
    if (FILE *f = fopen(path, "r"))
    {
      if (fscanf(f, "%d", &pkg) != 1) pkg = 0;
      fclose(f);
    }
    snprintf(path, sizeof(path), "/sys/devices/system/cpu/cpu%d/topology/core_id", c);
    if (FILE *f = fopen(path, "r"))
    {
      if (fscanf(f, "%d", &core) != 1) core = c;
      fclose(f);
    }
    bool dup = false;
    for (auto &pc : seen) dup = dup || (pc.first == pkg && pc.second == core);
    if (!dup)
    {
      seen.emplace_back(pkg, core);
      out.push_back(c);
    }
  }
  if (out.empty()) out.push_back(0);
  return out;
}

struct OrcBlocked
{
  int ne = 0, P = 0;
  long long lsize = 0;
  std::vector<long long> dof_ptr;   // [lsize + 1]
  std::vector<int> dof_ent;         // E-vector positions e * P + i, grouped by dof
  std::vector<double> Ye;           // [ne][P] element vectors
};
}  // namespace

extern "C"
{

int orc_physical_cores() { return (int)physical_core_cpus().size(); }

// nthreads <= 0: one thread per physical core, pinned.
void *orc_pool_create(int nthreads)
{
  const std::vector<int> cpus = physical_core_cpus();
  auto *p = new OrcPool;
  p->nt = nthreads > 0 ? nthreads : (int)cpus.size();
  for (int t = 0; t < p->nt; t++)
  {
    p->th.emplace_back(
        [p, t]()
        {
          long long seen = 0;
          for (;;)
          {
            std::function<void(int)> f;
            {
              std::unique_lock<std::mutex> lk(p->mu);
              p->cv_go.wait(lk, [&] { return p->stop || p->epoch != seen; });
              if (p->stop) return;
              seen = p->epoch;
              f = p->job;
            }
            f(t);
            {
              std::unique_lock<std::mutex> lk(p->mu);
              if (--p->pending == 0) p->cv_done.notify_all();
            }
          }
        });
    cpu_set_t set;
    CPU_ZERO(&set);
    CPU_SET(cpus[t % cpus.size()], &set);
    pthread_setaffinity_np(p->th.back().native_handle(), sizeof(set), &set);
  }
  return p;
}
int orc_pool_threads(void *pool) { return ((OrcPool *)pool)->nt; }
void orc_pool_destroy(void *pool)
{
  auto *p = (OrcPool *)pool;
  if (!p) return;
  {
    std::unique_lock<std::mutex> lk(p->mu);
    p->stop = true;
    p->cv_go.notify_all();
  }
  for (auto &t : p->th) t.join();
  delete p;
}

// Set-up (untimed, like CeedElemRestrictionCreate): transposed restriction + the E-vector.
void *orc_blocked_setup(int ne, int P, const int *idx, long long lsize)
{
  auto *b = new OrcBlocked;
  b->ne = ne;
  b->P = P;
  b->lsize = lsize;
  b->dof_ptr.assign((size_t)lsize + 1, 0);
  const size_t tot = (size_t)ne * P;
  for (size_t k = 0; k < tot; k++) b->dof_ptr[(size_t)idx[k] + 1]++;
  for (long long g = 0; g < lsize; g++) b->dof_ptr[g + 1] += b->dof_ptr[g];
  b->dof_ent.resize(tot);
  std::vector<long long> fill(b->dof_ptr.begin(), b->dof_ptr.end() - 1);
  for (size_t k = 0; k < tot; k++) b->dof_ent[(size_t)fill[idx[k]]++] = (int)k;
  b->Ye.assign(tot, 0.0);
  return b;
}
void orc_blocked_destroy(void *h) { delete (OrcBlocked *)h; }

// y += A x, blocks of BLK elements handed out dynamically to the pool's threads.
void orc_apply_add_blocked(void *pool, void *setup, int kind, int ne, int P, int Q, const double *interp, const double *deriv,
                           const int *idx, const signed char *orient, const double *qdata, const void *ctx, const double *x,
                           double *y)
{
  auto *pl = (OrcPool *)pool;
  auto *bs = (OrcBlocked *)setup;
  const bool need_u = needs_u(kind);
  const bool need_c = needs_c(kind);
  const int nblk = (ne + BLK - 1) / BLK, R = 3 * Q;
  std::atomic<int> next{0};
  pl->run(
      [&](int)
      {
        std::vector<double> ue((size_t)P * BLK), uq((size_t)R * BLK), cq((size_t)R * BLK), vq((size_t)R * BLK), wq((size_t)R * BLK),
            ye((size_t)P * BLK), u1(R), c1(R), v1(R), w1(R);
        for (;;)
        {
          const int blk = next.fetch_add(2);  // two blocks per grab (a preempted thread holds back at most 16 elements)
          if (blk >= nblk) break;
          for (int bb = blk; bb < std::min(blk + 2, nblk); bb++)
          {
            const int e0 = bb * BLK, nel = std::min(BLK, ne - e0);
            // gather with signs, element index innermost
            for (int i = 0; i < P; i++)
              for (int l = 0; l < BLK; l++)
              {
                const int e = e0 + (l < nel ? l : 0);
                const double sg = orient ? (double)orient[(size_t)e * P + i] : 1.0;
                ue[(size_t)i * BLK + l] = l < nel ? sg * x[idx[(size_t)e * P + i]] : 0.0;
              }
            // u = interp ue, c = deriv ue: four rows of the dense tables at a time against the BLK element lanes held as ONE
            // SIMD vector (GCC vector extension: eight independent register accumulators cover the FMA latency; every ue
            // vector is loaded once per four rows)
            for (int r0 = 0; r0 < R; r0 += 4)
            {
              const int nr = std::min(4, R - r0);
              vblk au[4], ac[4];
              for (int k = 0; k < 4; k++) au[k] = ac[k] = vblk{};
              const double *ri[4], *rd[4];
              for (int k = 0; k < 4; k++)
              {
                ri[k] = interp + (size_t)(r0 + std::min(k, nr - 1)) * P;
                rd[k] = deriv + (size_t)(r0 + std::min(k, nr - 1)) * P;
              }
              const vblk *uev = reinterpret_cast<const vblk *>(ue.data());
              if (need_u && need_c)
                for (int i = 0; i < P; i++)
                {
                  const vblk ul = uev[i];
                  for (int k = 0; k < 4; k++)
                  {
                    au[k] += ri[k][i] * ul;
                    ac[k] += rd[k][i] * ul;
                  }
                }
              else if (need_u)
                for (int i = 0; i < P; i++)
                {
                  const vblk ul = uev[i];
                  for (int k = 0; k < 4; k++) au[k] += ri[k][i] * ul;
                }
              else
                for (int i = 0; i < P; i++)
                {
                  const vblk ul = uev[i];
                  for (int k = 0; k < 4; k++) ac[k] += rd[k][i] * ul;
                }
              for (int k = 0; k < nr; k++)
              {
                *reinterpret_cast<vblk *>(&uq[(size_t)(r0 + k) * BLK]) = au[k];
                *reinterpret_cast<vblk *>(&cq[(size_t)(r0 + k) * BLK]) = ac[k];
              }
            }
            // pointwise D, element by element (the reference's QFunction arithmetic)
            for (int l = 0; l < nel; l++)
            {
              for (int r = 0; r < R; r++)
              {
                u1[r] = uq[(size_t)r * BLK + l];
                c1[r] = cq[(size_t)r * BLK + l];
              }
              orc_apply_D(kind, ctx, Q, qdata + (size_t)(e0 + l) * 11 * Q, u1.data(), c1.data(), v1.data(), w1.data());
              for (int r = 0; r < R; r++)
              {
                vq[(size_t)r * BLK + l] = need_u ? v1[r] : 0.0;
                wq[(size_t)r * BLK + l] = need_c ? w1[r] : 0.0;
              }
            }
            for (int l = nel; l < BLK; l++)
              for (int r = 0; r < R; r++) vq[(size_t)r * BLK + l] = wq[(size_t)r * BLK + l] = 0.0;
            // ye = interp^T v + deriv^T w: eight output dofs at a time held in registers across the whole row loop
            {
              const vblk *vv = reinterpret_cast<const vblk *>(vq.data()), *wv = reinterpret_cast<const vblk *>(wq.data());
              for (int i0 = 0; i0 < P; i0 += 8)
              {
                const int ni = std::min(8, P - i0);
                vblk acc[8];
                for (int k = 0; k < 8; k++) acc[k] = vblk{};
                for (int r = 0; r < R; r++)
                {
                  const double *ri = interp + (size_t)r * P + i0, *rd = deriv + (size_t)r * P + i0;
                  const vblk vl = vv[r], wl = wv[r];
                  if (need_u && need_c)
                    for (int k = 0; k < 8; k++)
                    {
                      const int kk = std::min(k, ni - 1);
                      acc[k] += ri[kk] * vl + rd[kk] * wl;
                    }
                  else if (need_u)
                    for (int k = 0; k < 8; k++) acc[k] += ri[std::min(k, ni - 1)] * vl;
                  else
                    for (int k = 0; k < 8; k++) acc[k] += rd[std::min(k, ni - 1)] * wl;
                }
                for (int k = 0; k < ni; k++) *reinterpret_cast<vblk *>(&ye[(size_t)(i0 + k) * BLK]) = acc[k];
              }
            }
            for (int l = 0; l < nel; l++)
              for (int i = 0; i < P; i++)
              {
                const size_t k = (size_t)(e0 + l) * P + i;
                bs->Ye[k] = (orient ? (double)orient[k] : 1.0) * ye[(size_t)i * BLK + l];
              }
          }
        }
      });
  // transposed restriction: every dof sums its E-vector entries (no races, fixed order)
  const long long lsize = bs->lsize, CH = 8192;
  std::atomic<long long> nextd{0};
  pl->run(
      [&](int)
      {
        for (;;)
        {
          const long long lo = nextd.fetch_add(CH);
          if (lo >= lsize) break;
          const long long hi = std::min(lsize, lo + CH);
          for (long long g = lo; g < hi; g++)
          {
            double sacc = 0.0;
            for (long long k = bs->dof_ptr[g]; k < bs->dof_ptr[g + 1]; k++) sacc += bs->Ye[bs->dof_ent[k]];
            y[g] += sacc;
          }
        }
      });
}

}  // extern "C"
