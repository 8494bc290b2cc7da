// TEST INFRASTRUCTURE ONLY (CPU oracle). Nothing under oracle/ is shipped or timed as product.
//
// 1-D point sets and Lagrange bases used by the tensor-product elements the reference
// builds through MFEM (un-vendored; pinned tag d9d6526c in /root/reference/cmake/ExternalGitTags.cmake):
//   * ND hex = ND_HexahedronElement(p, GaussLobatto closed, GaussLegendre open)
//     (/root/reference/palace/fem/multigrid.hpp:35,49)
//   * H1 hex = H1_HexahedronElement(p, GaussLobatto)  (multigrid.hpp:31-33)
//   * hex quadrature of order 2p = Gauss-Legendre (p+1)^3  (/root/reference/palace/fem/integrator.cpp:14-22)
// Everything is evaluated in long double and rounded once, so tables are correct to the last
// bit or so of double, independent of how MFEM's Poly_1D evaluates them.
#pragma once
#include <cmath>
#include <vector>

namespace orc
{

using ld = long double;

// Legendre P_n(x) and P_n'(x) on [-1,1].
inline void legendre(int n, ld x, ld &P, ld &dP)
{
  if (n == 0)
  {
    P = 1;
    dP = 0;
    return;
  }
  ld p0 = 1, p1 = x;
  for (int k = 2; k <= n; k++)
  {
    ld pk = ((2 * k - 1) * x * p1 - (k - 1) * p0) / k;
    p0 = p1;
    p1 = pk;
  }
  P = p1;
  dP = n * (x * p1 - p0) / (x * x - 1);
}

// n-point Gauss-Legendre rule mapped to [0,1], ascending.
inline void gauss_legendre(int n, std::vector<ld> &x, std::vector<ld> &w)
{
  x.assign(n, 0);
  w.assign(n, 0);
  const ld pi = acosl(-1.0L);
  for (int i = 0; i < n; i++)
  {
    ld z = -cosl(pi * (i + 0.75L) / (n + 0.5L));
    for (int it = 0; it < 100; it++)
    {
      ld P, dP;
      legendre(n, z, P, dP);
      ld dz = P / dP;
      z -= dz;
      if (fabsl(dz) < 1e-20L)
        break;
    }
    ld P, dP;
    legendre(n, z, P, dP);
    x[i] = 0.5L * (z + 1);
    w[i] = 1 / ((1 - z * z) * dP * dP);  // (2/((1-z^2)dP^2)) * 1/2 for [0,1]
  }
  // Symmetrise.
  for (int i = 0; i < n / 2; i++)
  {
    ld a = 0.5L * (x[i] + (1 - x[n - 1 - i]));
    x[i] = a;
    x[n - 1 - i] = 1 - a;
    ld b = 0.5L * (w[i] + w[n - 1 - i]);
    w[i] = w[n - 1 - i] = b;
  }
  if (n % 2)
    x[n / 2] = 0.5L;
}

// n-point Gauss-Lobatto nodes mapped to [0,1] (n >= 2), ascending: endpoints + roots of P'_{n-1}.
inline void gauss_lobatto(int n, std::vector<ld> &x)
{
  x.assign(n, 0);
  x[0] = 0;
  x[n - 1] = 1;
  const int m = n - 1;
  const ld pi = acosl(-1.0L);
  for (int i = 1; i < n - 1; i++)
  {
    ld z = -cosl(pi * i / m);
    for (int it = 0; it < 100; it++)
    {
      // f = P'_m, f' = P''_m from (1-z^2)P'' = 2zP' - m(m+1)P
      ld P, dP;
      legendre(m, z, P, dP);
      ld d2P = (2 * z * dP - m * (m + 1) * P) / (1 - z * z);
      ld dz = dP / d2P;
      z -= dz;
      if (fabsl(dz) < 1e-20L)
        break;
    }
    x[i] = 0.5L * (z + 1);
  }
  for (int i = 0; i < n / 2; i++)
  {
    ld a = 0.5L * (x[i] + (1 - x[n - 1 - i]));
    x[i] = a;
    x[n - 1 - i] = 1 - a;
  }
  if (n % 2)
    x[n / 2] = 0.5L;
}

// Lagrange basis through `nodes` at point t: values l[j], derivatives dl[j].
inline void lagrange(const std::vector<ld> &nodes, ld t, std::vector<ld> &l, std::vector<ld> &dl)
{
  const int n = (int)nodes.size();
  l.assign(n, 0);
  dl.assign(n, 0);
  for (int j = 0; j < n; j++)
  {
    ld den = 1;
    for (int m = 0; m < n; m++)
      if (m != j)
        den *= (nodes[j] - nodes[m]);
    ld val = 1;
    for (int m = 0; m < n; m++)
      if (m != j)
        val *= (t - nodes[m]);
    ld der = 0;
    for (int k = 0; k < n; k++)
    {
      if (k == j)
        continue;
      ld pr = 1;
      for (int m = 0; m < n; m++)
        if (m != j && m != k)
          pr *= (t - nodes[m]);
      der += pr;
    }
    l[j] = val / den;
    dl[j] = der / den;
  }
}

// Table T[q][j] = basis_j(pts[q]) and D[q][j] = basis_j'(pts[q]).
struct Table1D
{
  int nq = 0, nd = 0;
  std::vector<ld> B, G;
  ld b(int q, int j) const { return B[q * nd + j]; }
  ld g(int q, int j) const { return G[q * nd + j]; }
};

inline Table1D make_table(const std::vector<ld> &nodes, const std::vector<ld> &pts)
{
  Table1D t;
  t.nq = (int)pts.size();
  t.nd = (int)nodes.size();
  t.B.resize(t.nq * t.nd);
  t.G.resize(t.nq * t.nd);
  std::vector<ld> l, dl;
  for (int q = 0; q < t.nq; q++)
  {
    lagrange(nodes, pts[q], l, dl);
    for (int j = 0; j < t.nd; j++)
    {
      t.B[q * t.nd + j] = l[j];
      t.G[q * t.nd + j] = dl[j];
    }
  }
  return t;
}

}  // namespace orc
