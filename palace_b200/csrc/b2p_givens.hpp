// Plane (Givens) rotations of the GMRES least-squares update, host side (plain C++: also compiled alone by
// tests/test_givens_cpu.py).
#pragma once
#include <algorithm>
#include <cmath>
#include <complex>
#include <limits>

namespace b2p
{
using cplx = std::complex<double>;

// Complex Givens rotation [cs sn; -conj(sn) cs] (dx, dy)^T = (r, 0)^T with real cs, after LAPACK 3.10's zlartg (the algorithm
// the reference restates in iterative.cpp:112-226): the rotation is formed from |dx|^2 and |dy|^2 directly when both lie in
// the range whose squares neither underflow nor overflow, and from operands scaled into that range otherwise -- dx with its
// own scale when it is far smaller than dy, which the cosine then carries as a factor. (Round 1 had the unscaled branch only.)
inline void GeneratePlaneRotation(const cplx dx, const cplx dy, double &cs, cplx &sn)
{
  using lim = std::numeric_limits<double>;
  static const double safmin = std::ldexp(1.0, std::max(lim::min_exponent - 1, 1 - lim::max_exponent));
  static const double safmax = 1.0 / safmin;
  auto absq = [](const cplx z) { return z.real() * z.real() + z.imag() * z.imag(); };
  auto inf_norm = [](const cplx z) { return std::max(std::abs(z.real()), std::abs(z.imag())); };
  if (dy == 0.0)
  {
    cs = 1.0;
    sn = 0.0;
    return;
  }
  const double gmax = inf_norm(dy);
  if (dx == 0.0)
  {
    // r = |dy|: only the phase of dy is left in sn; |dy| from scaled parts when its square would leave the safe range
    cs = 0.0;
    const double lo = std::sqrt(safmin), hi = std::sqrt(safmax / 2);
    if (dy.real() == 0.0 || dy.imag() == 0.0)
      sn = std::conj(dy) / gmax;
    else if (gmax > lo && gmax < hi)
      sn = std::conj(dy) / std::sqrt(absq(dy));
    else
    {
      const cplx g = dy / std::min(safmax, std::max(safmin, gmax));
      sn = std::conj(g) / std::sqrt(absq(g));
    }
    return;
  }
  const double lo = std::sqrt(safmin), hi = std::sqrt(safmax / 4);
  const double fmax = inf_norm(dx);
  // the rotation from (scaled) f, g with f2 = |f|^2 and h2 = f2 (* w^2) + |g|^2
  auto form = [&](const cplx f, const cplx g, double f2, double h2)
  {
    if (f2 >= h2 * safmin)
    {
      cs = std::sqrt(f2 / h2);
      sn = (f2 > lo && h2 < 2 * hi) ? std::conj(g) * (f / std::sqrt(f2 * h2)) : std::conj(g) * ((f / cs) / h2);
    }
    else
    {
      const double d = std::sqrt(f2 * h2);  // f2 / h2 would underflow
      cs = f2 / d;
      sn = std::conj(g) * (f / d);
    }
  };
  if (fmax > lo && fmax < hi && gmax > lo && gmax < hi)
  {
    const double f2 = absq(dx);
    form(dx, dy, f2, f2 + absq(dy));
    return;
  }
  const double u = std::min(safmax, std::max(safmin, std::max(fmax, gmax)));
  const cplx g = dy / u;
  const double g2 = absq(g);
  if (fmax / u < lo)
  {
    // dx is negligible at dy's scale: give it its own scale v and account for w = v / u afterwards
    const double v = std::min(safmax, std::max(safmin, fmax)), w = v / u;
    const cplx f = dx / v;
    const double f2 = absq(f);
    form(f, g, f2, f2 * w * w + g2);
    cs *= w;
  }
  else
  {
    const cplx f = dx / u;
    const double f2 = absq(f);
    form(f, g, f2, f2 + g2);
  }
}

}  // namespace b2p
