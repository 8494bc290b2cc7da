"""CPU: the complex Givens rotation of the GMRES update (palace_b200/csrc/b2p_givens.hpp, LAPACK zlartg's safe scaling as the
reference restates it, linalg/iterative.cpp:112-226) against extended-precision arithmetic, including operands whose squares
underflow or overflow in double precision and operands of wildly different magnitude."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r"""
#include <cstdio>
#include <complex>
#include "b2p_givens.hpp"
using b2p::cplx;
typedef long double ld;
typedef std::complex<ld> lc;
int main()
{
  const double mags[] = {1e-300, 1e-170, 1e-155, 1e-20, 1.0, 3.7e10, 1e150, 1e160, 1e290};
  const cplx dirs[] = {cplx(1, 0), cplx(0, 1), cplx(0.6, -0.8), cplx(-0.28, 0.96), cplx(1e-9, 1)};
  double worst = 0;
  int n = 0;
  for (double ma : mags)
    for (double mb : mags)
      for (const cplx &da : dirs)
        for (const cplx &db : dirs)
        {
          const cplx dx = ma * da, dy = mb * db;
          double cs;
          cplx sn;
          b2p::GeneratePlaneRotation(dx, dy, cs, sn);
          // extended-precision reference on operands scaled by the larger magnitude: c = |f| / sqrt(|f|^2 + |g|^2),
          // s = (f / |f|) conj(g) / sqrt(|f|^2 + |g|^2)
          const ld u = std::max((ld)ma, (ld)mb);
          const lc f = lc(dx.real(), dx.imag()) / u, g = lc(dy.real(), dy.imag()) / u;
          const ld h = std::sqrt(std::norm(f) + std::norm(g)), af = std::abs(f);
          const ld c_ref = af / h;
          const lc s_ref = (f / af) * std::conj(g) / h;
          const double e = std::max(std::abs((double)(cs - c_ref)), (double)std::abs(lc(sn.real(), sn.imag()) - s_ref));
          // the rotation annihilates dy and is unitary
          const lc r2 = -std::conj(lc(sn.real(), sn.imag())) * f + (ld)cs * g;
          const double e2 = (double)(std::abs(r2) / h);
          const double e3 = std::abs(cs * cs + std::norm(sn) - 1.0);
          worst = std::max(worst, std::max(e, std::max(e2, e3)));
          n++;
        }
  // zero operands
  double cs;
  cplx sn;
  b2p::GeneratePlaneRotation(cplx(2, 1), cplx(0, 0), cs, sn);
  if (cs != 1.0 || sn != cplx(0, 0)) return 2;
  b2p::GeneratePlaneRotation(cplx(0, 0), cplx(3e-200, -4e-200), cs, sn);
  if (cs != 0.0 || std::abs(sn - cplx(0.6, 0.8)) > 1e-15) return 3;
  std::printf("GIVENS %s cases=%d worst=%.3e\n", worst < 1e-14 ? "OK" : "FAIL", n, worst);
  return worst < 1e-14 ? 0 : 1;
}
"""


def test_complex_givens_rotation_is_accurate_over_the_whole_exponent_range(tmp_path):
    src = tmp_path / "givens.cpp"
    src.write_text(SRC)
    exe = str(tmp_path / "givens")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "palace_b200", "csrc"), str(src), "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "GIVENS OK" in r.stdout, r.stdout + r.stderr
