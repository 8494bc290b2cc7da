// Pointwise quadrature-point operators D (device). Same linear maps as the reference QFunctions
//   /root/reference/palace/fem/qfunctions/33/hcurl_33_qf.h:10-29   v = w detJ * A^T C A u        (A = adjJ^T/detJ = J^-T)
//   /root/reference/palace/fem/qfunctions/33/hdiv_33_qf.h:10-30    v = w detJ * Jd^T C Jd c      (Jd = J/detJ = cofactor(A))
//   /root/reference/palace/fem/qfunctions/33/hdivmass_33_qf.h:10-44  both
// written for registers; all 3x3 matrices column-major ("0 3 6 / 1 4 7 / 2 5 8").
#pragma once

namespace b2p
{

// y = A^T (C (A x)) * s
__device__ __forceinline__ void AtCAx(const double A[9], const double C[9], const double x[3], double s,
                                      double y[3])
{
  const double t0 = A[0] * x[0] + A[3] * x[1] + A[6] * x[2];
  const double t1 = A[1] * x[0] + A[4] * x[1] + A[7] * x[2];
  const double t2 = A[2] * x[0] + A[5] * x[1] + A[8] * x[2];
  const double z0 = C[0] * t0 + C[3] * t1 + C[6] * t2;
  const double z1 = C[1] * t0 + C[4] * t1 + C[7] * t2;
  const double z2 = C[2] * t0 + C[5] * t1 + C[8] * t2;
  y[0] = s * (A[0] * z0 + A[1] * z1 + A[2] * z2);
  y[1] = s * (A[3] * z0 + A[4] * z1 + A[5] * z2);
  y[2] = s * (A[6] * z0 + A[7] * z1 + A[8] * z2);
}

// y = s * A^T (A x)   (isotropic coefficient: C = c I folded into s)
__device__ __forceinline__ void AtAx(const double A[9], const double x[3], double s, double y[3])
{
  const double t0 = A[0] * x[0] + A[3] * x[1] + A[6] * x[2];
  const double t1 = A[1] * x[0] + A[4] * x[1] + A[7] * x[2];
  const double t2 = A[2] * x[0] + A[5] * x[1] + A[8] * x[2];
  y[0] = s * (A[0] * t0 + A[1] * t1 + A[2] * t2);
  y[1] = s * (A[3] * t0 + A[4] * t1 + A[5] * t2);
  y[2] = s * (A[6] * t0 + A[7] * t1 + A[8] * t2);
}

// t = A x ; y = s * A^T z   (the two halves of AtCAx, for coefficients applied in between)
__device__ __forceinline__ void Ax33(const double A[9], const double x[3], double t[3])
{
  t[0] = A[0] * x[0] + A[3] * x[1] + A[6] * x[2];
  t[1] = A[1] * x[0] + A[4] * x[1] + A[7] * x[2];
  t[2] = A[2] * x[0] + A[5] * x[1] + A[8] * x[2];
}
__device__ __forceinline__ void Atx33(const double A[9], const double z[3], double s, double y[3])
{
  y[0] = s * (A[0] * z[0] + A[1] * z[1] + A[2] * z[2]);
  y[1] = s * (A[3] * z[0] + A[4] * z[1] + A[5] * z[2]);
  y[2] = s * (A[6] * z[0] + A[7] * z[1] + A[8] * z[2]);
}

// Cofactor matrix (adj^T) of a column-major 3x3: for A = J^-T this is J/detJ (utils_33_qf.h:20-37).
__device__ __forceinline__ void cofactor33(const double J[9], double A[9])
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

// S = s * A^T C A (full 3x3, column-major) -- the reference's "build" QFunctions
// (hcurl_build_33_qf.h / hdiv_build_33_qf.h via MultAtBA33, utils_33_qf.h:112-139).
__device__ __forceinline__ void AtCA(const double A[9], const double C[9], double s, double S[9])
{
  double R[9];  // R = C A
  for (int c = 0; c < 3; c++)
    for (int r = 0; r < 3; r++) R[r + 3 * c] = C[r] * A[3 * c] + C[r + 3] * A[3 * c + 1] + C[r + 6] * A[3 * c + 2];
  for (int c = 0; c < 3; c++)
    for (int r = 0; r < 3; r++)
      S[r + 3 * c] = s * (A[3 * r] * R[3 * c] + A[3 * r + 1] * R[3 * c + 1] + A[3 * r + 2] * R[3 * c + 2]);
}

}  // namespace b2p
