// eig3.cuh — symmetric 3x3 eigen-decomposition in float64 for the normal/planarity kernel.
// Reference: np.linalg.eig on np.cov output (python/simpleicp/pointcloud.py:190-198).
#pragma once
#include <math.h>

#include "../../include/sicp_b200.h"

namespace sicp {

// Cyclic Jacobi on a symmetric 3x3 matrix.  a is destroyed (its diagonal becomes the
// eigenvalues), v receives the eigenvectors as columns.
__host__ __device__ inline void eig3_jacobi(double a[3][3], double v[3][3]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) v[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 64; ++sweep) {
    const double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
    const double dia = a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2];
    if (off <= 1e-34 * dia || off == 0.0) break;
    for (int p = 0; p < 2; ++p) {
      for (int q = p + 1; q < 3; ++q) {
        const double apq = a[p][q];
        if (apq == 0.0) continue;
        const double theta = (a[q][q] - a[p][p]) / (2.0 * apq);
        double t;
        if (fabs(theta) > 1e150) {
          t = 0.5 / theta;
        } else {
          t = 1.0 / (fabs(theta) + sqrt(theta * theta + 1.0));
          if (theta < 0.0) t = -t;
        }
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        const int r = 3 - p - q;  // the third index
        const double app = a[p][p], aqq = a[q][q];
        a[p][p] = app - t * apq;
        a[q][q] = aqq + t * apq;
        a[p][q] = a[q][p] = 0.0;
        const double arp = a[r][p], arq = a[r][q];
        a[r][p] = a[p][r] = c * arp - s * arq;
        a[r][q] = a[q][r] = s * arp + c * arq;
        for (int i = 0; i < 3; ++i) {
          const double vip = v[i][p], viq = v[i][q];
          v[i][p] = c * vip - s * viq;
          v[i][q] = s * vip + c * viq;
        }
      }
    }
  }
}

// Eigenvalues sorted descending in w; n = unit eigenvector of the smallest eigenvalue.
__host__ __device__ inline void eig3_smallest(double c00, double c01, double c02, double c11,
                                              double c12, double c22, int sign_mode, double w[3],
                                              double n[3]) {
  double scale = fmax(fmax(fabs(c00), fabs(c11)), fmax(fabs(c22), fmax(fabs(c01), fmax(fabs(c02), fabs(c12)))));
  if (!(scale > 0.0) || !isfinite(scale)) {
    // degenerate neighbourhood (all points coincide) or non-finite input: the reference's
    // planarity is 0/0 = NaN, which every rejection test drops.
    w[0] = w[1] = w[2] = (scale == 0.0) ? 0.0 : scale;
    n[0] = 1.0;
    n[1] = n[2] = 0.0;
    return;
  }
  const double is = 1.0 / scale;
  double a[3][3] = {{c00 * is, c01 * is, c02 * is}, {c01 * is, c11 * is, c12 * is}, {c02 * is, c12 * is, c22 * is}};
  double v[3][3];
  eig3_jacobi(a, v);
  int i0 = 0, i1 = 1, i2 = 2;  // descending order of a[i][i]
  if (a[i0][i0] < a[i1][i1]) { int t = i0; i0 = i1; i1 = t; }
  if (a[i1][i1] < a[i2][i2]) { int t = i1; i1 = i2; i2 = t; }
  if (a[i0][i0] < a[i1][i1]) { int t = i0; i0 = i1; i1 = t; }
  w[0] = a[i0][i0] * scale;
  w[1] = a[i1][i1] * scale;
  w[2] = a[i2][i2] * scale;
  double x = v[0][i2], y = v[1][i2], z = v[2][i2];
  const double nn = 1.0 / sqrt(x * x + y * y + z * z);
  x *= nn;
  y *= nn;
  z *= nn;
  (void)sign_mode;
  // canonical sign: the component of largest magnitude is positive
  const double ax = fabs(x), ay = fabs(y), az = fabs(z);
  const double lead = (ax >= ay && ax >= az) ? x : ((ay >= az) ? y : z);
  if (lead < 0.0) {
    x = -x;
    y = -y;
    z = -z;
  }
  n[0] = x;
  n[1] = y;
  n[2] = z;
}

}  // namespace sicp
