// Host shim for tests: exposes the header-only eigen-solver of the normal kernel to ctypes.
#include "../../simpleicp_b200/csrc/eig3.cuh"

extern "C" {
int eig3_dgeev_host(const double* A, double* wr, double* vr) {
  double a[3][3], v[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) a[i][j] = A[i * 3 + j];
  const bool ok = sicp::eig3_dgeev(a, wr, v);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) vr[i * 3 + j] = v[i][j];
  return ok ? 1 : 0;
}
void eig3_smallest_host(const double* c6, int sign_mode, double* w, double* n) {
  sicp::eig3_smallest(c6[0], c6[1], c6[2], c6[3], c6[4], c6[5], sign_mode, w, n);
}
}
