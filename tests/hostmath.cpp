// Host build of slam3d_b200/csrc/gicp_math.h for the CPU test-suite (tests/test_hostmath.py).
// Test-only: lets the optimiser math that runs on one GPU thread be checked against the oracle without a GPU.
#include "../slam3d_b200/csrc/gicp_math.h"

extern "C" {
double hm_f(const double* mom, const double* x) { return s3d::moments_f(mom, x); }
void hm_dfddf(const double* mom, const double* x, double* g, double* H) {
  double gg[6], HH[6][6];
  s3d::moments_dfddf(mom, x, gg, HH);
  for (int i = 0; i < 6; ++i) { g[i] = gg[i]; for (int j = 0; j < 6; ++j) H[j * 6 + i] = HH[i][j]; }
}
int hm_newton(const double* mom, float* T, int max_inner, int* inner_done) {
  int d = 0;
  bool ok = s3d::newton_from_moments(mom, T, max_inner, &d);
  *inner_done = d;
  return ok ? 0 : 2;
}
void hm_direction(const double* H, const double* g, double* delta) {
  double HH[6][6], gg[6], dd[6];
  for (int i = 0; i < 6; ++i) { gg[i] = g[i]; for (int j = 0; j < 6; ++j) HH[i][j] = H[j * 6 + i]; }
  s3d::newton_direction(HH, gg, dd);
  for (int i = 0; i < 6; ++i) delta[i] = dd[i];
}
void hm_normal(const double* cov, double* n) {
  double c[3][3], nn[3];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) c[i][j] = cov[j * 3 + i];
  s3d::smallest_eigenvector3(c, nn);
  for (int i = 0; i < 3; ++i) n[i] = nn[i];
}
void hm_mahalanobis(const double* RRt, const double* a, const double* b, double* M6) {
  double r[3][3], m[6];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r[i][j] = RRt[j * 3 + i];
  s3d::mahalanobis6(r, a, b, m);
  for (int i = 0; i < 6; ++i) M6[i] = m[i];
}
int hm_sym3(int a, int b) { return s3d::sym3(a, b); }
int hm_sym4(int a, int b) { return s3d::sym4(a, b); }
}
