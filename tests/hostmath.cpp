// Host build of slam3d_b200/csrc/gicp_math.h for the CPU test-suite (tests/test_hostmath.py).
// Test-only: lets the optimiser math that runs on the GPU be checked against the oracle without a GPU.  The evaluation
// pass below does on the host what gicp_iter_kernel / gicp_eval_kernel do on the device (same float operation order).
#include <vector>

#include "../slam3d_b200/csrc/gicp_math.h"

namespace {
// Eigen Matrix4f * Vector4f with w = 1 (same order as common.cuh transform_mv); built with -ffp-contract=off
inline void transform_mv(const float* T, const float* p, float* o) {
  o[0] = ((T[0] * p[0] + T[4] * p[1]) + T[8] * p[2]) + T[12];
  o[1] = ((T[1] * p[0] + T[5] * p[1]) + T[9] * p[2]) + T[13];
  o[2] = ((T[2] * p[0] + T[6] * p[1]) + T[10] * p[2]) + T[14];
}
// 74 sums of one evaluation pass at state x: static part + residual part (d from the float transform, like PCL)
void evaluate(const float* moved, const float* fixed, const double* M6, int m, const double x[6], double* sums) {
  float T[16];
  s3d::matrix_from_state(x, T);
  for (int i = 0; i < s3d::kNumMoments; ++i) sums[i] = 0.0;
  for (int i = 0; i < m; ++i) {
    const float* p = moved + 4 * i; const float* q = fixed + 4 * i; const double* M = M6 + 6 * i;
    float pp[3];
    transform_mv(T, p, pp);
    const double d[3] = {(double)(pp[0] - q[0]), (double)(pp[1] - q[1]), (double)(pp[2] - q[2])};
    const double Mf[3][3] = {{M[0], M[1], M[2]}, {M[1], M[3], M[4]}, {M[2], M[4], M[5]}};
    const double phi[4] = {p[0], p[1], p[2], 1.0};
    double Md[3];
    for (int a = 0; a < 3; ++a) Md[a] = Mf[a][0] * d[0] + Mf[a][1] * d[1] + Mf[a][2] * d[2];
    for (int a = 0; a < 3; ++a) for (int b = a; b < 3; ++b) for (int c = 0; c < 4; ++c) for (int e = c; e < 4; ++e)
      sums[s3d::sym3(a, b) * 10 + s3d::sym4(c, e)] += Mf[a][b] * phi[c] * phi[e];
    for (int a = 0; a < 3; ++a) for (int c = 0; c < 4; ++c) sums[60 + a * 4 + c] += Md[a] * phi[c];
    sums[72] += d[0] * Md[0] + d[1] * Md[1] + d[2] * Md[2];
    sums[73] += 1.0;
  }
}
}  // namespace

extern "C" {
void hm_objective(const float* moved, const float* fixed, const double* M6, int m, const double* x, double* f, double* g, double* H) {
  double sums[s3d::kNumMoments], gg[6], HH[6][6];
  evaluate(moved, fixed, M6, m, x, sums);
  s3d::objective_from_sums(sums, x, *f, gg, HH);
  for (int i = 0; i < 6; ++i) { g[i] = gg[i]; for (int j = 0; j < 6; ++j) H[j * 6 + i] = HH[i][j]; }
}
// estimateRigidTransformationNewton through the resumable state machine with the speculative back-tracking of
// gicp_ctrl_kernel (trial 0 alone, then trials 1..9 in one go); T column-major float in/out
int hm_newton(const float* moved, const float* fixed, const double* M6, int m, float* T, int max_inner, int* inner_done, int* evaluations) {
  if (m < 4) return 2;
  s3d::NewtonState st;
  s3d::newton_begin(st, T);
  double sums[s3d::kNumMoments], trial_sums[s3d::kLineSearchTrials][s3d::kNumMoments];
  int evals = 1;
  evaluate(moved, fixed, M6, m, st.xc, sums);
  bool more = s3d::newton_advance(st, sums, max_inner);  // objective at x0, first step
  while (more) {
    int first = 0, count = 1, j = -1;
    for (;;) {
      double f_trial[s3d::kLineSearchTrials];
      for (int t = first; t < first + count; ++t) {
        double xc[6];
        s3d::newton_trial_state(st, t, xc);
        evaluate(moved, fixed, M6, m, xc, trial_sums[t]);
        f_trial[t] = trial_sums[t][72] / trial_sums[t][73];
      }
      ++evals;
      j = s3d::newton_pick_trial(st, f_trial, first, count);
      if (j >= 0 || first != 0) break;
      first = 1; count = s3d::kLineSearchTrials - 1;
    }
    if (j < 0) { st.phase = 2; break; }  // no improvement
    s3d::newton_select_trial(st, j);
    more = s3d::newton_advance(st, trial_sums[j], max_inner);
  }
  s3d::matrix_from_state(st.x, T);
  *inner_done = st.it;
  if (evaluations) *evaluations = evals;
  return 0;
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
}
