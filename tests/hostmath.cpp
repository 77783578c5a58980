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

// ---- NDT: host run of slam3d_b200/csrc/ndt_math.h --------------------------------------------------------------------
// Does on the host what ndt.cu does on the device (voxel keys, stable order inside a voxel, leaf finalisation, 27-voxel
// probe, candidates ordered by (d2, key), the 44 sums, the resumable optimiser), with plain sequential sums.
#include <algorithm>
#include <cmath>
#include <unordered_map>

#include "../slam3d_b200/csrc/ndt_math.h"

namespace {
inline void transform_se3(const float* T, const float* p, float* o) {  // x*c0 + (y*c1 + (z*c2 + c3))
  o[0] = p[0] * T[0] + (p[1] * T[4] + (p[2] * T[8] + T[12]));
  o[1] = p[0] * T[1] + (p[1] * T[5] + (p[2] * T[9] + T[13]));
  o[2] = p[0] * T[2] + (p[1] * T[6] + (p[2] * T[10] + T[14]));
}
struct HostNdtGrid {
  std::vector<s3d::NdtLeaf> leaves;
  std::unordered_map<uint32_t, uint32_t> index;  // voxel key -> leaf
  float inv; int min_b[3], div_b[3]; uint32_t mul1, mul2;
};
bool build_grid(const float* pts, int n, float resolution, HostNdtGrid& G) {
  G.inv = 1.0f / resolution;
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int i = 0; i < n; ++i) for (int a = 0; a < 3; ++a) { mn[a] = std::min(mn[a], pts[4 * i + a]); mx[a] = std::max(mx[a], pts[4 * i + a]); }
  const long long dx = (long long)((mx[0] - mn[0]) * G.inv) + 1, dy = (long long)((mx[1] - mn[1]) * G.inv) + 1, dz = (long long)((mx[2] - mn[2]) * G.inv) + 1;
  if (dx * dy * dz > 2147483647ll) return false;
  for (int a = 0; a < 3; ++a) { G.min_b[a] = (int)std::floor(mn[a] * G.inv); G.div_b[a] = (int)std::floor(mx[a] * G.inv) - G.min_b[a] + 1; }
  G.mul1 = (uint32_t)G.div_b[0]; G.mul2 = (uint32_t)G.div_b[0] * (uint32_t)G.div_b[1];
  std::vector<std::pair<uint32_t, uint32_t>> kv(n);
  for (int i = 0; i < n; ++i) {
    const int i0 = (int)std::floor(pts[4 * i] * G.inv) - G.min_b[0], i1 = (int)std::floor(pts[4 * i + 1] * G.inv) - G.min_b[1], i2 = (int)std::floor(pts[4 * i + 2] * G.inv) - G.min_b[2];
    kv[i] = {(uint32_t)i0 + (uint32_t)i1 * G.mul1 + (uint32_t)i2 * G.mul2, (uint32_t)i};
  }
  std::stable_sort(kv.begin(), kv.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
  size_t i = 0;
  while (i < kv.size()) {
    size_t j = i + 1;
    while (j < kv.size() && kv[j].first == kv[i].first) ++j;
    const int nr = (int)(j - i);
    if (nr >= 6) {
      float cs[3] = {0, 0, 0}; double ms[3] = {0, 0, 0}, cov[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
      for (size_t l = i; l < j; ++l) {
        const float* p = pts + 4 * kv[l].second;
        const double d[3] = {p[0], p[1], p[2]};
        for (int a = 0; a < 3; ++a) { ms[a] += d[a]; cs[a] += p[a]; for (int b = 0; b < 3; ++b) cov[a][b] += d[a] * d[b]; }
      }
      s3d::NdtLeaf L;
      L.cx = cs[0] / (float)nr; L.cy = cs[1] / (float)nr; L.cz = cs[2] / (float)nr; L.key = kv[i].first;
      s3d::ndt_finalize_leaf(nr, ms, cov, L.mean, L.icov);
      G.index[L.key] = (uint32_t)G.leaves.size();
      G.leaves.push_back(L);
    }
    i = j;
  }
  return true;
}
void ndt_evaluate(const HostNdtGrid& G, const float* src, int n, const float* T, const double x_eval[6], double d1, double d2, float r2, double* sums) {
  for (int i = 0; i < s3d::kNdtSums; ++i) sums[i] = 0.0;
  s3d::NdtAngular A;
  s3d::ndt_angle_derivatives(x_eval, A);
  for (int i = 0; i < n; ++i) {
    float q[3];
    transform_se3(T, src + 4 * i, q);
    const int c[3] = {(int)std::floor(q[0] * G.inv) - G.min_b[0], (int)std::floor(q[1] * G.inv) - G.min_b[1], (int)std::floor(q[2] * G.inv) - G.min_b[2]};
    struct Cand { float d; uint32_t key, leaf; };
    Cand cand[125]; int nc = 0;
    for (int dz = -2; dz <= 2; ++dz) for (int dy = -2; dy <= 2; ++dy) for (int dx = -2; dx <= 2; ++dx) {
      const int ix = c[0] + dx, iy = c[1] + dy, iz = c[2] + dz;
      if (ix < 0 || iy < 0 || iz < 0 || ix >= G.div_b[0] || iy >= G.div_b[1] || iz >= G.div_b[2]) continue;
      auto it = G.index.find((uint32_t)ix + (uint32_t)iy * G.mul1 + (uint32_t)iz * G.mul2);
      if (it == G.index.end()) continue;
      const s3d::NdtLeaf& L = G.leaves[it->second];
      const float ex = q[0] - L.cx, ey = q[1] - L.cy, ez = q[2] - L.cz;
      float d = ex * ex; d = d + ey * ey; d = d + ez * ez;
      if (d < r2) cand[nc++] = {d, L.key, it->second};
    }
    std::sort(cand, cand + nc, [](const Cand& a, const Cand& b) { return a.d < b.d || (a.d == b.d && a.key < b.key); });
    if (!nc) continue;
    const double xo[3] = {src[4 * i], src[4 * i + 1], src[4 * i + 2]};
    s3d::NdtPointDerivs P;
    s3d::ndt_point_derivatives(A, xo, P);
    for (int k = 0; k < nc; ++k) {
      const s3d::NdtLeaf& L = G.leaves[cand[k].leaf];
      const double xt[3] = {(double)q[0] - L.mean[0], (double)q[1] - L.mean[1], (double)q[2] - L.mean[2]};
      s3d::ndt_accumulate(P, d1, d2, xt, L.icov, sums);
      sums[43] += 1.0;
    }
  }
}
}  // namespace

extern "C" {
// pcl_source (moving, n_src x 4 floats), pcl_target (voxelised, n_tgt x 4 floats), guess column-major float 4x4.
// out: T[16] final_transformation_, info = {converged, nr_iterations, line_iterations, n_pairs_last, evaluations, n_leaves}
int hm_ndt_register(const float* src, int n_src, const float* tgt, int n_tgt, const float* guess, float resolution, double step_size,
                    double outlier_ratio, double trans_eps, int max_iter, float* T_out, int* info) {
  HostNdtGrid G;
  for (int i = 0; i < 16; ++i) T_out[i] = (i % 5 == 0) ? 1.f : 0.f;
  for (int i = 0; i < 6; ++i) info[i] = 0;
  if (!build_grid(tgt, n_tgt, resolution, G) || G.leaves.empty()) return 0;
  info[5] = (int)G.leaves.size();
  double d1, d2;
  s3d::ndt_gauss_constants((double)resolution, outlier_ratio, d1, d2);
  const float r2 = (float)((double)resolution * (double)resolution);
  float T[16];
  for (int i = 0; i < 16; ++i) T[i] = guess[i];
  double x0[6];
  s3d::ndt_initial_state(T, x0);
  s3d::NdtOptState st;
  s3d::ndt_opt_begin(st, x0, step_size, trans_eps, max_iter);
  double sums[s3d::kNdtSums];
  int evals = 0;
  for (;;) {
    ndt_evaluate(G, src, n_src, T, st.x_t, d1, d2, r2, sums);
    ++evals;
    if (!s3d::ndt_opt_on_eval(st, sums)) break;
    s3d::ndt_convert_transform(st.x_t, T);
  }
  for (int i = 0; i < 16; ++i) T_out[i] = T[i];
  info[0] = st.converged; info[1] = st.nr_iterations; info[2] = st.line_iterations; info[3] = (int)st.n_pairs_last; info[4] = evals;
  return 0;
}
void hm_ndt_svd_solve(const double* H, const double* b, double* x) {
  double HH[6][6];
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) HH[i][j] = H[i * 6 + j];
  s3d::ndt_svd_solve6(HH, b, x);
}
}
