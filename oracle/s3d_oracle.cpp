// s3d_oracle.cpp — CPU ORACLE for the slam3d PointCloudSensor scan-matching path.
//
// *** TEST INFRASTRUCTURE, NOT PRODUCT. ***  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library.  The product (slam3d_b200/) never does.
//
// *** PARITY UNPINNED. ***  The arithmetic of this path lives in PCL (+FLANN, Eigen), an external,
// un-vendored, un-pinned dependency of the reference (slam3d-dependencies.cmake:24, floor 1.8.1; the
// reference CI installs ubuntu:latest's libpcl-dev = PCL 1.14).  None of it exists in this environment and
// the reference holds no test or golden vector for downsample/align (PointCloudSensorTest.cpp:29-96 covers
// serialization and an empty cloud only).  This file therefore RESTATES, from the published algorithms,
//   slam3d glue   slam3d/sensor/pcl/PointCloudSensor.cpp:52-82 (doICP), :119-174 (align), :190-201 (downsample)
//                 slam3d/sensor/pcl/RegistrationParameters.hpp:36-97
//   PCL 1.14      pcl/filters/impl/voxel_grid.hpp  VoxelGrid<PointXYZ>::applyFilter          
//                 pcl/kdtree + FLANN KDTreeSingleIndex / L2_Simple<float>, exact search        
//                 pcl/registration/impl/gicp.hpp   computeCovariances / computeTransformation /
//                                                  estimateRigidTransformationNewton /
//                                                  OptimizationFunctorWithIndices::operator(), dfddf,
//                                                  applyState                                   
//                 pcl/registration/impl/registration.hpp  align / getFitnessScore(max_range)   
//                 pcl/common/impl/transforms.hpp   transformPointCloud (float, se3 form)
// and is pinned only against independent re-derivations (numpy voxel keys, brute force / scipy cKDTree
// neighbours, numpy eigh normals, finite-difference derivatives) in tests/test_oracle_*.py.
//
// Deliberate, documented definitions where PCL's behaviour is unspecified:
//   * VoxelGrid sorts (voxel, point) pairs with an unstable sort; here the sort is stable, so each centroid
//     sums its points in ascending input order (cannot change the leaf assignment).
//   * FLANN breaks exact float distance ties by traversal order; here ties go to the LOWEST index.
//   * Eigen's JacobiSVD / SelfAdjointEigenSolver are replaced by cyclic Jacobi eigen-solvers (same
//     decomposition; differs only in rounding, and arbitrarily for exactly degenerate neighbourhoods).
//   * applyState builds the rotation in float through Eigen AngleAxis/quaternion products; here
//     R = Rz*Ry*Rx is formed in double and rounded to float.
//
// Build: see oracle/Makefile (g++ -O2 -ffp-contract=off: float expressions must not be fused).

#include "../include/s3d_b200.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include <atomic>
#include <thread>

namespace {

struct P4 { float x, y, z, w; };

thread_local std::string g_err;

// Dynamic-chunk parallel loop over std::thread (n_threads <= 0: all hardware threads).
template <typename F>
static void parallel_for(int64_t n, int n_threads, int64_t chunk, F&& body) {
  unsigned hw = std::thread::hardware_concurrency();
  int nt = n_threads > 0 ? n_threads : (hw ? static_cast<int>(hw) : 1);
  nt = static_cast<int>(std::min<int64_t>(nt, (n + chunk - 1) / chunk));
  if (nt <= 1) { for (int64_t i = 0; i < n; ++i) body(i); return; }
  std::atomic<int64_t> next{0};
  std::vector<std::thread> pool;
  for (int t = 0; t < nt; ++t)
    pool.emplace_back([&]() {
      for (;;) {
        const int64_t b = next.fetch_add(chunk);
        if (b >= n) break;
        for (int64_t i = b; i < std::min(n, b + chunk); ++i) body(i);
      }
    });
  for (auto& th : pool) th.join();
}

// ------------------------------------------------------------------------------------------------
// Float primitives with PCL's operation order (compiled with -ffp-contract=off, SSE scalar floats).
// ------------------------------------------------------------------------------------------------

// flann::L2_Simple<float>: result = 0; for d: diff = a[d]-b[d]; result += diff*diff.   
static inline float dist2(const P4& a, const P4& b) {
  float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
  float r = dx * dx;
  r = r + dy * dy;
  r = r + dz * dz;
  return r;
}

// Column-major 4x4 float matrix.
struct M4f { float m[16]; float operator()(int r, int c) const { return m[c * 4 + r]; } float& operator()(int r, int c) { return m[c * 4 + r]; } };

static M4f m4f_identity() { M4f a; for (int i = 0; i < 16; ++i) a.m[i] = (i % 5 == 0) ? 1.f : 0.f; return a; }

// pcl::transformPointCloud, float se3 form: x*c0 + (y*c1 + (z*c2 + c3)).
static inline P4 transform_se3(const M4f& T, const P4& p) {
  P4 o;
  o.x = p.x * T(0, 0) + (p.y * T(0, 1) + (p.z * T(0, 2) + T(0, 3)));
  o.y = p.x * T(1, 0) + (p.y * T(1, 1) + (p.z * T(1, 2) + T(1, 3)));
  o.z = p.x * T(2, 0) + (p.y * T(2, 1) + (p.z * T(2, 2) + T(2, 3)));
  o.w = 1.f;
  return o;
}

// Eigen Matrix4f * Vector4f with w == 1: ((c0*x + c1*y) + c2*z) + c3.   
static inline P4 transform_mv(const M4f& T, const P4& p) {
  P4 o;
  o.x = ((T(0, 0) * p.x + T(0, 1) * p.y) + T(0, 2) * p.z) + T(0, 3);
  o.y = ((T(1, 0) * p.x + T(1, 1) * p.y) + T(1, 2) * p.z) + T(1, 3);
  o.z = ((T(2, 0) * p.x + T(2, 1) * p.y) + T(2, 2) * p.z) + T(2, 3);
  o.w = 1.f;
  return o;
}

// Eigen Matrix4f * Matrix4f: entry = ((a0*b0 + a1*b1) + a2*b2) + a3*b3.
static M4f m4f_mul(const M4f& A, const M4f& B) {
  M4f C;
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c)
      C(r, c) = ((A(r, 0) * B(0, c) + A(r, 1) * B(1, c)) + A(r, 2) * B(2, c)) + A(r, 3) * B(3, c);
  return C;
}

// ------------------------------------------------------------------------------------------------
// VoxelGrid<PointXYZ>::applyFilter   (called from PointCloudSensor.cpp:195-198)
// ------------------------------------------------------------------------------------------------
static inline bool finite3(const P4& p) { return std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z); }

static void voxel_filter(const P4* in, size_t n, float leaf, std::vector<P4>& out, uint32_t* leaf_index, int* overflow) {
  out.clear();
  if (overflow) *overflow = 0;
  if (leaf_index) std::fill(leaf_index, leaf_index + n, 0xFFFFFFFFu);
  if (n == 0) return;  // PointCloudSensor.cpp:193  (in->size() > 0)
  const float inv = 1.0f / leaf;  // setLeafSize: inverse_leaf_size_ = 1 / leaf_size_ (float)
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  size_t nfinite = 0;
  for (size_t i = 0; i < n; ++i) {  // getMinMax3D (non-finite points skipped, A.1 step 2)
    if (!finite3(in[i])) continue;
    ++nfinite;
    mn[0] = std::min(mn[0], in[i].x); mx[0] = std::max(mx[0], in[i].x);
    mn[1] = std::min(mn[1], in[i].y); mx[1] = std::max(mx[1], in[i].y);
    mn[2] = std::min(mn[2], in[i].z); mx[2] = std::max(mx[2], in[i].z);
  }
  if (nfinite == 0) return;
  // A.1 step 3: overflow guard, float arithmetic then truncation to int64.
  int64_t dx = static_cast<int64_t>((mx[0] - mn[0]) * inv) + 1;
  int64_t dy = static_cast<int64_t>((mx[1] - mn[1]) * inv) + 1;
  int64_t dz = static_cast<int64_t>((mx[2] - mn[2]) * inv) + 1;
  if (dx * dy * dz > static_cast<int64_t>(std::numeric_limits<int32_t>::max())) {
    out.assign(in, in + n);  // "output = *input_"
    if (overflow) *overflow = 1;
    return;
  }
  int min_b[3], max_b[3], div_b[3];
  for (int a = 0; a < 3; ++a) {  // A.1 step 4
    min_b[a] = static_cast<int>(std::floor(mn[a] * inv));
    max_b[a] = static_cast<int>(std::floor(mx[a] * inv));
    div_b[a] = max_b[a] - min_b[a] + 1;
  }
  const int mul1 = div_b[0];
  const int mul2 = static_cast<int>(static_cast<uint32_t>(div_b[0]) * static_cast<uint32_t>(div_b[1]));
  struct KI { uint32_t key; uint32_t idx; };
  std::vector<KI> kv;
  kv.reserve(nfinite);
  for (size_t i = 0; i < n; ++i) {  // A.1 step 5
    if (!finite3(in[i])) continue;
    int ijk0 = static_cast<int>(std::floor(in[i].x * inv) - static_cast<float>(min_b[0]));
    int ijk1 = static_cast<int>(std::floor(in[i].y * inv) - static_cast<float>(min_b[1]));
    int ijk2 = static_cast<int>(std::floor(in[i].z * inv) - static_cast<float>(min_b[2]));
    uint32_t key = static_cast<uint32_t>(ijk0) + static_cast<uint32_t>(ijk1) * static_cast<uint32_t>(mul1) +
                   static_cast<uint32_t>(ijk2) * static_cast<uint32_t>(mul2);
    kv.push_back({key, static_cast<uint32_t>(i)});
    if (leaf_index) leaf_index[i] = key;
  }
  // A.1 step 6 (stable here: ascending input index inside a voxel).
  std::stable_sort(kv.begin(), kv.end(), [](const KI& a, const KI& b) { return a.key < b.key; });
  // A.1 steps 7-9: groups, float centroid (CentroidPoint<PointXYZ>: float sum, / float(count)).
  size_t i = 0;
  while (i < kv.size()) {
    size_t j = i + 1;
    while (j < kv.size() && kv[j].key == kv[i].key) ++j;
    float sx = 0.f, sy = 0.f, sz = 0.f;
    for (size_t l = i; l < j; ++l) { sx += in[kv[l].idx].x; sy += in[kv[l].idx].y; sz += in[kv[l].idx].z; }
    const float cnt = static_cast<float>(j - i);
    out.push_back({sx / cnt, sy / cnt, sz / cnt, 1.0f});
    i = j;
  }
}

// ------------------------------------------------------------------------------------------------
// Exact nearest-neighbour search  (SURVEY A.2).  A bounding-box kd-tree whose pruning bound is computed
// with the same monotone float operations as dist2(), so it never prunes a point that brute force with
// dist2() and lexicographic (d2, index) order would have returned.
// ------------------------------------------------------------------------------------------------
struct KdTree {
  struct Node { float lo[3], hi[3]; int left, right; int begin, end; };
  std::vector<Node> nodes;
  std::vector<P4> pts;        // reordered copy
  std::vector<uint32_t> idx;  // reordered -> original index
  static constexpr int kLeaf = 15;  // flann::KDTreeSingleIndexParams(15)

  void build(const P4* p, size_t n) {
    pts.assign(p, p + n);
    idx.resize(n);
    for (size_t i = 0; i < n; ++i) idx[i] = static_cast<uint32_t>(i);
    nodes.clear();
    nodes.reserve(n / 4 + 16);
    if (n) build_rec(0, static_cast<int>(n));
  }
  int build_rec(int b, int e) {
    Node nd;
    for (int a = 0; a < 3; ++a) { nd.lo[a] = FLT_MAX; nd.hi[a] = -FLT_MAX; }
    for (int i = b; i < e; ++i) {
      const float v[3] = {pts[i].x, pts[i].y, pts[i].z};
      for (int a = 0; a < 3; ++a) { nd.lo[a] = std::min(nd.lo[a], v[a]); nd.hi[a] = std::max(nd.hi[a], v[a]); }
    }
    nd.left = nd.right = -1; nd.begin = b; nd.end = e;
    const int me = static_cast<int>(nodes.size());
    nodes.push_back(nd);
    if (e - b > kLeaf) {
      int ax = 0;
      float ext = nd.hi[0] - nd.lo[0];
      for (int a = 1; a < 3; ++a) if (nd.hi[a] - nd.lo[a] > ext) { ext = nd.hi[a] - nd.lo[a]; ax = a; }
      const int mid = (b + e) / 2;
      // median split on the widest axis; ties by original index keep the build deterministic
      std::vector<int> ord(e - b);
      for (int i = b; i < e; ++i) ord[i - b] = i;
      auto coord = [&](int i) { return ax == 0 ? pts[i].x : (ax == 1 ? pts[i].y : pts[i].z); };
      std::nth_element(ord.begin(), ord.begin() + (mid - b), ord.end(), [&](int a, int c) {
        float va = coord(a), vc = coord(c);
        return va < vc || (va == vc && idx[a] < idx[c]);
      });
      std::vector<P4> tp(e - b);
      std::vector<uint32_t> ti(e - b);
      for (int i = 0; i < e - b; ++i) { tp[i] = pts[ord[i]]; ti[i] = idx[ord[i]]; }
      std::copy(tp.begin(), tp.end(), pts.begin() + b);
      std::copy(ti.begin(), ti.end(), idx.begin() + b);
      int l = build_rec(b, mid);
      int r = build_rec(mid, e);
      nodes[me].left = l; nodes[me].right = r;
    }
    return me;
  }
  // Lower bound of dist2(q, p) for every p inside the node's box, in the float arithmetic of dist2().
  static inline float box_bound(const Node& nd, const P4& q) {
    const float v[3] = {q.x, q.y, q.z};
    float e[3];
    for (int a = 0; a < 3; ++a) e[a] = v[a] < nd.lo[a] ? nd.lo[a] - v[a] : (v[a] > nd.hi[a] ? v[a] - nd.hi[a] : 0.f);
    float r = e[0] * e[0];
    r = r + e[1] * e[1];
    r = r + e[2] * e[2];
    return r;
  }
  struct Cand { float d2; uint32_t id; };
  static inline bool less(const Cand& a, const Cand& b) { return a.d2 < b.d2 || (a.d2 == b.d2 && a.id < b.id); }

  // k best by (d2, original index), ascending.  Returns number found (min(k, n)).
  int knn(const P4& q, int k, Cand* best) const {
    int found = 0;
    if (nodes.empty() || k <= 0) return 0;
    search(0, q, k, best, found);
    return found;
  }
  void search(int ni, const P4& q, int k, Cand* best, int& found) const {
    const Node& nd = nodes[ni];
    if (nd.left < 0) {
      for (int i = nd.begin; i < nd.end; ++i) {
        Cand c{dist2(q, pts[i]), idx[i]};
        if (found < k) {
          int j = found++;
          while (j > 0 && less(c, best[j - 1])) { best[j] = best[j - 1]; --j; }
          best[j] = c;
        } else if (less(c, best[k - 1])) {
          int j = k - 1;
          while (j > 0 && less(c, best[j - 1])) { best[j] = best[j - 1]; --j; }
          best[j] = c;
        }
      }
      return;
    }
    const float bl = box_bound(nodes[nd.left], q), br = box_bound(nodes[nd.right], q);
    const int first = bl <= br ? nd.left : nd.right, second = bl <= br ? nd.right : nd.left;
    const float bf = std::min(bl, br), bs = std::max(bl, br);
    if (found < k || !(bf > best[k - 1].d2)) search(first, q, k, best, found);   // bound == worst is still visited
    if (found < k || !(bs > best[k - 1].d2)) search(second, q, k, best, found);  // (a lower index may tie)
  }
};

// ------------------------------------------------------------------------------------------------
// Small dense linear algebra (double).
// ------------------------------------------------------------------------------------------------
struct M3 { double a[3][3]; };  // a[row][col]

static M3 m3_zero() { M3 m; std::memset(&m, 0, sizeof m); return m; }
static M3 m3_mul(const M3& A, const M3& B) {
  M3 C = m3_zero();
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double s = 0; for (int k = 0; k < 3; ++k) s += A.a[i][k] * B.a[k][j]; C.a[i][j] = s; }
  return C;
}
static M3 m3_transpose(const M3& A) { M3 C; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) C.a[i][j] = A.a[j][i]; return C; }
// Eigen 3x3 inverse: cofactors / determinant.
static M3 m3_inverse(const M3& A) {
  const double(*a)[3] = A.a;
  M3 c;
  c.a[0][0] = a[1][1] * a[2][2] - a[1][2] * a[2][1];
  c.a[0][1] = a[0][2] * a[2][1] - a[0][1] * a[2][2];
  c.a[0][2] = a[0][1] * a[1][2] - a[0][2] * a[1][1];
  c.a[1][0] = a[1][2] * a[2][0] - a[1][0] * a[2][2];
  c.a[1][1] = a[0][0] * a[2][2] - a[0][2] * a[2][0];
  c.a[1][2] = a[0][2] * a[1][0] - a[0][0] * a[1][2];
  c.a[2][0] = a[1][0] * a[2][1] - a[1][1] * a[2][0];
  c.a[2][1] = a[0][1] * a[2][0] - a[0][0] * a[2][1];
  c.a[2][2] = a[0][0] * a[1][1] - a[0][1] * a[1][0];
  const double det = a[0][0] * c.a[0][0] + a[0][1] * c.a[1][0] + a[0][2] * c.a[2][0];
  const double inv = 1.0 / det;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) c.a[i][j] *= inv;
  return c;
}

// Cyclic Jacobi for a symmetric NxN matrix: A = V diag(w) V^T.  A is destroyed.
template <int N>
static void jacobi_eigen(double A[N][N], double V[N][N], double w[N]) {
  for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = 0, diag = 0;
    for (int i = 0; i < N; ++i) { diag += A[i][i] * A[i][i]; for (int j = i + 1; j < N; ++j) off += A[i][j] * A[i][j]; }
    if (off <= 1e-34 * diag || off == 0.0) break;
    for (int p = 0; p < N - 1; ++p)
      for (int q = p + 1; q < N; ++q) {
        const double apq = A[p][q];
        if (apq == 0.0) continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < N; ++k) {  // A <- A J
          const double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq;
          A[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < N; ++k) {  // A <- J^T A
          const double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk;
          A[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < N; ++k) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
  }
  for (int i = 0; i < N; ++i) w[i] = A[i][i];
}

// ------------------------------------------------------------------------------------------------
// GICP computeCovariances   
// ------------------------------------------------------------------------------------------------
static const double kGicpEpsilon = 1e-3;  // gicp_epsilon_, not exposed by slam3d

static void covariance_from_neighbours(const P4* cloud, const KdTree::Cand* nb, int k, M3& C) {
  double mean[3] = {0, 0, 0};
  double cov[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int j = 0; j < k; ++j) {
    const P4& pt = cloud[nb[j].id];
    mean[0] += pt.x; mean[1] += pt.y; mean[2] += pt.z;
    cov[0][0] += pt.x * pt.x;  // float product, double accumulate (A.3 step 2)
    cov[1][0] += pt.y * pt.x;
    cov[1][1] += pt.y * pt.y;
    cov[2][0] += pt.z * pt.x;
    cov[2][1] += pt.z * pt.y;
    cov[2][2] += pt.z * pt.z;
  }
  for (int a = 0; a < 3; ++a) mean[a] /= static_cast<double>(k);
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b <= a; ++b) {
      cov[a][b] /= static_cast<double>(k);
      cov[a][b] -= mean[a] * mean[b];
      cov[b][a] = cov[a][b];
    }
  double V[3][3], w[3];
  jacobi_eigen<3>(cov, V, w);
  // JacobiSVD orders singular values (= |eigenvalues|) descending; the last one gets gicp_epsilon.
  int order[3] = {0, 1, 2};
  std::stable_sort(order, order + 3, [&](int a, int b) { return std::fabs(w[a]) > std::fabs(w[b]); });
  C = m3_zero();
  for (int kk = 0; kk < 3; ++kk) {
    const int c = order[kk];
    const double v = (kk == 2) ? kGicpEpsilon : 1.0;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) C.a[i][j] += v * V[i][c] * V[j][c];
  }
}

static bool compute_covariances(const std::vector<P4>& cloud, const KdTree& tree, int k, std::vector<M3>& covs,
                                uint32_t* knn_index, float* knn_dist2) {
  if (k > static_cast<int>(cloud.size())) return false;  // PCL: error, covariances stay empty
  covs.resize(cloud.size());
  std::vector<KdTree::Cand> nb(k);
  for (size_t i = 0; i < cloud.size(); ++i) {
    tree.knn(cloud[i], k, nb.data());
    if (knn_index) for (int j = 0; j < k; ++j) knn_index[i * k + j] = nb[j].id;
    if (knn_dist2) for (int j = 0; j < k; ++j) knn_dist2[i * k + j] = nb[j].d2;
    covariance_from_neighbours(cloud.data(), nb.data(), k, covs[i]);
  }
  return true;
}

// ------------------------------------------------------------------------------------------------
// Euler ZYX rotation and its derivatives   (gicp.hpp applyState / computeRDerivative / dfddf)
// x = (tx,ty,tz, phi(x-axis), theta(y-axis), psi(z-axis));  R = Rz(psi) Ry(theta) Rx(phi).
// ------------------------------------------------------------------------------------------------
struct EulerDerivs { M3 R; M3 dR[3]; M3 ddR[3][3]; };

static void euler_factors(double ang, int axis, M3& R, M3& dR, M3& ddR) {
  const double c = std::cos(ang), s = std::sin(ang);
  R = m3_zero(); dR = m3_zero(); ddR = m3_zero();
  const int i = (axis + 1) % 3, j = (axis + 2) % 3;  // rotation in the (i,j) plane
  R.a[axis][axis] = 1.0;
  R.a[i][i] = c;  R.a[i][j] = -s; R.a[j][i] = s;  R.a[j][j] = c;
  dR.a[i][i] = -s; dR.a[i][j] = -c; dR.a[j][i] = c;  dR.a[j][j] = -s;
  ddR.a[i][i] = -c; ddR.a[i][j] = s; ddR.a[j][i] = -s; ddR.a[j][j] = -c;
}

static void euler_derivs(const double x[6], EulerDerivs& E, bool second) {
  M3 X, dX, ddX, Y, dY, ddY, Z, dZ, ddZ;
  euler_factors(x[3], 0, X, dX, ddX);
  euler_factors(x[4], 1, Y, dY, ddY);
  euler_factors(x[5], 2, Z, dZ, ddZ);
  const M3 ZY = m3_mul(Z, Y);
  E.R = m3_mul(ZY, X);
  E.dR[0] = m3_mul(ZY, dX);
  E.dR[1] = m3_mul(m3_mul(Z, dY), X);
  E.dR[2] = m3_mul(m3_mul(dZ, Y), X);
  if (!second) return;
  E.ddR[0][0] = m3_mul(ZY, ddX);
  E.ddR[1][1] = m3_mul(m3_mul(Z, ddY), X);
  E.ddR[2][2] = m3_mul(m3_mul(ddZ, Y), X);
  E.ddR[0][1] = E.ddR[1][0] = m3_mul(m3_mul(Z, dY), dX);
  E.ddR[0][2] = E.ddR[2][0] = m3_mul(m3_mul(dZ, Y), dX);
  E.ddR[1][2] = E.ddR[2][1] = m3_mul(m3_mul(dZ, dY), X);
}

// applyState on base_transformation_ = I:  T = [R(x) | t] rounded to float.
static M4f state_to_matrix(const double x[6]) {
  EulerDerivs E;
  euler_derivs(x, E, false);
  M4f T = m4f_identity();
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) T(r, c) = static_cast<float>(E.R.a[r][c]);
  for (int r = 0; r < 3; ++r) T(r, 3) = static_cast<float>(x[r]);
  return T;
}

// ------------------------------------------------------------------------------------------------
// GICP registration   (SURVEY A.4 / A.5 — PCL 1.14 Newton inner optimiser)
// ------------------------------------------------------------------------------------------------
struct GicpProblem {
  const std::vector<P4>* moved;   // `output`: PCL source (= slam3d target) after guess, float
  const std::vector<P4>* fixed;   // PCL target (= slam3d source)
  std::vector<int> src_idx, tgt_idx;
  std::vector<M3> mahalanobis;    // indexed by source point
};

// OptimizationFunctorWithIndices::operator()
static double gicp_f(const GicpProblem& P, const double x[6]) {
  const M4f T = state_to_matrix(x);
  double f = 0;
  const int m = static_cast<int>(P.src_idx.size());
  for (int i = 0; i < m; ++i) {
    const P4& ps = (*P.moved)[P.src_idx[i]];
    const P4& pt = (*P.fixed)[P.tgt_idx[i]];
    const P4 pp = transform_mv(T, ps);
    const double d[3] = {static_cast<double>(pp.x - pt.x), static_cast<double>(pp.y - pt.y), static_cast<double>(pp.z - pt.z)};
    const M3& M = P.mahalanobis[P.src_idx[i]];
    double Md[3];
    for (int a = 0; a < 3; ++a) Md[a] = M.a[a][0] * d[0] + M.a[a][1] * d[1] + M.a[a][2] * d[2];
    f += d[0] * Md[0] + d[1] * Md[1] + d[2] * Md[2];
  }
  return f / m;
}

// OptimizationFunctorWithIndices::dfddf: exact gradient and Hessian of f w.r.t. (t, phi, theta, psi).
static void gicp_dfddf(const GicpProblem& P, const double x[6], double g[6], double H[6][6]) {
  const M4f T = state_to_matrix(x);
  const int m = static_cast<int>(P.src_idx.size());
  double gt[3] = {0, 0, 0};
  double Htt[3][3] = {{0}}, dC[3][3] = {{0}};     // dC[b][a] = sum p_b (Md)_a
  double Tb[3][3][3] = {{{0}}};                   // Tb[b][a][c] = sum p_b M[a][c]
  double Hr[3][3][3][3] = {{{{0}}}};              // Hr[a][c][b][e] = sum M[a][c] p_b p_e
  for (int i = 0; i < m; ++i) {
    const P4& ps = (*P.moved)[P.src_idx[i]];
    const P4& pt = (*P.fixed)[P.tgt_idx[i]];
    const P4 pp = transform_mv(T, ps);
    const double d[3] = {static_cast<double>(pp.x - pt.x), static_cast<double>(pp.y - pt.y), static_cast<double>(pp.z - pt.z)};
    const M3& M = P.mahalanobis[P.src_idx[i]];
    const double p[3] = {ps.x, ps.y, ps.z};  // base_transformation_ = I
    double Md[3];
    for (int a = 0; a < 3; ++a) Md[a] = M.a[a][0] * d[0] + M.a[a][1] * d[1] + M.a[a][2] * d[2];
    for (int a = 0; a < 3; ++a) {
      gt[a] += Md[a];
      for (int c = 0; c < 3; ++c) {
        Htt[a][c] += M.a[a][c];
        for (int b = 0; b < 3; ++b) {
          Tb[b][a][c] += p[b] * M.a[a][c];
          for (int e = 0; e < 3; ++e) Hr[a][c][b][e] += M.a[a][c] * (p[b] * p[e]);
        }
      }
      for (int b = 0; b < 3; ++b) dC[b][a] += p[b] * Md[a];
    }
  }
  const double s = 2.0 / m;
  EulerDerivs E;
  euler_derivs(x, E, true);
  for (int a = 0; a < 6; ++a) { g[a] = 0; for (int b = 0; b < 6; ++b) H[a][b] = 0; }
  for (int a = 0; a < 3; ++a) {
    g[a] = s * gt[a];
    for (int c = 0; c < 3; ++c) H[a][c] = s * Htt[a][c];
  }
  for (int k = 0; k < 3; ++k) {
    double gr = 0;
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) gr += E.dR[k].a[a][b] * (s * dC[b][a]);
    g[3 + k] = gr;
    for (int a = 0; a < 3; ++a) {  // translation-rotation block
      double h = 0;
      for (int c = 0; c < 3; ++c) for (int b = 0; b < 3; ++b) h += E.dR[k].a[c][b] * (s * Tb[b][a][c]);
      H[a][3 + k] = H[3 + k][a] = h;
    }
    for (int l = 0; l < 3; ++l) {  // rotation-rotation block
      double h = 0;
      for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) {
        for (int c = 0; c < 3; ++c) for (int e = 0; e < 3; ++e) h += E.dR[k].a[a][b] * E.dR[l].a[c][e] * (s * Hr[a][c][b][e]);
        h += E.ddR[k][l].a[a][b] * (s * dC[b][a]);
      }
      H[3 + k][3 + l] = h;
    }
  }
}

// estimateRigidTransformationNewton.  Returns false when there are fewer than 4 correspondences
// (PCL throws NotEnoughPointsException, which computeTransformation catches and breaks on).
static bool gicp_newton(const GicpProblem& P, M4f& T, int max_inner, int* inner_done) {
  if (P.src_idx.size() < 4) return false;  // min_number_correspondences_
  double x[6];
  x[0] = T(0, 3); x[1] = T(1, 3); x[2] = T(2, 3);
  x[3] = std::atan2(static_cast<double>(T(2, 1)), static_cast<double>(T(2, 2)));
  x[4] = std::asin(std::min(1.0, std::max(-1.0, -static_cast<double>(T(2, 0)))));
  x[5] = std::atan2(static_cast<double>(T(1, 0)), static_cast<double>(T(0, 0)));
  double g[6], H[6][6];
  double fcur = gicp_f(P, x);
  gicp_dfddf(P, x, g, H);
  int it = 0;
  do {
    ++it;
    double A[6][6], V[6][6], w[6];
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) A[i][j] = H[i][j];
    jacobi_eigen<6>(A, V, w);
    double wmax = w[0];
    for (int i = 1; i < 6; ++i) wmax = std::max(wmax, w[i]);
    double delta[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 6; ++i) {  // delta = V diag(1/w') V^T g, negative eigenvalues -> 1/largest
      const double inv = (w[i] < 0) ? 1.0 / wmax : 1.0 / w[i];
      double proj = 0;
      for (int r = 0; r < 6; ++r) proj += V[r][i] * g[r];
      for (int r = 0; r < 6; ++r) delta[r] += V[r][i] * (inv * proj);
    }
    double alpha = 1.0;
    bool improved = false;
    for (int ls = 0; ls < 10; ++ls, alpha /= 2) {  // back-tracking until f decreases
      double xc[6];
      for (int r = 0; r < 6; ++r) xc[r] = x[r] - alpha * delta[r];
      const double fc = gicp_f(P, xc);
      if (fc < fcur) { for (int r = 0; r < 6; ++r) x[r] = xc[r]; fcur = fc; improved = true; break; }
    }
    if (!improved) break;
    gicp_dfddf(P, x, g, H);
    const double gtn = std::sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
    const double grn = std::sqrt(g[3] * g[3] + g[4] * g[4] + g[5] * g[5]);
    if (gtn < 1e-2 && grn < 1e-2) break;  // translation_/rotation_gradient_tolerance_
  } while (it < max_inner);
  if (inner_done) *inner_done += it;
  T = state_to_matrix(x);
  return true;
}

#include "bfgs_oracle.inc"  // PCL <= 1.13's inner optimiser (and useBFGS() later)

// which inner optimiser gicp_register uses: 0 = Newton (PCL >= 1.14 default; what the CUDA path mirrors), 1 = BFGS
static int g_gicp_optimizer = 0;

struct GicpOutput { M4f final_T; bool converged; int outer_iterations; int inner_iterations; uint32_t n_corr; double fitness; };

// Registration::align + GICP::computeTransformation + getFitnessScore, as driven by doICP (:52-82).
// `pcl_source` = slam3d target (moved), `pcl_target` = slam3d source (fixed).
static bool gicp_register(const std::vector<P4>& pcl_source, const std::vector<P4>& pcl_target, const M4f& guess,
                          const s3d_registration_parameters& cfg, GicpOutput& out) {
  KdTree tree_target, tree_source;
  tree_target.build(pcl_target.data(), pcl_target.size());
  tree_source.build(pcl_source.data(), pcl_source.size());
  std::vector<M3> cov_target, cov_source;
  const int k = cfg.correspondence_randomness;
  const bool have_t = compute_covariances(pcl_target, tree_target, k, cov_target, nullptr, nullptr);
  const bool have_s = compute_covariances(pcl_source, tree_source, k, cov_source, nullptr, nullptr);
  if (!have_t || !have_s) { g_err = "correspondence_randomness larger than the cloud"; return false; }

  const size_t N = pcl_source.size();
  std::vector<P4> moved(N);
  for (size_t i = 0; i < N; ++i) moved[i] = transform_se3(guess, pcl_source[i]);  // transformPointCloud(output, output, guess)

  GicpProblem P;
  P.moved = &moved; P.fixed = &pcl_target;
  P.mahalanobis.resize(N);
  M4f T = m4f_identity(), prev = m4f_identity();
  const double dist_thr = cfg.max_correspondence_distance * cfg.max_correspondence_distance;
  out.converged = false; out.outer_iterations = 0; out.inner_iterations = 0; out.n_corr = 0;
  while (!out.converged) {
    // R = top-left 3x3 of double(transformation_) * double(guess)
    M3 R;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int kk = 0; kk < 4; ++kk) s += static_cast<double>(T(i, kk)) * static_cast<double>(guess(kk, j));
      R.a[i][j] = s;
    }
    const M3 Rt = m3_transpose(R);
    P.src_idx.clear(); P.tgt_idx.clear();
    for (size_t i = 0; i < N; ++i) {
      const P4 q = transform_mv(T, moved[i]);
      KdTree::Cand nn;
      tree_target.knn(q, 1, &nn);
      if (static_cast<double>(nn.d2) < dist_thr) {
        M3 tmp = m3_mul(m3_mul(R, cov_source[i]), Rt);
        const M3& C2 = cov_target[nn.id];
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) tmp.a[a][b] += C2.a[a][b];
        P.mahalanobis[i] = m3_inverse(tmp);
        P.src_idx.push_back(static_cast<int>(i));
        P.tgt_idx.push_back(static_cast<int>(nn.id));
      }
    }
    out.n_corr = static_cast<uint32_t>(P.src_idx.size());
    prev = T;
    const bool solved = g_gicp_optimizer == 1 ? gicp_bfgs(P, T, cfg.maximum_optimizer_iterations, &out.inner_iterations)
                                              : gicp_newton(P, T, cfg.maximum_optimizer_iterations, &out.inner_iterations);
    if (!solved) break;  // exception path: converged stays false
    double delta = 0;
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) {
      const double ratio = (r < 3 && c < 3) ? 1.0 / cfg.rotation_epsilon : 1.0 / cfg.transformation_epsilon;
      const double cd = ratio * std::fabs(static_cast<double>(prev(r, c)) - static_cast<double>(T(r, c)));
      if (cd > delta) delta = cd;
    }
    ++out.outer_iterations;
    if (std::getenv("S3D_ORACLE_TRACE"))
      std::fprintf(stderr, "[oracle trace] outer=%d inner=%d ncorr=%u delta=%.4g t=(%.9g %.9g %.9g) r10=%.9g r20=%.9g r21=%.9g\n", out.outer_iterations,
                   out.inner_iterations, out.n_corr, delta, T(0, 3), T(1, 3), T(2, 3), T(1, 0), T(2, 0), T(2, 1));
    if (out.outer_iterations >= cfg.maximum_iterations || delta < 1) { out.converged = true; prev = T; }
  }
  out.final_T = m4f_mul(prev, guess);
  // getFitnessScore(max_range = max_correspondence_distance): squared distance vs un-squared range (A.6)
  double sum = 0; int nr = 0;
  for (size_t i = 0; i < N; ++i) {
    const P4 q = transform_se3(out.final_T, pcl_source[i]);
    KdTree::Cand nn;
    tree_target.knn(q, 1, &nn);
    if (static_cast<double>(nn.d2) <= cfg.max_correspondence_distance) { sum += nn.d2; ++nr; }
  }
  out.fitness = nr > 0 ? sum / nr : std::numeric_limits<double>::max();
  return true;
}

#include "ndt_oracle.inc"  // NDT branch (PointCloudSensor.cpp:84-117)

// ------------------------------------------------------------------------------------------------
// slam3d align()   (PointCloudSensor.cpp:119-174)
// ------------------------------------------------------------------------------------------------
static void iso_inverse(const double T[16], double out[16]) {  // column-major rigid inverse
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) out[c * 4 + r] = T[r * 4 + c];
  for (int r = 0; r < 3; ++r) {
    double s = 0;
    for (int c = 0; c < 3; ++c) s += out[c * 4 + r] * T[12 + c];
    out[12 + r] = -s;
  }
  out[3] = out[7] = out[11] = 0; out[15] = 1;
}
static void m4d_mul(const double A[16], const double B[16], double C[16]) {
  for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) {
    double s = 0;
    for (int k = 0; k < 4; ++k) s += A[k * 4 + r] * B[c * 4 + k];
    C[c * 4 + r] = s;
  }
}
// Eigen::AngleAxisd(R).angle(): through the quaternion, angle = 2 atan2(|vec|, |w|) in [0, pi].
static double rotation_angle(const double T[16]) {
  const double m00 = T[0], m11 = T[5], m22 = T[10];
  const double m01 = T[4], m02 = T[8], m10 = T[1], m12 = T[9], m20 = T[2], m21 = T[6];
  double w, x, y, z;
  const double tr = m00 + m11 + m22;
  if (tr > 0) {
    double t = std::sqrt(tr + 1.0); w = 0.5 * t; t = 0.5 / t;
    x = (m21 - m12) * t; y = (m02 - m20) * t; z = (m10 - m01) * t;
  } else {
    int i = 0;
    if (m11 > m00) i = 1;
    if (m22 > (i == 0 ? m00 : m11)) i = 2;
    const double M[3][3] = {{m00, m01, m02}, {m10, m11, m12}, {m20, m21, m22}};
    const int j = (i + 1) % 3, k = (i + 2) % 3;
    double t = std::sqrt(M[i][i] - M[j][j] - M[k][k] + 1.0);
    double q[3];
    q[i] = 0.5 * t; t = 0.5 / t;
    w = (M[k][j] - M[j][k]) * t;
    q[j] = (M[j][i] + M[i][j]) * t;
    q[k] = (M[k][i] + M[i][k]) * t;
    x = q[0]; y = q[1]; z = q[2];
  }
  const double n = std::sqrt(x * x + y * y + z * z);
  return 2.0 * std::atan2(n, std::fabs(w));
}

static int align_impl(const P4* src, size_t nsrc, const P4* tgt, size_t ntgt, const double guess[16],
                      const s3d_registration_parameters& cfg, s3d_result& res) {
  std::memset(&res, 0, sizeof res);
  for (int i = 0; i < 16; ++i) res.T[i] = (i % 5 == 0) ? 1.0 : 0.0;
  std::vector<P4> fsrc, ftgt;
  if (cfg.point_cloud_density > 0) {  // :127-131
    voxel_filter(src, nsrc, static_cast<float>(cfg.point_cloud_density), fsrc, nullptr, nullptr);
    voxel_filter(tgt, ntgt, static_cast<float>(cfg.point_cloud_density), ftgt, nullptr, nullptr);
  } else {
    fsrc.assign(src, src + nsrc);
    ftgt.assign(tgt, tgt + ntgt);
  }
  res.n_source = static_cast<uint32_t>(fsrc.size());
  res.n_target = static_cast<uint32_t>(ftgt.size());
  if (ftgt.size() < 100 || fsrc.size() < 100) {  // :134-135
    g_err = "Too few points after filtering, you may have to decrease 'point_cloud_density'.";
    return res.status = S3D_TOO_FEW_POINTS;
  }
  switch (cfg.registration_algorithm) {  // :139-165
    case S3D_ALG_GICP: break;
    case S3D_ALG_GICP_OMP:
    case S3D_ALG_NDT_OMP:
      g_err = "OMP is not available, you need to rebuild SLAM3D with OMP or use another matching algorithm.";
      return res.status = S3D_UNKNOWN_ALGORITHM;
    case S3D_ALG_NDT: break;
    default:
      g_err = "Unknown registration algorithm specified.";
      return res.status = S3D_UNKNOWN_ALGORITHM;
  }
  M4f guess_f;
  for (int i = 0; i < 16; ++i) guess_f.m[i] = static_cast<float>(guess[i]);  // guess.matrix().cast<float>()  :70
  GicpOutput out{};
  const bool ndt = cfg.registration_algorithm == S3D_ALG_NDT;
  if (ndt) {
    // ndt.setInputSource(target); ndt.setInputTarget(source);   :99-100
    NdtOutput no{};
    if (!ndt_register(ftgt, fsrc, guess_f, cfg, no)) return res.status = S3D_INTERNAL_ERROR;
    out.final_T = no.final_T; out.converged = no.converged; out.outer_iterations = no.outer_iterations;
    out.inner_iterations = no.inner_iterations; out.n_corr = no.n_corr; out.fitness = no.fitness;
  } else
  // icp.setInputSource(target); icp.setInputTarget(source);   :68-69
  if (!gicp_register(ftgt, fsrc, guess_f, cfg, out)) return res.status = S3D_INTERNAL_ERROR;
  for (int i = 0; i < 16; ++i) res.T[i] = static_cast<double>(out.final_T.m[i]);  // Isometry3f -> Transform  :80
  res.fitness = out.fitness;
  res.converged = out.converged ? 1 : 0;
  res.outer_iterations = out.outer_iterations;
  res.inner_iterations = out.inner_iterations;
  res.n_correspondences = out.n_corr;
  if (!out.converged || out.fitness > cfg.max_fitness_score) {  // :74-77
    g_err = std::string(ndt ? "NDT" : "ICP") + " failed with Fitness-Score " + std::to_string(out.fitness) + " > " + std::to_string(cfg.max_fitness_score);
    return res.status = S3D_NOT_CONVERGED;
  }
  double ginv[16], delta[16];  // :167-172
  iso_inverse(guess, ginv);
  m4d_mul(ginv, res.T, delta);
  const double tn = std::sqrt(delta[12] * delta[12] + delta[13] * delta[13] + delta[14] * delta[14]);
  if (tn > cfg.max_translation || rotation_angle(delta) > cfg.max_rotation) {
    g_err = "ICP result is to far away from guess";
    return res.status = S3D_TOO_FAR_FROM_GUESS;
  }
  return res.status = S3D_OK;
}

// ------------------------------------------------------------------------------------------------
// Map building: PointCloudSensor::transform :228-233, getAccumulatedCloud :235-256,
// removeOutliers :211-226 (pcl::RadiusOutlierRemoval), buildMap :301-318.
// ------------------------------------------------------------------------------------------------

// pcl::transformPointCloud(cloud, out, Matrix4d): double se3 form x*c0 + (y*c1 + (z*c2 + c3)), result cast to float.
static inline P4 transform_se3_d(const double T[16], const P4& p) {
  const double x = p.x, y = p.y, z = p.z;
  P4 o;
  o.x = static_cast<float>(x * T[0] + (y * T[4] + (z * T[8] + T[12])));
  o.y = static_cast<float>(x * T[1] + (y * T[5] + (z * T[9] + T[13])));
  o.z = static_cast<float>(x * T[2] + (y * T[6] + (z * T[10] + T[14])));
  o.w = 1.f;
  return o;
}

// pcl::RadiusOutlierRemoval on a dense cloud: nearestKSearch(min_pts + 1) (the query itself included) and the point is
// kept iff the last of them is within the radius:  !(radius^2 < d2_k)  with d2 float, radius^2 double.
static void radius_outlier_removal(const std::vector<P4>& in, double radius, unsigned min_pts, std::vector<P4>& out, uint8_t* keep) {
  out.clear();
  KdTree tree;
  tree.build(in.data(), in.size());
  const int k = static_cast<int>(min_pts) + 1;
  const double r2 = radius * radius;
  std::vector<KdTree::Cand> nb(k);
  for (size_t i = 0; i < in.size(); ++i) {
    const int found = tree.knn(in[i], k, nb.data());
    const bool ok = found == k && !(r2 < static_cast<double>(nb[k - 1].d2));
    if (keep) keep[i] = ok ? 1 : 0;
    if (ok) out.push_back(in[i]);
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C entry points (same shapes as include/s3d_b200.h, prefix s3d_oracle_, host pointers only).
// ------------------------------------------------------------------------------------------------
extern "C" {

const char* s3d_oracle_last_error(void) { return g_err.c_str(); }

void s3d_oracle_default_parameters(s3d_registration_parameters* p) {
  p->registration_algorithm = S3D_ALG_GICP;
  p->point_cloud_density = 0.2;
  p->max_fitness_score = 2.0;
  p->max_translation = 1.0;
  p->max_rotation = 1.0;
  p->euclidean_fitness_epsilon = 1.0;
  p->transformation_epsilon = 1e-5;
  p->max_correspondence_distance = 2.5;
  p->maximum_iterations = 50;
  p->rotation_epsilon = 2e-3;
  p->correspondence_randomness = 20;
  p->maximum_optimizer_iterations = 20;
  p->resolution = 1.0f;
  p->step_size = 0.05;
  p->outlier_ratio = 0.35;
}

int s3d_oracle_voxel_downsample(s3d_cloud in, float leaf, float* out_xyzw, uint64_t* n_out, uint32_t* leaf_index, int32_t* overflow) {
  std::vector<P4> out;
  int ov = 0;
  voxel_filter(reinterpret_cast<const P4*>(in.xyzw), in.n, leaf, out, leaf_index, &ov);
  if (out_xyzw && !out.empty()) std::memcpy(out_xyzw, out.data(), out.size() * sizeof(P4));
  if (n_out) *n_out = out.size();
  if (overflow) *overflow = ov;
  return S3D_OK;
}

int s3d_oracle_knn_covariances(s3d_cloud cloud, int k, uint32_t* knn_index, float* knn_dist2, double* covariances) {
  const P4* p = reinterpret_cast<const P4*>(cloud.xyzw);
  std::vector<P4> pts(p, p + cloud.n);
  KdTree tree;
  tree.build(pts.data(), pts.size());
  std::vector<M3> covs;
  if (!compute_covariances(pts, tree, k, covs, knn_index, knn_dist2)) { g_err = "k larger than cloud"; return S3D_INVALID_ARGUMENT; }
  if (covariances)
    for (size_t i = 0; i < covs.size(); ++i)
      for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) covariances[i * 9 + c * 3 + r] = covs[i].a[r][c];
  return S3D_OK;
}

// Brute-force kNN with the same (d2, index) order: validates the kd-tree.
int s3d_oracle_knn_bruteforce(s3d_cloud reference, s3d_cloud queries, int k, uint32_t* knn_index, float* knn_dist2) {
  const P4* r = reinterpret_cast<const P4*>(reference.xyzw);
  const P4* q = reinterpret_cast<const P4*>(queries.xyzw);
  if (k > static_cast<int>(reference.n)) return S3D_INVALID_ARGUMENT;
  parallel_for(static_cast<int64_t>(queries.n), 0, 64, [&](int64_t i) {
    std::vector<KdTree::Cand> best(k);
    int found = 0;
    for (size_t j = 0; j < reference.n; ++j) {
      KdTree::Cand c{dist2(q[i], r[j]), static_cast<uint32_t>(j)};
      if (found < k) {
        int t = found++;
        while (t > 0 && KdTree::less(c, best[t - 1])) { best[t] = best[t - 1]; --t; }
        best[t] = c;
      } else if (KdTree::less(c, best[k - 1])) {
        int t = k - 1;
        while (t > 0 && KdTree::less(c, best[t - 1])) { best[t] = best[t - 1]; --t; }
        best[t] = c;
      }
    }
    for (int j = 0; j < k; ++j) { knn_index[i * k + j] = best[j].id; knn_dist2[i * k + j] = best[j].d2; }
  });
  return S3D_OK;
}

int s3d_oracle_nearest_neighbors(s3d_cloud reference, s3d_cloud queries, const double* transform, uint32_t* nn_index, float* nn_dist2) {
  const P4* r = reinterpret_cast<const P4*>(reference.xyzw);
  const P4* q = reinterpret_cast<const P4*>(queries.xyzw);
  if (reference.n == 0) return S3D_INVALID_ARGUMENT;
  KdTree tree;
  tree.build(r, reference.n);
  M4f T = m4f_identity();
  if (transform) for (int i = 0; i < 16; ++i) T.m[i] = static_cast<float>(transform[i]);
  for (size_t i = 0; i < queries.n; ++i) {
    const P4 qq = transform ? transform_mv(T, q[i]) : q[i];
    KdTree::Cand nn;
    tree.knn(qq, 1, &nn);
    if (nn_index) nn_index[i] = nn.id;
    if (nn_dist2) nn_dist2[i] = nn.d2;
  }
  return S3D_OK;
}

int s3d_oracle_gicp_align(s3d_cloud source, s3d_cloud target, const double guess[16], const s3d_registration_parameters* params, s3d_result* out) {
  return align_impl(reinterpret_cast<const P4*>(source.xyzw), source.n, reinterpret_cast<const P4*>(target.xyzw), target.n, guess, *params, *out);
}

// One pair per host thread ("all host cores", the analogue of sharding pairs over GPUs).
int s3d_oracle_gicp_align_batch(const s3d_cloud* sources, const s3d_cloud* targets, const double* guesses,
                                const s3d_registration_parameters* params, int n_pairs, int n_threads, s3d_result* out) {
  parallel_for(n_pairs, n_threads, 1, [&](int64_t i) {
    align_impl(reinterpret_cast<const P4*>(sources[i].xyzw), sources[i].n, reinterpret_cast<const P4*>(targets[i].xyzw),
               targets[i].n, guesses + 16 * i, *params, out[i]);
  });
  return S3D_OK;
}

int s3d_oracle_max_threads(void) {
  const unsigned n = std::thread::hardware_concurrency();
  return n ? static_cast<int>(n) : 1;
}

// Test hooks: analytic objective / gradient / Hessian on explicit correspondences, so tests can check the
// derivatives by finite differences.  pts_moved/pts_fixed: m x 4 floats; mahal: m x 9 doubles (col-major).
int s3d_oracle_test_objective(const float* pts_moved, const float* pts_fixed, const double* mahal, int m, const double x[6],
                              double* f, double g[6], double H[36]) {
  std::vector<P4> a(reinterpret_cast<const P4*>(pts_moved), reinterpret_cast<const P4*>(pts_moved) + m);
  std::vector<P4> b(reinterpret_cast<const P4*>(pts_fixed), reinterpret_cast<const P4*>(pts_fixed) + m);
  GicpProblem P;
  P.moved = &a; P.fixed = &b;
  P.mahalanobis.resize(m);
  for (int i = 0; i < m; ++i) {
    P.src_idx.push_back(i); P.tgt_idx.push_back(i);
    for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) P.mahalanobis[i].a[r][c] = mahal[i * 9 + c * 3 + r];
  }
  if (f) *f = gicp_f(P, x);
  if (g && H) {
    double HH[6][6];
    gicp_dfddf(P, x, g, HH);
    for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) H[c * 6 + r] = HH[r][c];
  }
  return S3D_OK;
}

// 0 = Newton (default), 1 = BFGS; process-wide (test infrastructure).  Returns the previous setting.
int s3d_oracle_set_gicp_optimizer(int which) {
  const int old = g_gicp_optimizer;
  g_gicp_optimizer = which == 1 ? 1 : 0;
  return old;
}

// Test hook: estimateRigidTransformationBFGS on explicit correspondences. T: column-major float 4x4 in/out.
int s3d_oracle_test_bfgs(const float* pts_moved, const float* pts_fixed, const double* mahal, int m, float T[16], int max_inner, int* inner_done) {
  std::vector<P4> a(reinterpret_cast<const P4*>(pts_moved), reinterpret_cast<const P4*>(pts_moved) + m);
  std::vector<P4> b(reinterpret_cast<const P4*>(pts_fixed), reinterpret_cast<const P4*>(pts_fixed) + m);
  GicpProblem P;
  P.moved = &a; P.fixed = &b;
  P.mahalanobis.resize(m);
  for (int i = 0; i < m; ++i) {
    P.src_idx.push_back(i); P.tgt_idx.push_back(i);
    for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) P.mahalanobis[i].a[r][c] = mahal[i * 9 + c * 3 + r];
  }
  M4f Tm;
  std::memcpy(Tm.m, T, sizeof Tm.m);
  int done = 0;
  const bool ok = gicp_bfgs(P, Tm, max_inner, &done);
  std::memcpy(T, Tm.m, sizeof Tm.m);
  if (inner_done) *inner_done = done;
  return ok ? S3D_OK : S3D_NOT_CONVERGED;
}

// Test hook: estimateRigidTransformationNewton on explicit correspondences. T: column-major float 4x4 in/out.
int s3d_oracle_test_newton(const float* pts_moved, const float* pts_fixed, const double* mahal, int m, float T[16],
                           int max_inner, int* inner_done) {
  std::vector<P4> a(reinterpret_cast<const P4*>(pts_moved), reinterpret_cast<const P4*>(pts_moved) + m);
  std::vector<P4> b(reinterpret_cast<const P4*>(pts_fixed), reinterpret_cast<const P4*>(pts_fixed) + m);
  GicpProblem P;
  P.moved = &a; P.fixed = &b;
  P.mahalanobis.resize(m);
  for (int i = 0; i < m; ++i) {
    P.src_idx.push_back(i); P.tgt_idx.push_back(i);
    for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) P.mahalanobis[i].a[r][c] = mahal[i * 9 + c * 3 + r];
  }
  M4f Tm;
  std::memcpy(Tm.m, T, sizeof Tm.m);
  int done = 0;
  const bool ok = gicp_newton(P, Tm, max_inner, &done);
  std::memcpy(T, Tm.m, sizeof Tm.m);
  if (inner_done) *inner_done = done;
  return ok ? S3D_OK : S3D_NOT_CONVERGED;
}

int s3d_oracle_test_eigen3(const double A[9], double V[9], double w[3]) {
  double a[3][3], v[3][3];
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) a[r][c] = A[c * 3 + r];
  jacobi_eigen<3>(a, v, w);
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) V[c * 3 + r] = v[r][c];
  return S3D_OK;
}

int s3d_oracle_test_eigen6(const double A[36], double V[36], double w[6]) {
  double a[6][6], v[6][6];
  for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) a[r][c] = A[c * 6 + r];
  jacobi_eigen<6>(a, v, w);
  for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) V[c * 6 + r] = v[r][c];
  return S3D_OK;
}

}  // extern "C"

extern "C" {

int s3d_oracle_transform_cloud(s3d_cloud cloud, const double T[16], float* out_xyzw) {
  const P4* p = reinterpret_cast<const P4*>(cloud.xyzw);
  P4* o = reinterpret_cast<P4*>(out_xyzw);
  for (size_t i = 0; i < cloud.n; ++i) o[i] = transform_se3_d(T, p[i]);
  return S3D_OK;
}

// PointCloudSensor::removeOutliers :211-226 — returns the input unchanged unless size > 0, radius > 0 and min_neighbors > 0
int s3d_oracle_remove_outliers(s3d_cloud cloud, double radius, unsigned min_neighbors, float* out_xyzw, uint64_t* n_out, uint8_t* keep) {
  const P4* p = reinterpret_cast<const P4*>(cloud.xyzw);
  std::vector<P4> in(p, p + cloud.n), out;
  if (cloud.n > 0 && radius > 0 && min_neighbors > 0) radius_outlier_removal(in, radius, min_neighbors, out, keep);
  else { out = in; if (keep) std::fill(keep, keep + cloud.n, 1); }
  if (!out.empty()) std::memcpy(out_xyzw, out.data(), out.size() * sizeof(P4));
  *n_out = out.size();
  return S3D_OK;
}

// FNV-1a over raw bytes: the array hash baseline/doicp_driver.cpp writes into tests/golden/pcl_<version>.json
uint64_t s3d_oracle_fnv1a(const void* p, uint64_t n) {
  const unsigned char* b = static_cast<const unsigned char*>(p);
  uint64_t h = 1469598103934665603ull;
  for (uint64_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
  return h;
}

// PointCloudSensor::createCombinedMeasurement :258-266 with explicit (cloud, pose) lists: getAccumulatedCloud (:235-256,
// transform(cloud_i, pose_i) appended in list order — the order of a build without OpenMP), then pcl::transformPointCloud with
// pose.inverse().matrix(); `inv_patch_pose` is that inverse (the caller forms it, as Eigen::Isometry3d::inverse() does).
int s3d_oracle_combined_measurement(const s3d_cloud* clouds, const double* poses, int n, const double inv_patch_pose[16], float* out_xyzw,
                                    uint64_t* n_out) {
  P4* o = reinterpret_cast<P4*>(out_xyzw);
  size_t m = 0;
  for (int i = 0; i < n; ++i) {
    const P4* p = reinterpret_cast<const P4*>(clouds[i].xyzw);
    for (size_t j = 0; j < clouds[i].n; ++j) o[m++] = transform_se3_d(inv_patch_pose, transform_se3_d(poses + 16 * i, p[j]));
  }
  *n_out = m;
  return S3D_OK;
}

// PointCloudSensor::buildMap :301-318 with explicit (cloud, pose) lists: accumulate in list order, removeOutliers, downsample.
// poses: n x 16 doubles (column-major) = vertex.correctedPose * measurement.sensorPose  (:248)
int s3d_oracle_build_map(const s3d_cloud* clouds, const double* poses, int n, double outlier_radius, unsigned outlier_neighbors,
                         double resolution, float* out_xyzw, uint64_t* n_out) {
  std::vector<P4> accu;
  for (int i = 0; i < n; ++i) {
    const P4* p = reinterpret_cast<const P4*>(clouds[i].xyzw);
    for (size_t j = 0; j < clouds[i].n; ++j) accu.push_back(transform_se3_d(poses + 16 * i, p[j]));
  }
  std::vector<P4> filtered;
  if (!accu.empty() && outlier_radius > 0 && outlier_neighbors > 0) radius_outlier_removal(accu, outlier_radius, outlier_neighbors, filtered, nullptr);
  else filtered = accu;
  std::vector<P4> map;
  if (!filtered.empty()) voxel_filter(filtered.data(), filtered.size(), static_cast<float>(resolution), map, nullptr, nullptr);  // downsample(): empty in, empty out
  if (!map.empty()) std::memcpy(out_xyzw, map.data(), map.size() * sizeof(P4));
  *n_out = map.size();
  return S3D_OK;
}

}  // extern "C"


extern "C" {
// Test hook: NDT score, gradient and Hessian (row-major) of `source` against the voxel grid of `target` at the state x
// (the cloud is transformed by convertTransform(x) as computeStepLengthMT does), for finite-difference checks.
int s3d_oracle_test_ndt_derivatives(s3d_cloud target, s3d_cloud source, float resolution, double outlier_ratio, const double x[6],
                                    double* score, double g[6], double H[36], uint32_t* n_leaves) {
  std::vector<P4> tgt(reinterpret_cast<const P4*>(target.xyzw), reinterpret_cast<const P4*>(target.xyzw) + target.n);
  std::vector<P4> src(reinterpret_cast<const P4*>(source.xyzw), reinterpret_cast<const P4*>(source.xyzw) + source.n);
  NdtGrid G;
  ndt_build_grid(tgt, resolution, G);
  if (n_leaves) *n_leaves = static_cast<uint32_t>(G.leaves.size());
  NdtProblem P;
  P.input = &src; P.grid = &G;
  std::memset(&P.S, 0, sizeof P.S);
  const double res = static_cast<double>(resolution);
  const double c1 = 10.0 * (1 - outlier_ratio), c2 = outlier_ratio / std::pow(res, 3), d3 = -std::log(c2);
  P.S.gauss_d1 = -std::log(c1 + c2) - d3;
  P.S.gauss_d2 = -2 * std::log((-std::log(c1 * std::exp(-0.5) + c2) - d3) / P.S.gauss_d1);
  P.S.r2 = static_cast<float>(res * res);
  for (int r = 0; r < 3; ++r) P.S.pj[r][r] = 1.0;
  const M4f T = ndt_convert_transform(x);
  std::vector<P4> tc(src.size());
  for (size_t i = 0; i < src.size(); ++i) tc[i] = transform_se3(T, src[i]);
  double gg[6], HH[6][6];
  *score = ndt_compute_derivatives(P, gg, HH, tc, x, true);
  for (int i = 0; i < 6; ++i) { g[i] = gg[i]; for (int j = 0; j < 6; ++j) H[6 * i + j] = HH[i][j]; }
  return S3D_OK;
}
}
