// gicp.cu — the Generalized-ICP loop of pcl::GeneralizedIterativeClosestPoint as driven by slam3d's doICP
// (slam3d/sensor/pcl/PointCloudSensor.cpp:52-82) and the fitness score of :73.   Semantics: SURVEY A.4-A.6.
//
// The loop runs in ROUNDS of four launches for the whole batch; every pair carries its own phase, so pairs advance
// independently and the host only polls an "active pairs" counter:
//   gicp_iter_kernel   pairs that start an outer iteration.  One thread per moving point:  q = T_f32 * (guess_f32 * p)  ->
//                      exact 1-NN in the fixed cloud (nn_search.cuh; skipped when the previous correspondence is provably
//                      still nearest)  ->  d2 < max_corr^2  ->  M = (R C1 R^T + C2)^-1 (FP64, stored)  ->  the CTA reduces
//                      the point's 14 features to the 74 sums of gicp_math.h in a fixed order: 60 x-independent sums that
//                      give the whole Hessian, and the 13 residual sums + count of PCL's first objective evaluation.
//   gicp_ctrl_kernel   one CTA per pair: fixed-order sum of the tile partials, then ONE thread advances PCL's inner
//                      optimiser (estimateRigidTransformationNewton as a resumable state machine) to its next objective
//                      evaluation, or finishes the outer iteration (convergence test of computeTransformation).
//   gicp_eval_kernel   pairs whose optimiser asked for f at a trial state: one pass over the stored correspondences with the
//                      residual formed exactly as PCL forms it (float32 T(x)*p, float subtraction) -> 13 sums per tile.
//   gicp_ctrl_kernel   again.
// (Tried and dropped, B200, 64 pairs per step: a separate search kernel in which one thread handles 2/3/4 consecutive points
// and passes each result on as a hint for the next — 10.1 / 11.6 / 13.0 ms per step against 8.6 ms: the saved instructions
// do not make up for the probe latency that fewer threads in flight can hide.  Also tried and dropped: splitting this kernel
// into a barrier-free search kernel (CTAs of 32 / 64 / 128 / 256 threads) and a separate 74-sum tile kernel, because 37 % of the
// stall samples of a late-iteration launch sit at the reduction barrier (profiles/r01g_summary.md) — results bit-identical,
// stage time 4.6-4.9 ms against 4.55 ms fused per 32 pairs: the waiting warps cost no issue slots and the extra launch does.
// Also dropped: per-lane cell cursors (+12 %) and CTA-level compaction of the points that still need a search (equal):
// profiles/r01h_summary.md.
// Kept: the two-pass 27-block of nn_search.cuh (scan_block<true>; cells noted in s_cells) — gicp_iter 8.60 -> 8.26 ms per 64 pairs.)
// Because every float operation that PCL's decisions depend on is mirrored and the double sums differ only in order, the
// GPU follows the oracle's iterate sequence (same inner/outer iteration counts, bit-identical poses in the test-suite).
// Reductions use fixed trees: results are bit-reproducible run to run and independent of batch composition.
// Roofline: per outer iteration 16 B (moving point) + 32 B (its normal) + gathered 16 B + 32 B (fixed point + normal) per
// correspondence, + 48 B (M) written; per evaluation 16 + 16 + 48 B.  The working set is L2 resident; the search is latency
// bound on the hash probes.
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "internal.h"
#include "nn_search.cuh"

namespace s3d {

constexpr int kFeat = 14;  // M00 M01 M02 M11 M12 M22 | px py pz | (Md)0 (Md)1 (Md)2 | d^T M d | valid

struct MomentSpec { uint8_t a, b, c; };  // moment = sum f[a]*f[b]*f[c]
__constant__ MomentSpec c_spec[kNumMoments];

static void build_moment_spec(MomentSpec* spec) {
  const int Mi[6] = {0, 1, 2, 3, 4, 5};
  const int phi[4] = {6, 7, 8, 13};
  int ab = 0;
  for (int a = 0; a < 3; ++a)
    for (int b = a; b < 3; ++b, ++ab) {
      int ce = 0;
      for (int c = 0; c < 4; ++c)
        for (int e = c; e < 4; ++e, ++ce) spec[ab * 10 + ce] = {(uint8_t)Mi[ab], (uint8_t)phi[c], (uint8_t)phi[e]};
    }
  for (int a = 0; a < 3; ++a) for (int c = 0; c < 4; ++c) spec[60 + a * 4 + c] = {(uint8_t)(9 + a), (uint8_t)phi[c], 13};
  spec[72] = {12, 13, 13};
  spec[73] = {13, 13, 13};
}

// R = top-left 3x3 of double(T) * double(guess);  RRt = R R^T      (SURVEY A.4 "R <- ...")
__device__ void update_rotation(PairState& ps) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 4; ++k) s += (double)ps.T[k * 4 + i] * (double)ps.guess[j * 4 + k];
      ps.R[i][j] = s;
    }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += ps.R[i][k] * ps.R[j][k];
      ps.RRt[i][j] = s;
    }
}

// Eigen Matrix4f * Matrix4f: ((a0*b0 + a1*b1) + a2*b2) + a3*b3, column-major
__device__ void mat4f_mul(const float* A, const float* B, float* C) {
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c)
      C[c * 4 + r] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(A[r], B[c * 4]), __fmul_rn(A[4 + r], B[c * 4 + 1])), __fmul_rn(A[8 + r], B[c * 4 + 2])),
                               __fmul_rn(A[12 + r], B[c * 4 + 3]));
}

// start of an outer iteration: x0 from transformation_, the float matrix PCL's first evaluation uses, and R for the Mahalanobis matrices
__device__ void begin_outer(PairState& ps) {
  newton_begin(ps.nst, ps.T);
  matrix_from_state(ps.nst.xc, ps.T_eval);
  update_rotation(ps);
  ps.phase = kPhaseNeedNN;
}

// Registration::align set-up: gates of align() :134-135, output = guess * input (transformPointCloud), state reset.
__global__ void __launch_bounds__(256) gicp_prepare_kernel(const SlotInfo* __restrict__ slots, PairState* __restrict__ pairs,
                                                           float4* __restrict__ moved, uint32_t* __restrict__ prev_nn, float* __restrict__ sec_lb,
                                                           int32_t* __restrict__ flags) {
  const uint32_t p = blockIdx.y;
  PairState& ps = pairs[p];
  const SlotInfo& sb = slots[2 * p];      // fixed cloud B  (slam3d source = PCL target)
  const SlotInfo& sa = slots[2 * p + 1];  // moving cloud A (slam3d target = PCL source)
  const bool enough = sa.n_pts >= 100 && sb.n_pts >= 100;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    ps.active = enough ? 1 : 0;
    ps.phase = kPhaseFinished;
    if (enough) { begin_outer(ps); atomicAdd(&flags[1], 1); }
  }
  if (!enough) return;
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= sa.n_pts) return;
  const float4 v = sa.gpts[r];
  const float3 m = transform_se3(ps.guess, v.x, v.y, v.z);
  moved[ps.pt_off + r] = make_float4(m.x, m.y, m.z, v.w);
  prev_nn[ps.pt_off + r] = kNoIndex;
  sec_lb[ps.pt_off + r] = 0.f;
}

// Temporal-coherence certificate (exactness preserving).  The last search for this point ran at position q_old and left
// (j, lb): every fixed point other than j was at least `lb` away from q_old.  If the query moved by m and
// dist(q_new, p_j) < lb - m, then j is still the unique nearest neighbour at q_new and no search is needed; the bound for
// the next iteration becomes lb - m.  Generous relative slack (1e-5) covers the float rounding of dist2_pcl (4e-7).
__device__ __forceinline__ bool certified_same_nn(const GridView& g, float3 q, float3 q_old, uint32_t j, float lb, NNResult& nn, float& lb_new) {
  if (j >= g.n || !(lb > 0.f)) return false;
  const float mx = q.x - q_old.x, my = q.y - q_old.y, mz = q.z - q_old.z;
  const float m = sqrtf(mx * mx + my * my + mz * mz) * 1.00001f + 1e-7f;
  const float rest = (lb - m) * 0.99999f;
  if (!(rest > 0.f)) return false;
  const float4 v = __ldg(g.pts + j);
  const float d2 = dist2_pcl(q.x, q.y, q.z, v.x, v.y, v.z);
  if (!(d2 * 1.00001f < rest * rest)) return false;
  nn.d2 = d2; nn.idx = __float_as_uint(v.w); nn.pos = j; nn.lb2 = rest * rest;
  lb_new = rest;
  return true;
}

// (register allocation: the compiler's own choice — 64 registers, 4 CTAs/SM — beat forced 5/6/7 CTAs/SM on B200)
#ifndef S3D_NN_GATHER
#define S3D_NN_GATHER 1
#endif
#if S3D_NN_GATHER
#define S3D_ITER_BOUNDS __launch_bounds__(kIterTile, 4)  // 64 registers: 4 CTAs per SM as before (the compiler's own choice would be 78)
#else
#define S3D_ITER_BOUNDS __launch_bounds__(kIterTile)
#endif
__global__ void S3D_ITER_BOUNDS gicp_iter_kernel(const SlotInfo* __restrict__ slots, const PairState* __restrict__ pairs,
                                                              const float4* __restrict__ moved,
                                                              uint32_t* __restrict__ prev_nn, float* __restrict__ sec_lb, uint32_t* __restrict__ corr,
                                                              double* __restrict__ mahal, double* __restrict__ moments) {
  __shared__ double feat[kIterTile][kFeat];
  __shared__ double part[3][kNumMoments];
#if S3D_NN_GATHER
  __shared__ uint2 s_cells[kNNGatherCap * kIterTile];  // cell ranges noted by the gathering search (nn_search.cuh)
#endif
  const uint32_t p = blockIdx.y;
  const PairState& ps = pairs[p];
  if (ps.phase != kPhaseNeedNN) return;
  const SlotInfo& sb = slots[2 * p];
  const SlotInfo& sa = slots[2 * p + 1];
  const uint32_t first = blockIdx.x * kIterTile;
  if (first >= sa.n_pts) return;
  const uint32_t r = first + threadIdx.x;
  double f[kFeat];
#pragma unroll
  for (int i = 0; i < kFeat; ++i) f[i] = 0.0;
  if (r < sa.n_pts) {
    const GridView g = make_grid_view(sb);
    const float4 mv = moved[ps.pt_off + r];
    const float3 q = transform_mv(ps.T, mv.x, mv.y, mv.z);
    const double thr = ps.max_corr2;
    const float cutoff = __double2float_ru(thr);
    const uint32_t hint = prev_nn[ps.pt_off + r];
    NNResult nn;
    float lb_new;
    if (!certified_same_nn(g, q, transform_mv(ps.T_search, mv.x, mv.y, mv.z), hint, sec_lb[ps.pt_off + r], nn, lb_new)) {
#if S3D_NN_GATHER
      nn = nn_search<true>(g, q.x, q.y, q.z, cutoff, hint, kNoIndex, s_cells + threadIdx.x);
#else
      nn = nn_search(g, q.x, q.y, q.z, cutoff, hint);
#endif
      prev_nn[ps.pt_off + r] = nn.pos;
      lb_new = sqrtf(nn.lb2) * 0.99999f;
    }
    sec_lb[ps.pt_off + r] = lb_new;
    uint32_t c = kNoIndex;
    if (nn.pos != kNoIndex && (double)nn.d2 < thr) {
      c = nn.pos;
      const double4 n1 = sa.normals[r];
      const double4 n2 = sb.normals[nn.pos];
      double a[3], b[3] = {n2.x, n2.y, n2.z};
      for (int i = 0; i < 3; ++i) a[i] = ps.R[i][0] * n1.x + ps.R[i][1] * n1.y + ps.R[i][2] * n1.z;
      double M[6];
      mahalanobis6(ps.RRt, a, b, M);
      double* mo = mahal + 6 * (size_t)(ps.pt_off + r);
#pragma unroll
      for (int i = 0; i < 6; ++i) mo[i] = M[i];
      // PCL's first objective evaluation of this outer iteration: d = float(T(x0) * p) - q, float subtraction, then double
      const float4 qb = g.pts[nn.pos];
      const float3 pp = transform_mv(ps.T_eval, mv.x, mv.y, mv.z);
      const double d0 = (double)__fsub_rn(pp.x, qb.x), d1 = (double)__fsub_rn(pp.y, qb.y), d2 = (double)__fsub_rn(pp.z, qb.z);
      const double Md0 = M[0] * d0 + M[1] * d1 + M[2] * d2;
      const double Md1 = M[1] * d0 + M[3] * d1 + M[4] * d2;
      const double Md2 = M[2] * d0 + M[4] * d1 + M[5] * d2;
#pragma unroll
      for (int i = 0; i < 6; ++i) f[i] = M[i];
      f[6] = mv.x; f[7] = mv.y; f[8] = mv.z;
      f[9] = Md0; f[10] = Md1; f[11] = Md2;
      f[12] = d0 * Md0 + d1 * Md1 + d2 * Md2;
      f[13] = 1.0;
    }
    corr[ps.pt_off + r] = c;
  }
#pragma unroll
  for (int i = 0; i < kFeat; ++i) feat[threadIdx.x][i] = f[i];
  __syncthreads();
  // 74 sums x 3 sub-ranges of the tile, each summed in ascending point order
  if (threadIdx.x < 3 * kNumMoments) {
    const int m = threadIdx.x % kNumMoments, sub = threadIdx.x / kNumMoments;
    const MomentSpec sp = c_spec[m];
    const int lo = sub * 86, hi = min(kIterTile, lo + 86);
    double s = 0.0;
    for (int i = lo; i < hi; ++i) s += feat[i][sp.a] * feat[i][sp.b] * feat[i][sp.c];
    part[sub][m] = s;
  }
  __syncthreads();
  if (threadIdx.x < kNumMoments) {
    const size_t tile = (size_t)p * gridDim.x + blockIdx.x;
    moments[tile * kNumMoments + threadIdx.x] = (part[0][threadIdx.x] + part[1][threadIdx.x]) + part[2][threadIdx.x];
  }
}

// Objective evaluations at the pending back-tracking trials of every pair in kPhaseEval: 13 residual sums per trial and
// 256-point tile, residual formed exactly as PCL forms it.
__global__ void __launch_bounds__(kIterTile) gicp_eval_kernel(const SlotInfo* __restrict__ slots, const PairState* __restrict__ pairs,
                                                              const float4* __restrict__ moved,
                                                              const uint32_t* __restrict__ corr, const double* __restrict__ mahal,
                                                              double* __restrict__ eval_part) {
  __shared__ double feat[kIterTile][8];  // px py pz | (Md)0..2 | d^T M d | 1
  __shared__ double part[16][kEvalSums];
  const uint32_t p = blockIdx.y;
  const PairState& ps = pairs[p];
  if (ps.phase != kPhaseEval) return;
  const SlotInfo& sb = slots[2 * p];
  const SlotInfo& sa = slots[2 * p + 1];
  const uint32_t first = blockIdx.x * kIterTile;
  if (first >= sa.n_pts) return;
  const uint32_t r = first + threadIdx.x;
  bool valid = false;
  float4 mv = make_float4(0.f, 0.f, 0.f, 0.f), qb = mv;
  double M[6] = {0, 0, 0, 0, 0, 0};
  if (r < sa.n_pts) {
    const uint32_t c = corr[ps.pt_off + r];
    if (c != kNoIndex) {
      valid = true;
      mv = moved[ps.pt_off + r];
      qb = sb.gpts[c];
      const double* Mp = mahal + 6 * (size_t)(ps.pt_off + r);
#pragma unroll
      for (int i = 0; i < 6; ++i) M[i] = Mp[i];
    }
  }
  const size_t tile = (size_t)p * gridDim.x + blockIdx.x;
  for (int j = ps.trial_first; j < ps.trial_first + ps.trial_count; ++j) {
    double f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (valid) {
      const float3 pp = transform_mv(ps.T_trial[j], mv.x, mv.y, mv.z);
      const double d0 = (double)__fsub_rn(pp.x, qb.x), d1 = (double)__fsub_rn(pp.y, qb.y), d2 = (double)__fsub_rn(pp.z, qb.z);
      const double Md0 = M[0] * d0 + M[1] * d1 + M[2] * d2;
      const double Md1 = M[1] * d0 + M[3] * d1 + M[4] * d2;
      const double Md2 = M[2] * d0 + M[4] * d1 + M[5] * d2;
      f[0] = mv.x; f[1] = mv.y; f[2] = mv.z; f[3] = Md0; f[4] = Md1; f[5] = Md2;
      f[6] = d0 * Md0 + d1 * Md1 + d2 * Md2; f[7] = 1.0;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) feat[threadIdx.x][i] = f[i];
    __syncthreads();
    // sum s (0..11: (Md)_a * phi_c with a = s / 4, c = s % 4; 12: d^T M d) over 16 sub-ranges of 16 points, ascending order
    if (threadIdx.x < 16 * kEvalSums) {
      const int sidx = threadIdx.x % kEvalSums, sub = threadIdx.x / kEvalSums;
      const int ia = sidx < 12 ? 3 + sidx / 4 : 6;
      const int ib = sidx < 12 ? (sidx % 4 < 3 ? sidx % 4 : 7) : 7;
      double acc = 0.0;
      for (int i = sub * 16; i < sub * 16 + 16; ++i) acc += feat[i][ia] * feat[i][ib];
      part[sub][sidx] = acc;
    }
    __syncthreads();
    if (threadIdx.x < kEvalSums) {
      double acc = 0.0;
      for (int sub = 0; sub < 16; ++sub) acc += part[sub][threadIdx.x];
      eval_part[(tile * kLineSearchTrials + j) * kEvalSums + threadIdx.x] = acc;
    }
  }
}

// End of an outer iteration (computeTransformation): convergence test, counters, next iteration or final transform.
__device__ void finish_outer(PairState& ps, bool optimiser_failed, int32_t* flags) {
  bool stop = false;
  if (optimiser_failed) {
    ps.failed = 1; ps.converged = 0; stop = true;  // optimiser exception: loop breaks, converged_ stays false, final = previous * guess
  } else {
    float T[16];
    matrix_from_state(ps.nst.x, T);  // transformation_matrix.setIdentity(); applyState(transformation_matrix, x)
    ps.inner_iterations += ps.nst.it;
    double delta = 0.0;
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < 4; ++c) {
        const double ratio = (r < 3 && c < 3) ? 1.0 / ps.rot_eps : 1.0 / ps.trans_eps;
        const double cd = ratio * fabs((double)ps.prev[c * 4 + r] - (double)T[c * 4 + r]);
        if (cd > delta) delta = cd;
      }
    for (int i = 0; i < 16; ++i) ps.T[i] = T[i];
    ps.outer_iterations += 1;
    if (ps.outer_iterations >= ps.max_iter || delta < 1.0) {
      ps.converged = 1; stop = true;
      for (int i = 0; i < 16; ++i) ps.prev[i] = T[i];  // previous_transformation_ = transformation_
    }
  }
  if (stop) {
    mat4f_mul(ps.prev, ps.guess, ps.final_T);  // final_transformation_ = previous_transformation_ * guess
    ps.active = 0;
    ps.phase = kPhaseFinished;
    atomicSub(&flags[1], 1);
  } else {
    begin_outer(ps);
  }
}

// euler_derivs(x, E, true) spread over the CTA: threads 0..2 build the factor matrices of one axis each (the three sincos),
// threads 0..9 one product (A B) C each, with the same operand order as the serial function, so E is bit-identical.
__device__ void euler_derivs_cta(const double* x, Euler& E, double (*fac)[3][3][3]) {  // fac[axis][order] in shared memory
  const int t = threadIdx.x;
  if (t < 3) euler_factor(x[3 + t], t, fac[t][0], fac[t][1], fac[t][2]);  // axis 0 = X(phi), 1 = Y(theta), 2 = Z(psi)
  __syncthreads();
  if (t < 10) {
    // derivative orders (z, y, x) of: R | dR0 dR1 dR2 | ddR00 ddR11 ddR22 | ddR01 ddR02 ddR12
    const int oz = (t == 3) ? 1 : (t == 6) ? 2 : (t == 8 || t == 9) ? 1 : 0;
    const int oy = (t == 2) ? 1 : (t == 5) ? 2 : (t == 7 || t == 9) ? 1 : 0;
    const int ox = (t == 1) ? 1 : (t == 4) ? 2 : (t == 7 || t == 8) ? 1 : 0;
    double T[3][3], O[3][3];
    mat3_mul(fac[2][oz], fac[1][oy], T);
    mat3_mul(T, fac[0][ox], O);
    double (*dst)[3] = t == 0 ? E.R : t == 1 ? E.dR[0] : t == 2 ? E.dR[1] : t == 3 ? E.dR[2] : t == 4 ? E.ddR[0][0] : t == 5 ? E.ddR[1][1]
                     : t == 6 ? E.ddR[2][2] : t == 7 ? E.ddR[0][1] : t == 8 ? E.ddR[0][2] : E.ddR[1][2];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) dst[i][j] = O[i][j];
    if (t >= 7) {  // mixed second derivatives are symmetric in (k, l)
      double (*sym)[3] = t == 7 ? E.ddR[1][0] : t == 8 ? E.ddR[2][0] : E.ddR[2][1];
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) sym[i][j] = O[i][j];
    }
  }
  __syncthreads();
}

// one CTA per pair: ordered reduction of the tile partials of the pass that just ran, then the optimiser advances
__global__ void __launch_bounds__(256) gicp_ctrl_kernel(const SlotInfo* __restrict__ slots, PairState* __restrict__ pairs,
                                                        const double* __restrict__ moments, const double* __restrict__ eval_part,
                                                        uint32_t tiles_per_pair, int after_eval, int32_t* __restrict__ flags) {
  __shared__ double part[3][kLineSearchTrials * kEvalSums];
  __shared__ double red[kLineSearchTrials * kEvalSums];
  __shared__ Euler E;
  __shared__ double fac[3][3][3][3];
  __shared__ double gH[42];
  __shared__ int mode;      // 0: nothing to assemble, 1: assemble the objective at ps.nst.xc from ps.sums
  __shared__ int newstep;   // a new Newton step started: its trial matrices are needed
  const uint32_t p = blockIdx.x;
  PairState& ps = pairs[p];
  // the first ctrl launch of a round serves pairs that just searched, the later ones pairs that were just evaluated
  if (ps.phase != (after_eval ? kPhaseEval : kPhaseNeedNN)) return;
  const uint32_t n_tiles = (slots[2 * p + 1].n_pts + kIterTile - 1) / kIterTile;
  const int n_sums = after_eval ? ps.trial_count * kEvalSums : kNumMoments;
  const int stride = after_eval ? kLineSearchTrials * kEvalSums : kNumMoments;
  const double* src0 = after_eval ? eval_part + (size_t)p * tiles_per_pair * stride + ps.trial_first * kEvalSums
                                  : moments + (size_t)p * tiles_per_pair * stride;
  const int subs = n_sums * 3 <= 256 ? 3 : (n_sums * 2 <= 256 ? 2 : 1);
  if ((int)threadIdx.x < subs * n_sums) {
    const int m = threadIdx.x % n_sums, sub = threadIdx.x / n_sums;
    const uint32_t per = (n_tiles + subs - 1) / subs;
    const uint32_t lo = sub * per, hi = min(n_tiles, lo + per);
    const double* src = src0 + m;
    double s = 0.0;
#pragma unroll 4
    for (uint32_t t = lo; t < hi; ++t) s += src[(size_t)t * stride];
    part[sub][m] = s;
  }
  __syncthreads();
  if ((int)threadIdx.x < n_sums) {
    double s = part[0][threadIdx.x];
    for (int sub = 1; sub < subs; ++sub) s += part[sub][threadIdx.x];
    red[threadIdx.x] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mode = 1; newstep = 0;
    if (!after_eval) {
      for (int i = 0; i < kNumMoments; ++i) ps.sums[i] = red[i];
      ps.n_corr = (uint32_t)ps.sums[73];
      for (int i = 0; i < 16; ++i) { ps.prev[i] = ps.T[i]; ps.T_search[i] = ps.T[i]; }  // previous_transformation_ = transformation_
      if (ps.sums[73] < 4.0) { finish_outer(ps, true, flags); mode = 0; }  // min_number_correspondences_: PCL throws
    } else {
      double f_trial[kLineSearchTrials];
      for (int j = 0; j < ps.trial_count; ++j) f_trial[ps.trial_first + j] = red[j * kEvalSums + 12] / ps.sums[73];
      const int j = newton_pick_trial(ps.nst, f_trial, ps.trial_first, ps.trial_count);
      if (j >= 0) {
        newton_select_trial(ps.nst, j);
        for (int i = 0; i < kEvalSums; ++i) ps.sums[60 + i] = red[(j - ps.trial_first) * kEvalSums + i];
      } else if (ps.trial_first == 0 && kLineSearchTrials > 1) {
        ps.trial_first = 1; ps.trial_count = kLineSearchTrials - 1;  // trial 0 did not improve: evaluate all remaining trials in one pass
        mode = 0;
      } else {
        ps.nst.phase = 2;  // no improvement found
        finish_outer(ps, false, flags);
        mode = 0;
      }
    }
  }
  __syncthreads();
  if (!mode) return;
  euler_derivs_cta(ps.nst.xc, E, fac);
  // objective at the evaluated state: 42 threads contract one gradient / Hessian entry each
  if (threadIdx.x < 42) gH[threadIdx.x] = objective_entry(ps.sums, E, threadIdx.x);
  __syncthreads();
  if (threadIdx.x == 0) {
    double g[6], H[6][6];
    for (int i = 0; i < 6; ++i) { g[i] = gH[i]; for (int j = 0; j < 6; ++j) H[i][j] = gH[6 + 6 * i + j]; }
    if (newton_advance_pre(ps.nst, ps.sums[72] / ps.sums[73], g, H, ps.max_inner)) {
      ps.trial_first = 0; ps.trial_count = 1;
      ps.phase = kPhaseEval;
      newstep = 1;
    } else {
      finish_outer(ps, false, flags);
    }
  }
  __syncthreads();
  if (newstep && threadIdx.x < kLineSearchTrials) {  // float matrices of all back-tracking trials of the new step
    double xc[6];
    newton_trial_state(ps.nst, threadIdx.x, xc);
    matrix_from_state(xc, ps.T_trial[threadIdx.x]);
  }
}

// getFitnessScore(max_range): transformPointCloud(input, final), 1-NN, d2 <= max_range (sic), mean of d2   (A.6)
__global__ void __launch_bounds__(kIterTile) gicp_fitness_kernel(const SlotInfo* __restrict__ slots, const PairState* __restrict__ pairs,
                                                                 const float4* __restrict__ moved, const uint32_t* __restrict__ prev_nn,
                                                                 const float* __restrict__ sec_lb, double* __restrict__ fit_partial) {
  __shared__ double ssum[kIterTile];
  __shared__ uint32_t scnt[kIterTile];
  const uint32_t p = blockIdx.y;
  const PairState& ps = pairs[p];
  const SlotInfo& sb = slots[2 * p];
  const SlotInfo& sa = slots[2 * p + 1];
  if (sa.n_pts < 100 || sb.n_pts < 100) return;
  const uint32_t first = blockIdx.x * kIterTile;
  if (first >= sa.n_pts) return;
  const uint32_t r = first + threadIdx.x;
  double s = 0.0; uint32_t c = 0;
  if (r < sa.n_pts) {
    const GridView g = make_grid_view(sb);
    const float4 v = sa.gpts[r];
    const float3 q = transform_se3(ps.final_T, v.x, v.y, v.z);
    const float4 mv = moved[ps.pt_off + r];
    const uint32_t hint = prev_nn[ps.pt_off + r];
    NNResult nn;
    float lb_new;
    if (!certified_same_nn(g, q, transform_mv(ps.T_search, mv.x, mv.y, mv.z), hint, sec_lb[ps.pt_off + r], nn, lb_new))
      nn = nn_search(g, q.x, q.y, q.z, __double2float_ru(ps.fit_range), hint);
    if (nn.pos != kNoIndex && (double)nn.d2 <= ps.fit_range) { s = (double)nn.d2; c = 1; }
  }
  ssum[threadIdx.x] = s; scnt[threadIdx.x] = c;
  __syncthreads();
  for (int o = kIterTile / 2; o > 0; o >>= 1) {  // fixed tree
    if (threadIdx.x < o) { ssum[threadIdx.x] += ssum[threadIdx.x + o]; scnt[threadIdx.x] += scnt[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const size_t tile = (size_t)p * gridDim.x + blockIdx.x;
    fit_partial[2 * tile] = ssum[0];
    fit_partial[2 * tile + 1] = (double)scnt[0];
  }
}

__global__ void gicp_fitness_reduce_kernel(const SlotInfo* __restrict__ slots, PairState* __restrict__ pairs, const double* __restrict__ fit_partial,
                                           uint32_t tiles_per_pair, uint32_t n_pairs) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pairs) return;
  PairState& ps = pairs[p];
  const SlotInfo& sb = slots[2 * p];
  const SlotInfo& sa = slots[2 * p + 1];
  ps.fit_sum = 0.0; ps.fit_n = 0;
  if (sa.n_pts < 100 || sb.n_pts < 100) return;
  const uint32_t n_tiles = (sa.n_pts + kIterTile - 1) / kIterTile;
  double s = 0.0, c = 0.0;
  for (uint32_t t = 0; t < n_tiles; ++t) { s += fit_partial[2 * ((size_t)p * tiles_per_pair + t)]; c += fit_partial[2 * ((size_t)p * tiles_per_pair + t) + 1]; }
  ps.fit_sum = s; ps.fit_n = (uint32_t)c;
}

// ------------------------------------------------------------------------------------------------------------------
static void iso_inverse(const double T[16], double out[16]) {
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) out[c * 4 + r] = T[r * 4 + c];
  for (int r = 0; r < 3; ++r) { double s = 0; for (int c = 0; c < 3; ++c) s += out[c * 4 + r] * T[12 + c]; out[12 + r] = -s; }
  out[3] = out[7] = out[11] = 0; out[15] = 1;
}
static void m4d_mul(const double A[16], const double B[16], double C[16]) {
  for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) { double s = 0; for (int k = 0; k < 4; ++k) s += A[k * 4 + r] * B[c * 4 + k]; C[c * 4 + r] = s; }
}
// Eigen::AngleAxisd(R).angle() — via the quaternion, in [0, pi]   (align() :168)
static double rotation_angle(const double T[16]) {
  const double m00 = T[0], m11 = T[5], m22 = T[10], m01 = T[4], m02 = T[8], m10 = T[1], m12 = T[9], m20 = T[2], m21 = T[6];
  double w, x, y, z;
  const double tr = m00 + m11 + m22;
  if (tr > 0) {
    double t = sqrt(tr + 1.0); w = 0.5 * t; t = 0.5 / t;
    x = (m21 - m12) * t; y = (m02 - m20) * t; z = (m10 - m01) * t;
  } else {
    int i = 0;
    if (m11 > m00) i = 1;
    if (m22 > (i == 0 ? m00 : m11)) i = 2;
    const double M[3][3] = {{m00, m01, m02}, {m10, m11, m12}, {m20, m21, m22}};
    const int j = (i + 1) % 3, k = (i + 2) % 3;
    double t = sqrt(M[i][i] - M[j][j] - M[k][k] + 1.0);
    double q[3];
    q[i] = 0.5 * t; t = 0.5 / t;
    w = (M[k][j] - M[j][k]) * t; q[j] = (M[j][i] + M[i][j]) * t; q[k] = (M[k][i] + M[i][k]) * t;
    x = q[0]; y = q[1]; z = q[2];
  }
  return 2.0 * atan2(sqrt(x * x + y * y + z * z), fabs(w));
}

double rotation_angle_of(const double T[16]) { return rotation_angle(T); }  // shared with ndt.cu

void check_arena(Workspace& ws, const int32_t* h_flags) {
  if (h_flags[0] & kErrHashArena) throw ArenaOverflow{(size_t)h_flags[3] + (size_t)h_flags[3] / 8 + 64};
}

// Runs the GICP loop + fitness for every pair of the batch (grids and covariances must be ready) and fills `out`
// with the decisions of doICP (:74-77) and align() (:134-135, :167-172).
void run_gicp(Workspace& ws, const std::vector<s3d_registration_parameters>& params, const double* guesses, s3d_result* out) {
  static bool spec_ready[64] = {false};
  if (!spec_ready[ws.device]) {
    MomentSpec spec[kNumMoments];
    build_moment_spec(spec);
    S3D_CUDA(cudaMemcpyToSymbol(c_spec, spec, sizeof spec));
    spec_ready[ws.device] = true;
  }
  const uint32_t np = ws.n_pairs;
  cudaStream_t st = ws.stream;
  const uint32_t max_na = ws.max_na;
  const uint32_t tiles_per_pair = std::max<uint32_t>(1, (max_na + kIterTile - 1) / kIterTile);
  ws.pairs.reserve(sizeof(PairState) * np);
  ws.h_pairs.reserve(sizeof(PairState) * np);
  ws.moved.reserve(16 * std::max<size_t>(ws.total, 4));
  ws.prev_nn.reserve(4 * std::max<size_t>(ws.total, 4));
  ws.sec_lb.reserve(4 * std::max<size_t>(ws.total, 4));
  ws.moments.reserve(sizeof(double) * kNumMoments * size_t(tiles_per_pair) * np);
  ws.eval_part.reserve(sizeof(double) * kEvalSums * kLineSearchTrials * size_t(tiles_per_pair) * np);
  ws.corr.reserve(4 * std::max<size_t>(ws.total, 4));
  ws.mahal.reserve(48 * std::max<size_t>(ws.total, 4));
  ws.fit_partial.reserve(sizeof(double) * 2 * size_t(tiles_per_pair) * np);
  PairState* hp = ws.h_pairs.as<PairState>();
  int max_iter = 0;
  for (uint32_t p = 0; p < np; ++p) {
    const s3d_registration_parameters& cfg = params[p];
    PairState& ps = hp[p];
    memset(&ps, 0, sizeof ps);
    for (int i = 0; i < 16; ++i) {
      ps.guess[i] = (float)guesses[16 * p + i];  // guess.matrix().cast<float>()  :70
      ps.T[i] = ps.prev[i] = ps.final_T[i] = (i % 5 == 0) ? 1.f : 0.f;
    }
    ps.max_corr2 = cfg.max_correspondence_distance * cfg.max_correspondence_distance;
    ps.rot_eps = cfg.rotation_epsilon; ps.trans_eps = cfg.transformation_epsilon;
    ps.fit_range = cfg.max_correspondence_distance;
    ps.pt_off = ws.pair_off[p];
    ps.max_iter = cfg.maximum_iterations; ps.max_inner = cfg.maximum_optimizer_iterations; ps.k = cfg.correspondence_randomness;
    max_iter = std::max(max_iter, cfg.maximum_iterations);
  }
  S3D_CUDA(cudaMemcpyAsync(ws.pairs.p, hp, sizeof(PairState) * np, cudaMemcpyHostToDevice, st));
  const SlotInfo* slots = ws.slots.as<SlotInfo>();
  PairState* pairs = ws.pairs.as<PairState>();
  int32_t* flags = ws.flags.as<int32_t>();
  int32_t* h_flags = ws.h_small.as<int32_t>();
  dim3 grid(tiles_per_pair, np);
  gicp_prepare_kernel<<<grid, 256, 0, st>>>(slots, pairs, ws.moved.as<float4>(), ws.prev_nn.as<uint32_t>(), ws.sec_lb.as<float>(), flags);
  ++ws.launches;
  // Rounds: [search + first evaluation] -> ctrl -> [trial evaluation] -> ctrl.  A pair needs one round per outer iteration
  // plus one per extra objective evaluation; PCL's loop is a do-while, so maximum_iterations <= 0 still runs one iteration.
  const bool trace = getenv("S3D_TRACE") != nullptr;
  const long max_rounds = (long)std::max(max_iter, 1) * 64 + 64;
  auto t_round = std::chrono::steady_clock::now();
  for (long round = 0; round < max_rounds; ++round) {
    {
      StageTimer timer(ws, kStageIter);
      gicp_iter_kernel<<<grid, kIterTile, 0, st>>>(slots, pairs, ws.moved.as<float4>(), ws.prev_nn.as<uint32_t>(), ws.sec_lb.as<float>(), ws.corr.as<uint32_t>(),
                                                   ws.mahal.as<double>(), ws.moments.as<double>());
      ++ws.launches;
    }
    {
      StageTimer timer(ws, kStageSolve);
      gicp_ctrl_kernel<<<np, 256, 0, st>>>(slots, pairs, ws.moments.as<double>(), ws.eval_part.as<double>(), tiles_per_pair, 0, flags);
      ++ws.launches;
      for (int e = 0; e < 2; ++e) {  // two evaluate/advance passes per round: most outer iterations finish without another host poll
        gicp_eval_kernel<<<grid, kIterTile, 0, st>>>(slots, pairs, ws.moved.as<float4>(), ws.corr.as<uint32_t>(),
                                                     ws.mahal.as<double>(), ws.eval_part.as<double>());
        gicp_ctrl_kernel<<<np, 256, 0, st>>>(slots, pairs, ws.moments.as<double>(), ws.eval_part.as<double>(), tiles_per_pair, 1, flags);
        ws.launches += 2;
      }
    }
    S3D_CUDA(cudaMemcpyAsync(h_flags, flags, 16, cudaMemcpyDeviceToHost, st));
    S3D_CUDA(cudaStreamSynchronize(st));
    ws.d2h += 16;
    if (trace) {
      S3D_CUDA(cudaMemcpy(hp, pairs, sizeof(PairState) * np, cudaMemcpyDeviceToHost));
      int n_nn = 0, n_ev = 0, n_fin = 0;
      for (uint32_t p = 0; p < np; ++p) { n_nn += hp[p].phase == kPhaseNeedNN; n_ev += hp[p].phase == kPhaseEval; n_fin += hp[p].phase == kPhaseFinished; }
      const auto now = std::chrono::steady_clock::now();
      fprintf(stderr, "[s3d trace] round=%ld dt=%.1f us  after round: need_nn=%d eval=%d finished=%d\n", round,
              std::chrono::duration<double, std::micro>(now - t_round).count(), n_nn, n_ev, n_fin);
      t_round = std::chrono::steady_clock::now();
      if (np <= 4)
      for (uint32_t p = 0; p < np; ++p)
        fprintf(stderr, "[s3d trace] round=%ld pair=%u phase=%d outer=%d inner=%d ncorr=%u t=(%.9g %.9g %.9g) r10=%.9g r20=%.9g r21=%.9g\n", round, p,
                hp[p].phase, hp[p].outer_iterations, hp[p].inner_iterations, hp[p].n_corr, hp[p].T[12], hp[p].T[13], hp[p].T[14], hp[p].T[1],
                hp[p].T[2], hp[p].T[6]);
    }
    if (h_flags[1] <= 0) break;
  }
  {
    StageTimer timer(ws, kStageFitness);
    gicp_fitness_kernel<<<grid, kIterTile, 0, st>>>(slots, pairs, ws.moved.as<float4>(),
                                                    ws.prev_nn.as<uint32_t>(), ws.sec_lb.as<float>(), ws.fit_partial.as<double>());
    gicp_fitness_reduce_kernel<<<(np + 63) / 64, 64, 0, st>>>(slots, pairs, ws.fit_partial.as<double>(), tiles_per_pair, np);
    ws.launches += 2;
  }
  S3D_CUDA(cudaMemcpyAsync(hp, pairs, sizeof(PairState) * np, cudaMemcpyDeviceToHost, st));
  SlotInfo* hs = ws.h_slots.as<SlotInfo>();
  S3D_CUDA(cudaMemcpyAsync(hs, slots, sizeof(SlotInfo) * ws.n_slots, cudaMemcpyDeviceToHost, st));
  S3D_CUDA(cudaMemcpyAsync(h_flags, flags, 16, cudaMemcpyDeviceToHost, st));
  S3D_CUDA(cudaStreamSynchronize(st));
  ws.d2h += sizeof(PairState) * np + sizeof(SlotInfo) * ws.n_slots + 16;
  ws.collect_spans();
  check_arena(ws, h_flags);
  for (uint32_t p = 0; p < np; ++p) {
    const s3d_registration_parameters& cfg = params[p];
    const PairState& ps = hp[p];
    s3d_result& r = out[p];
    memset(&r, 0, sizeof r);
    for (int i = 0; i < 16; ++i) r.T[i] = (i % 5 == 0) ? 1.0 : 0.0;
    r.n_source = hs[2 * p].n_pts; r.n_target = hs[2 * p + 1].n_pts;
    if (r.n_target < 100 || r.n_source < 100) { r.status = S3D_TOO_FEW_POINTS; continue; }  // :134-135
    for (int i = 0; i < 16; ++i) r.T[i] = (double)ps.final_T[i];  // Transform(Eigen::Isometry3f(final))  :80
    r.fitness = ps.fit_n > 0 ? ps.fit_sum / (double)ps.fit_n : 1.7976931348623157e308;
    r.converged = ps.converged; r.outer_iterations = ps.outer_iterations; r.inner_iterations = ps.inner_iterations;
    r.n_correspondences = ps.n_corr;
    if (!ps.converged || r.fitness > cfg.max_fitness_score) { r.status = S3D_NOT_CONVERGED; continue; }  // :74-77
    double ginv[16], delta[16];
    iso_inverse(guesses + 16 * p, ginv);
    m4d_mul(ginv, r.T, delta);
    const double tn = sqrt(delta[12] * delta[12] + delta[13] * delta[13] + delta[14] * delta[14]);
    r.status = (tn > cfg.max_translation || rotation_angle(delta) > cfg.max_rotation) ? S3D_TOO_FAR_FROM_GUESS : S3D_OK;  // :167-172
  }
}

}  // namespace s3d
