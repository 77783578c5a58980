// knn.cu — GICP computeCovariances: exact k-nearest-neighbour search + 3x3 covariance + regularisation, fused.
//
// Replaces GeneralizedIterativeClosestPoint::computeCovariances (reached from icp.align(), PointCloudSensor.cpp:70;
// SURVEY A.3, 8a row a4.2 — the largest single block of CPU time, ~270 ms per cloud).
//   * one warp per query, 32 consecutive (Morton-adjacent) queries per warp;
//   * lanes 0..26 probe the 27 cells of the block, a warp prefix sum flattens their point ranges, and the 32 lanes
//     stream the candidates as coalesced float4 loads; the k best are kept as a sorted list distributed over the lanes
//     in lexicographic (d2, original index) order — FLANN's exact search with ties to the lowest index;
//   * the level (cell size) is picked per query from the occupancy of its own ancestors and widened until the k-th
//     distance is certified by the block's coverage radius (nn_search.cuh);
//   * moments exactly as PCL: float products accumulated in double, in neighbour order (A.3 step 2);
//   * the neighbour list never leaves the SM: per query only a 32-byte unit normal is written, because
//     C = U diag(1,1,eps) U^T = I - (1-eps) n n^T;  the 3x3 eigen-decompositions run one query per lane.
// Bound: L2/latency (the working set of a scan, ~2 MB, is L2 resident); algorithmic bytes 16 B read + 32 B written
// per point.
#include "internal.h"
#include "nn_search.cuh"

namespace s3d {

constexpr int kKnnQueriesPerWarp = 32;
constexpr int kKnnWarps = 8;
constexpr int kKnnTile = kKnnQueriesPerWarp * kKnnWarps;  // 256 queries per CTA

__global__ void __launch_bounds__(kKnnWarps * 32) knn_cov_kernel(const SlotInfo* __restrict__ slots, const HashEntry* __restrict__ arena,
                                                                 const float4* __restrict__ gpts, const float4* __restrict__ work,
                                                                 double4* __restrict__ normals, int k, uint32_t* __restrict__ knn_index,
                                                                 float* __restrict__ knn_dist2) {
  const SlotInfo& si = slots[blockIdx.y];
  const uint32_t n = si.n_pts;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t q0 = blockIdx.x * kKnnTile + warp * kKnnQueriesPerWarp;
  if (q0 >= n) return;
  const GridView g = make_grid_view(si, arena, gpts);
  const float4* cloud = work + si.off;  // original (voxel-key) order
  const uint32_t FULL = 0xFFFFFFFFu;
  const int kk = k < (int)n ? k : (int)n;  // FLANN clamps k to the cloud size

  double my_cov[6] = {0, 0, 0, 0, 0, 0};  // covariance of query q0 + lane, filled in as the warp walks its queries

#pragma unroll 1
  for (int qi = 0; qi < kKnnQueriesPerWarp; ++qi) {
    const uint32_t r = q0 + qi;
    if (r >= n) break;
    const float4 qv = g.pts[r];
    const uint32_t q_orig = __float_as_uint(qv.w);
    const float ux = clamp_coord(grid_coord(qv.x, g.ox, g.inv_h0));
    const float uy = clamp_coord(grid_coord(qv.y, g.oy, g.inv_h0));
    const float uz = clamp_coord(grid_coord(qv.z, g.oz, g.inv_h0));

    // ---- start level: smallest L whose parent cell (level L+1) already holds >= 12 points --------------------------
    int L = g.nlev - 1;
    {
      const int lv = lane + 1;  // lane probes level lane+1
      bool enough = false;
      if (lv < g.nlev) {
        uint32_t b, e;
        const int cx = (int)floorf(ux) >> lv, cy = (int)floorf(uy) >> lv, cz = (int)floorf(uz) >> lv;
        if (cell_range(g.table, g.cap, g.nlev, lv, cx, cy, cz, b, e)) enough = (e - b) >= 12u;
      }
      const uint32_t m = __ballot_sync(FULL, enough);
      if (m) L = __ffs(m) - 1;
    }

    // lane l holds the l-th best candidate as one 64-bit key: (float bits of d2) << 32 | original index.
    // d2 >= +0 so the bit pattern orders like the value; lexicographic (d2, idx) == unsigned key order.
    const uint64_t KMAX = 0xFFFFFFFFFFFFFFFFull;
    uint64_t lk = KMAX;        // sorted ascending over lanes; lanes >= kk stay KMAX
    uint64_t bound = KMAX;     // inclusive bound carried over from a finer level
    uint64_t tau = KMAX;       // current k-th best (KMAX while the list is not full)
    for (;; ++L) {
      int cx, cy, cz;
      const float g2 = block_guarantee2(g, ux, uy, uz, L, cx, cy, cz);
      const bool top = L >= g.nlev - 1;
      if (top) cx = cy = cz = 0;
      // lanes 0..26: one cell each
      uint32_t cb = 0, cn = 0;
      if (lane < 27) {
        uint32_t b, e;
        if (cell_range(g.table, g.cap, g.nlev, L, cx + lane % 3 - 1, cy + (lane / 3) % 3 - 1, cz + lane / 9 - 1, b, e)) { cb = b; cn = e - b; }
      }
      uint32_t incl = cn;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += v; }
      const uint32_t total = __shfl_sync(FULL, incl, 31);
      const uint32_t excl = incl - cn;
      const int32_t seg_base = (int32_t)cb - (int32_t)excl;  // sorted position = candidate number + seg_base
      lk = KMAX; tau = KMAX;
      for (uint32_t base = 0; base < total; base += 32) {
        const uint32_t c = base + lane;
        // segment of candidate c: last lane whose exclusive offset <= c (binary search over lanes by shuffles)
        int seg = 0;
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
          const uint32_t v = __shfl_sync(FULL, excl, seg + step);
          if (v <= c) seg += step;
        }
        const int32_t sb = __shfl_sync(FULL, seg_base, seg);
        uint64_t ck = KMAX;
        if (c < total) {
          const float4 v = __ldg(g.pts + (int32_t)c + sb);
          const float cd = dist2_pcl(qv.x, qv.y, qv.z, v.x, v.y, v.z);
          if (cd == cd) {  // NaN never enters
            ck = ((uint64_t)__float_as_uint(cd) << 32) | (uint64_t)__float_as_uint(v.w);
            if (ck > bound) ck = KMAX;
          }
        }
        uint32_t m = __ballot_sync(FULL, ck < tau);
        if (!m) continue;
        if (__popc(m) <= 3) {
          // few newcomers: insert one by one
          while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const uint64_t x = __shfl_sync(FULL, ck, src);
            if (x >= tau) continue;  // tau moved
            const int pos = __popc(__ballot_sync(FULL, lk < x));
            const uint64_t up = __shfl_up_sync(FULL, lk, 1);
            if (lane == pos) lk = x; else if (lane > pos) lk = up;
            if (lane >= kk) lk = KMAX;
            tau = __shfl_sync(FULL, lk, kk - 1);
          }
        } else {
          // many newcomers: bitonic-sort the 32 candidates, then one bitonic merge with the list
          uint64_t v = ck;
#pragma unroll
          for (int kb = 2; kb <= 32; kb <<= 1) {
#pragma unroll
            for (int j = kb >> 1; j > 0; j >>= 1) {
              const uint64_t o = __shfl_xor_sync(FULL, v, j);
              const bool keep_min = ((lane & j) == 0) == ((lane & kb) == 0);
              v = keep_min ? (v < o ? v : o) : (v < o ? o : v);
            }
          }
          const uint64_t r = __shfl_sync(FULL, v, 31 - lane);  // descending
          v = lk < r ? lk : r;                                  // the 32 smallest of the union, bitonic
#pragma unroll
          for (int j = 16; j > 0; j >>= 1) {
            const uint64_t o = __shfl_xor_sync(FULL, v, j);
            v = ((lane & j) == 0) ? (v < o ? v : o) : (v < o ? o : v);
          }
          lk = lane < kk ? v : KMAX;
          tau = __shfl_sync(FULL, lk, kk - 1);
        }
      }
      const bool full = tau != KMAX;
      const float td = __uint_as_float((uint32_t)(tau >> 32));
      if ((full && td <= g2) || top) break;
      if (full) bound = tau;
    }
    const uint32_t li = (uint32_t)lk;                           // neighbour `lane`: original index
    const float ld = lk == KMAX ? INFINITY : __uint_as_float((uint32_t)(lk >> 32));

    // ---- moments in neighbour order, float products accumulated in double (A.3 step 2) ------------------------------
    // lane a (< 9) owns one accumulator: mean x,y,z | cov 00,10,11,20,21,22; every lane walks the neighbours in order.
    float px = 0.f, py = 0.f, pz = 0.f;
    if (lane < kk && lk != KMAX) { const float4 v = __ldg(cloud + li); px = v.x; py = v.y; pz = v.z; }
    const int ia = lane < 3 ? lane : (lane == 3 ? 0 : (lane <= 5 ? 1 : 2));
    const int ib = lane < 3 ? 3 : (lane == 3 || lane == 4 || lane == 6 ? 0 : (lane == 5 || lane == 7 ? 1 : 2));
    double acc = 0.0;
    for (int j = 0; j < kk; ++j) {
      const float x = __shfl_sync(FULL, px, j), y = __shfl_sync(FULL, py, j), z = __shfl_sync(FULL, pz, j);
      const float u = ia == 0 ? x : (ia == 1 ? y : z);
      const float w = ib == 0 ? x : (ib == 1 ? y : (ib == 2 ? z : 1.0f));
      acc += (double)__fmul_rn(u, w);
    }
    double s9[9];
#pragma unroll
    for (int a = 0; a < 9; ++a) s9[a] = __shfl_sync(FULL, acc, a);
    if (lane == qi) {
      const double dk = (double)k;  // PCL divides by k_correspondences_
      const double m0 = s9[0] / dk, m1 = s9[1] / dk, m2 = s9[2] / dk;
      my_cov[0] = s9[3] / dk - m0 * m0;
      my_cov[1] = s9[4] / dk - m1 * m0;
      my_cov[2] = s9[5] / dk - m1 * m1;
      my_cov[3] = s9[6] / dk - m2 * m0;
      my_cov[4] = s9[7] / dk - m2 * m1;
      my_cov[5] = s9[8] / dk - m2 * m2;
    }
    if (knn_index && lane < k) knn_index[((size_t)si.off + q_orig) * k + lane] = li;
    if (knn_dist2 && lane < k) knn_dist2[((size_t)si.off + q_orig) * k + lane] = ld;
  }

  // ---- one eigen-decomposition per lane ---------------------------------------------------------------------------
  const uint32_t r = q0 + lane;
  if (r < n) {
    double c[3][3] = {{my_cov[0], my_cov[1], my_cov[3]}, {my_cov[1], my_cov[2], my_cov[4]}, {my_cov[3], my_cov[4], my_cov[5]}};
    double nrm[3];
    smallest_eigenvector3(c, nrm);
    normals[si.off + r] = make_double4(nrm[0], nrm[1], nrm[2], 0.0);
  }
}

// stage-API helper: full regularised covariance per ORIGINAL index, column-major 3x3
__global__ void expand_cov_kernel(const SlotInfo* __restrict__ slots, const float4* __restrict__ gpts, const double4* __restrict__ normals,
                                  double* __restrict__ cov_out) {
  const SlotInfo& si = slots[blockIdx.y];
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= si.n_pts) return;
  const uint32_t orig = __float_as_uint(gpts[si.off + r].w);
  const double4 nv = normals[si.off + r];
  const double nn[3] = {nv.x, nv.y, nv.z};
  double* o = cov_out + ((size_t)si.off + orig) * 9;
  for (int c = 0; c < 3; ++c) for (int rr = 0; rr < 3; ++rr) o[c * 3 + rr] = (rr == c ? 1.0 : 0.0) - (1.0 - kGicpEpsilon) * nn[rr] * nn[c];
}

void run_knn_covariances(Workspace& ws, int k, uint32_t* knn_index, float* knn_dist2) {
  if (ws.n_tiles == 0) return;
  uint32_t max_n = 0;
  for (uint32_t s = 0; s < ws.n_slots; ++s) max_n = std::max(max_n, ws.h_n[s]);
  ws.normals.reserve(sizeof(double4) * std::max<size_t>(ws.total, 4));
  StageTimer timer(ws, kStageKnn);
  dim3 grid((max_n + kKnnTile - 1) / kKnnTile, ws.n_slots);
  knn_cov_kernel<<<grid, kKnnWarps * 32, 0, ws.stream>>>(ws.slots.as<SlotInfo>(), ws.hash.as<HashEntry>(), ws.gpts.as<float4>(),
                                                         ws.work.as<float4>(), ws.normals.as<double4>(), k, knn_index, knn_dist2);
  ++ws.launches;
  S3D_CUDA(cudaGetLastError());
}

void run_expand_cov(Workspace& ws, double* cov_out) {
  uint32_t max_n = 0;
  for (uint32_t s = 0; s < ws.n_slots; ++s) max_n = std::max(max_n, ws.h_n[s]);
  if (max_n == 0) return;
  dim3 grid((max_n + 255) / 256, ws.n_slots);
  expand_cov_kernel<<<grid, 256, 0, ws.stream>>>(ws.slots.as<SlotInfo>(), ws.gpts.as<float4>(), ws.normals.as<double4>(), cov_out);
  ++ws.launches;
}

// ---- stage API: exact 1-NN of T*query in a reference slot (thread per query) --------------------------------------
__global__ void __launch_bounds__(256) nn_stage_kernel(const SlotInfo* __restrict__ slots, const HashEntry* __restrict__ arena,
                                                       const float4* __restrict__ gpts, uint32_t ref_slot, uint32_t qry_slot,
                                                       const float* __restrict__ T, uint32_t* __restrict__ nn_index, float* __restrict__ nn_dist2) {
  const SlotInfo& rs = slots[ref_slot];
  const SlotInfo& qs = slots[qry_slot];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= qs.n_raw) return;
  const GridView g = make_grid_view(rs, arena, gpts);
  const float4 v = qs.raw[i];
  float3 q = make_float3(v.x, v.y, v.z);
  if (T) q = transform_mv(T, v.x, v.y, v.z);
  const NNResult r = nn_search(g, q.x, q.y, q.z, INFINITY, kNoIndex);
  if (nn_index) nn_index[i] = r.idx;
  if (nn_dist2) nn_dist2[i] = r.d2;
}

void run_nn_stage(Workspace& ws, uint32_t ref_slot, uint32_t qry_slot, const float* T16_dev, uint32_t* nn_index, float* nn_dist2) {
  const uint32_t nq = ws.h_n[qry_slot];
  if (nq == 0) return;
  nn_stage_kernel<<<(nq + 255) / 256, 256, 0, ws.stream>>>(ws.slots.as<SlotInfo>(), ws.hash.as<HashEntry>(), ws.gpts.as<float4>(), ref_slot,
                                                           qry_slot, T16_dev, nn_index, nn_dist2);
  ++ws.launches;
  S3D_CUDA(cudaGetLastError());
}

}  // namespace s3d
