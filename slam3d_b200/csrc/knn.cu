// knn.cu — GICP computeCovariances: exact k-nearest-neighbour search + 3x3 covariance + regularisation, fused.
//
// Replaces GeneralizedIterativeClosestPoint::computeCovariances (reached from icp.align(), PointCloudSensor.cpp:70;
// SURVEY A.3, 8a row a4.2 — the largest single block of CPU time, ~270 ms per cloud).
//   * one thread per query; the 128 queries of a CTA are Morton-adjacent, so the cell walks of a warp touch the same
//     L1 lines (a warp-cooperative variant, one warp per query with a shuffle-merged top-k, cost 6x more instructions
//     per query: profiles/r01a_summary.md);
//   * candidates are 64-bit keys (d2 bits << 32 | original index): unsigned order == lexicographic (d2, index) order,
//     i.e. FLANN's exact search with ties to the lowest index; the k best sit in a per-thread max-heap in shared memory;
//   * the level (cell size) is picked per query from the occupancy of its own ancestors and widened until the k-th
//     distance is certified by the block's coverage radius (nn_search.cuh); cells farther than the k-th best are skipped;
//   * moments exactly as PCL: float products accumulated in double, in neighbour order (A.3 step 2);
//   * the neighbour list never leaves the SM: per query only a 32-byte unit normal is written, because
//     C = U diag(1,1,eps) U^T = I - (1-eps) n n^T.
// Bound: L2/latency (the working set of a scan, ~2 MB, is L2 resident); algorithmic bytes 16 B read + 32 B written
// per point.
#include "internal.h"
#include "knn_walk.cuh"

namespace s3d {

// Moments exactly as PCL (float products accumulated in double, in neighbour order = FLANN's ascending (d2, idx) order; A.3
// step 2), smallest eigenvector; optional neighbour lists for the stage API.
//
// The neighbour ORDER only matters where a double sum rounds.  Every term is a float (24-bit significand) widened to double,
// so a sum of k of them is exact — hence the same in any order — whenever the binary exponents of its non-zero terms span
// few enough bits: with e in [lo, hi] per axis and L = ceil(log2 k), all partial sums of x_i are multiples of 2^(lo-23) below
// 2^(hi+1+L), and those of the products x_i*y_i multiples of 2^(lox+loy-23) below 2^(hix+hiy+2+L); both fit 53 bits if every
// axis has hi - lo <= (28 - L) / 2 (11 for k = 20).  The align() path therefore sums the heap in storage order while tracking
// the exponent range (a few integer min/max per neighbour) and only falls back to heapsort + ordered sums when that exactness
// certificate fails (a neighbourhood that straddles a coordinate plane within ~1e-4 of its extent) or when the caller wants the
// sorted lists.  The result is bit-identical to the ordered sums either way; the heapsort was 18 % of this kernel.
template <typename Heap>
__device__ __forceinline__ void knn_finish(const SlotInfo& si, const float4* __restrict__ cloud, const Heap& h, int cnt, int k, uint32_t r, uint32_t q_orig,
                                           double4* __restrict__ normals, uint32_t* __restrict__ knn_index, float* __restrict__ knn_dist2) {
  double mean[3] = {0, 0, 0}, cov[6] = {0, 0, 0, 0, 0, 0};  // cov: 00,10,11,20,21,22
  bool ordered = knn_index != nullptr || knn_dist2 != nullptr;
  if (!ordered) {
    uint32_t elo[3] = {255u, 255u, 255u}, ehi[3] = {0u, 0u, 0u};
    for (int j = 0; j < cnt; ++j) {
      const float4 pt = __ldg(cloud + (uint32_t)h.at(j));
      const float c3v[3] = {pt.x, pt.y, pt.z};
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const uint32_t e = (__float_as_uint(c3v[a]) >> 23) & 255u;
        if (c3v[a] != 0.f) { elo[a] = min(elo[a], e); ehi[a] = max(ehi[a], e); }  // exact zeros add nothing; denormals (e = 0) fail the test below
      }
      mean[0] += (double)pt.x; mean[1] += (double)pt.y; mean[2] += (double)pt.z;
      cov[0] += (double)__fmul_rn(pt.x, pt.x);
      cov[1] += (double)__fmul_rn(pt.y, pt.x);
      cov[2] += (double)__fmul_rn(pt.y, pt.y);
      cov[3] += (double)__fmul_rn(pt.z, pt.x);
      cov[4] += (double)__fmul_rn(pt.z, pt.y);
      cov[5] += (double)__fmul_rn(pt.z, pt.z);
    }
    const int L = 32 - __clz(max(cnt, 2) - 1);  // ceil(log2 cnt)
    const uint32_t span = (uint32_t)max(0, (28 - L) / 2);
#pragma unroll
    for (int a = 0; a < 3; ++a)
      if (ehi[a] >= elo[a] && (ehi[a] - elo[a] > span || elo[a] < 64u || ehi[a] > 190u)) ordered = true;  // products must stay normal floats
    if (ordered) { for (int a = 0; a < 3; ++a) mean[a] = 0.0; for (int a = 0; a < 6; ++a) cov[a] = 0.0; }
  }
  if (ordered) {
    for (int m = cnt - 1; m > 0; --m) {  // heapsort: ascending (d2, idx)
      const uint64_t last = h.at(m);
      h.at(m) = h.at(0);
      heap_sift_down(h, m, last);
    }
    for (int j = 0; j < cnt; ++j) {
      const uint64_t key = h.at(j);
      const uint32_t id = (uint32_t)key;
      const float4 pt = __ldg(cloud + id);
      mean[0] += (double)pt.x; mean[1] += (double)pt.y; mean[2] += (double)pt.z;
      cov[0] += (double)__fmul_rn(pt.x, pt.x);
      cov[1] += (double)__fmul_rn(pt.y, pt.x);
      cov[2] += (double)__fmul_rn(pt.y, pt.y);
      cov[3] += (double)__fmul_rn(pt.z, pt.x);
      cov[4] += (double)__fmul_rn(pt.z, pt.y);
      cov[5] += (double)__fmul_rn(pt.z, pt.z);
      if (knn_index) knn_index[((size_t)si.off + q_orig) * k + j] = id;
      if (knn_dist2) knn_dist2[((size_t)si.off + q_orig) * k + j] = __uint_as_float((uint32_t)(key >> 32));
    }
    for (int j = cnt; j < k; ++j) {
      if (knn_index) knn_index[((size_t)si.off + q_orig) * k + j] = kNoIndex;
      if (knn_dist2) knn_dist2[((size_t)si.off + q_orig) * k + j] = INFINITY;
    }
  }
  const double dk = (double)k;  // PCL divides by k_correspondences_
  mean[0] /= dk; mean[1] /= dk; mean[2] /= dk;
  double c3[3][3];
  c3[0][0] = cov[0] / dk - mean[0] * mean[0];
  c3[1][0] = c3[0][1] = cov[1] / dk - mean[1] * mean[0];
  c3[1][1] = cov[2] / dk - mean[1] * mean[1];
  c3[2][0] = c3[0][2] = cov[3] / dk - mean[2] * mean[0];
  c3[2][1] = c3[1][2] = cov[4] / dk - mean[2] * mean[1];
  c3[2][2] = cov[5] / dk - mean[2] * mean[2];
  double nrm[3];
  smallest_eigenvector3(c3, nrm);
  normals[si.off + r] = make_double4(nrm[0], nrm[1], nrm[2], 0.0);
}

template <typename Heap>
__device__ __forceinline__ void knn_cov_query(const SlotInfo& si, const GridView& g, const float4* __restrict__ work, double4* __restrict__ normals, int k, uint32_t r,
                                              const Heap& h, uint32_t* __restrict__ knn_index, float* __restrict__ knn_dist2) {
  const int kk = k < (int)si.n_pts ? k : (int)si.n_pts;  // FLANN clamps k to the cloud size
  const float4 qv = g.pts[r];
  const float ux = clamp_coord(grid_coord(qv.x, g.ox, g.inv_h0));
  const float uy = clamp_coord(grid_coord(qv.y, g.oy, g.inv_h0));
  const float uz = clamp_coord(grid_coord(qv.z, g.oz, g.inv_h0));
  KSTAT(0, 1);
  const int cnt = thread_walk(g, qv, ux, uy, uz, knn_start_level(g, ux, uy, uz), KMAX, h, kk);
  if (cnt < kk) for (int i = heap_last_parent(cnt); i >= 0; --i) heap_sift_down(h, cnt, h.at(i), i);  // top level ended before the list filled
  knn_finish(si, work + si.off, h, cnt, k, r, __float_as_uint(qv.w), normals, knn_index, knn_dist2);
}

__global__ void __launch_bounds__(kKnnThreads, S3D_KNN_MINBLOCKS) knn_cov_kernel(const SlotInfo* __restrict__ slots, const float4* __restrict__ work,
                                                              double4* __restrict__ normals, int k, uint32_t* __restrict__ knn_index,
                                                              float* __restrict__ knn_dist2) {
  extern __shared__ uint64_t heap_smem[];  // k * kKnnThreads keys
  const SlotInfo& si = slots[blockIdx.y];
  if (blockIdx.x * kKnnThreads >= si.n_pts) return;
  const GridView g = make_grid_view(si);
  if (g.cap == 0) return;  // grid build overflowed its arena: the host re-runs the batch
  // the grid is sized from an estimate of the filtered size (Workspace::grid_frac): stride over the chunks it did not cover
  for (uint32_t r = blockIdx.x * kKnnThreads + threadIdx.x; r < si.n_pts; r += gridDim.x * kKnnThreads)
    knn_cov_query(si, g, work, normals, k, r, SmemHeap{heap_smem + threadIdx.x}, knn_index, knn_dist2);
}

// correspondence_randomness beyond what the shared-memory heap holds: the same walk on a heap in global memory (one column per
// thread of the launch).  PCL accepts any k; this keeps the GPU path from refusing what the reference runs.
__global__ void __launch_bounds__(kKnnThreads) knn_cov_bigk_kernel(const SlotInfo* __restrict__ slots, const float4* __restrict__ work,
                                                                   double4* __restrict__ normals, int k, uint64_t* __restrict__ arena,
                                                                   uint32_t* __restrict__ knn_index, float* __restrict__ knn_dist2) {
  const SlotInfo& si = slots[blockIdx.y];
  if (blockIdx.x * kKnnThreads >= si.n_pts) return;
  const GridView g = make_grid_view(si);
  if (g.cap == 0) return;
  const size_t stride = (size_t)gridDim.x * gridDim.y * kKnnThreads;
  const size_t column = ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * kKnnThreads + threadIdx.x;
  for (uint32_t r = blockIdx.x * kKnnThreads + threadIdx.x; r < si.n_pts; r += gridDim.x * kKnnThreads)
    knn_cov_query(si, g, work, normals, k, r, GlobalHeap{arena + column, stride}, knn_index, knn_dist2);
}

// stage-API helper: full regularised covariance per ORIGINAL index, column-major 3x3
__global__ void expand_cov_kernel(const SlotInfo* __restrict__ slots, double* __restrict__ cov_out) {
  const SlotInfo& si = slots[blockIdx.y];
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= si.n_pts || si.hash_cap == 0) return;  // hash_cap == 0: the grid build overflowed, the batch is being re-run
  const uint32_t orig = __float_as_uint(si.gpts[r].w);
  const double4 nv = si.normals[r];
  const double nn[3] = {nv.x, nv.y, nv.z};
  double* o = cov_out + ((size_t)si.off + orig) * 9;
  for (int c = 0; c < 3; ++c) for (int rr = 0; rr < 3; ++rr) o[c * 3 + rr] = (rr == c ? 1.0 : 0.0) - (1.0 - kGicpEpsilon) * nn[rr] * nn[c];
}

void run_knn_covariances(Workspace& ws, int k, uint32_t* knn_index, float* knn_dist2) {
  if (ws.n_tiles == 0) return;
  uint32_t max_n = 0;
  for (uint32_t s = 0; s < ws.n_slots; ++s) max_n = std::max(max_n, ws.h_n[s]);
  StageTimer timer(ws, kStageKnn);
  const uint32_t chunks = (max_n + kKnnThreads - 1) / kKnnThreads;
  dim3 grid(ws.grid_x(chunks, ws.n_slots, S3D_KNN_MINBLOCKS), ws.n_slots);
  const size_t heap_bytes = sizeof(uint64_t) * k * kKnnThreads;
  if (k <= kMaxKShared) {
    if (heap_bytes > 48 * 1024) S3D_CUDA(cudaFuncSetAttribute(knn_cov_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)heap_bytes));
    knn_cov_kernel<<<grid, kKnnThreads, heap_bytes, ws.stream>>>(ws.slots.as<SlotInfo>(), ws.work.as<float4>(), ws.normals.as<double4>(), k, knn_index, knn_dist2);
  } else {
    ws.knn_arena.reserve(heap_bytes * grid.x * grid.y);
    knn_cov_bigk_kernel<<<grid, kKnnThreads, 0, ws.stream>>>(ws.slots.as<SlotInfo>(), ws.work.as<float4>(), ws.normals.as<double4>(), k, ws.knn_arena.as<uint64_t>(),
                                                             knn_index, knn_dist2);
  }
  ++ws.launches;
  S3D_CUDA(cudaGetLastError());
}

#ifdef S3D_KNN_STATS
extern "C" void s3d_debug_knn_stats(unsigned long long* out, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_knn_stats, sizeof(unsigned long long) * 8);
  if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(g_knn_stats, z, sizeof z); }
}
#endif

void run_expand_cov(Workspace& ws, double* cov_out) {
  uint32_t max_n = 0;
  for (uint32_t s = 0; s < ws.n_slots; ++s) max_n = std::max(max_n, ws.h_n[s]);
  if (max_n == 0) return;
  dim3 grid((max_n + 255) / 256, ws.n_slots);
  expand_cov_kernel<<<grid, 256, 0, ws.stream>>>(ws.slots.as<SlotInfo>(), cov_out);
  ++ws.launches;
}

// ---- stage API: exact 1-NN of T*query in a reference slot (thread per query) --------------------------------------
__global__ void __launch_bounds__(256) nn_stage_kernel(const SlotInfo* __restrict__ slots, uint32_t ref_slot, uint32_t qry_slot,
                                                       const float* __restrict__ T, uint32_t* __restrict__ nn_index, float* __restrict__ nn_dist2) {
  const SlotInfo& rs = slots[ref_slot];
  const SlotInfo& qs = slots[qry_slot];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= qs.n_raw) return;
  const GridView g = make_grid_view(rs);
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
  nn_stage_kernel<<<(nq + 255) / 256, 256, 0, ws.stream>>>(ws.slots.as<SlotInfo>(), ref_slot, qry_slot, T16_dev, nn_index, nn_dist2);
  ++ws.launches;
  S3D_CUDA(cudaGetLastError());
}

}  // namespace s3d
