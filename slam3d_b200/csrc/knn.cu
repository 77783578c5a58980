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
#include <cstdlib>

#include "internal.h"
#include "nn_search.cuh"

namespace s3d {

#ifndef S3D_KNN_THREADS
#define S3D_KNN_THREADS 128
#endif
#ifndef S3D_KNN_MINBLOCKS
#define S3D_KNN_MINBLOCKS 8
#endif
constexpr int kKnnThreads = S3D_KNN_THREADS;  // queries per CTA (one thread per query)

#ifdef S3D_KNN_STATS
__device__ unsigned long long g_knn_stats[8];  // queries, level scans, cells probed, cells pruned, candidates, pushes, sift-downs, start-level probes
#define KSTAT(i, v) atomicAdd(&g_knn_stats[i], (unsigned long long)(v))
#else
#define KSTAT(i, v)
#endif

// max-heap of 64-bit keys in shared memory, element j of thread t at h[j * kKnnThreads]
// puts x at node i (whose subtrees are heaps) and restores the heap below it; i = 0 replaces the root
__device__ __forceinline__ void heap_sift_down(uint64_t* h, int n, uint64_t x, int i = 0) {
  for (;;) {
    int c = 2 * i + 1;
    if (c >= n) break;
    uint64_t hc = h[c * kKnnThreads];
    if (c + 1 < n) { const uint64_t hr = h[(c + 1) * kKnnThreads]; if (hr > hc) { hc = hr; ++c; } }
    if (hc <= x) break;
    h[i * kKnnThreads] = hc;
    i = c;
  }
  h[i * kKnnThreads] = x;
}

// ---- per-thread walk ---------------------------------------------------------------------------------------------
// One query, one thread: scans the 27-block around the query at level L and widens by doubling the cell size until the
// k-th distance is certified by the block's coverage radius (nn_search.cuh).  Candidates are (d2, original index)
// packed into one 64-bit key — d2 >= +0, so unsigned key order is the lexicographic (d2, idx) order of the parity
// contract — and the k best live in a per-thread max-heap in shared memory (bank-conflict free: element j of thread t
// at [j][t]).  A cell is skipped when the lower bound of its distance exceeds the current k-th best.  `bound`: inclusive
// admission bound carried over from an earlier scan (KMAX: none).  Returns the number of neighbours in the heap.
constexpr uint64_t KMAX = 0xFFFFFFFFFFFFFFFFull;

__device__ __forceinline__ int thread_walk(const GridView& g, const float4 qv, float ux, float uy, float uz, int L, uint64_t bound, uint64_t* h, int kk) {
  int cnt = 0;
  for (;; ++L) {
    int cx, cy, cz;
    float ax, ay, az;
    const float g2 = block_guarantee2(g, ux, uy, uz, L, cx, cy, cz, ax, ay, az);
    const bool top = L >= g.nlev - 1;
    if (top) cx = cy = cz = 0;
    cnt = 0;
    KSTAT(1, 1);
    uint64_t tau = bound;                                   // current admission threshold (inclusive while not full)
    float tau_d2 = __uint_as_float((uint32_t)(tau >> 32));  // KMAX -> NaN bits: comparisons stay false, nothing is pruned
    const int dim = 1 << (g.nlev - L);
    const float hl = g.h0 * (float)(1 << L) * 0.9999f;
    // Morton bits of the three cell coordinates per axis, spread once per block instead of once per cell
    const uint32_t sx0 = spread3((uint32_t)(cx - 1)), sx1 = spread3((uint32_t)cx), sx2 = spread3((uint32_t)(cx + 1));
    const uint32_t sy0 = spread3((uint32_t)(cy - 1)) << 1, sy1 = spread3((uint32_t)cy) << 1, sy2 = spread3((uint32_t)(cy + 1)) << 1;
    const uint32_t sz0 = spread3((uint32_t)(cz - 1)) << 2, sz1 = spread3((uint32_t)cz) << 2, sz2 = spread3((uint32_t)(cz + 1)) << 2;
    float glx = 0.f, gux = 0.f, gly = 0.f, guy = 0.f, glz = 0.f, guz = 0.f;  // squared axis gaps to the lower / upper neighbour cell
    if (!top) {
      const float a0 = fmaxf(ax * hl - g.margin, 0.f), a1 = fmaxf((1.f - ax) * hl - g.margin, 0.f);
      const float b0 = fmaxf(ay * hl - g.margin, 0.f), b1 = fmaxf((1.f - ay) * hl - g.margin, 0.f);
      const float c0 = fmaxf(az * hl - g.margin, 0.f), c1 = fmaxf((1.f - az) * hl - g.margin, 0.f);
      glx = a0 * a0; gux = a1 * a1; gly = b0 * b0; guy = b1 * b1; glz = c0 * c0; guz = c1 * c1;
    }
#pragma unroll 1
    for (int i = 0; i < 27; ++i) {
      const int c = cell_order(i);  // own cell, faces, edges, corners
      const int dx = c % 3, dy = (c / 3) % 3, dz = c / 9;
      const int ix = cx + dx - 1, iy = cy + dy - 1, iz = cz + dz - 1;
      if ((unsigned)ix >= (unsigned)dim || (unsigned)iy >= (unsigned)dim || (unsigned)iz >= (unsigned)dim) continue;
      if (!top) {
        const float cell_lb = ((dx == 0 ? glx : (dx == 1 ? 0.f : gux)) + (dy == 0 ? gly : (dy == 1 ? 0.f : guy)) + (dz == 0 ? glz : (dz == 1 ? 0.f : guz))) * 0.99999f;
        if (cell_lb > tau_d2) { KSTAT(3, 1); continue; }  // the whole cell is farther than the k-th best / the bound
      }
      const uint32_t key = (dx == 0 ? sx0 : (dx == 1 ? sx1 : sx2)) | (dy == 0 ? sy0 : (dy == 1 ? sy1 : sy2)) | (dz == 0 ? sz0 : (dz == 1 ? sz1 : sz2));
      uint32_t begin, end;
      KSTAT(2, 1);
      if (!cell_range_key(g.table, g.cap, key, L, begin, end)) continue;
      KSTAT(4, end - begin);
#ifndef S3D_KNN_BATCH
#define S3D_KNN_BATCH 4
#endif
      // S3D_KNN_BATCH points are fetched before the first of them is examined, so their load latencies overlap
      // (B200, 32 pairs per step: 4.41 / 4.30 / 4.06 ms for batches of 1 / 2 / 4)
      for (uint32_t p0 = begin; p0 < end; p0 += S3D_KNN_BATCH) {
        float4 vb[S3D_KNN_BATCH];
#pragma unroll
        for (int u = 0; u < S3D_KNN_BATCH; ++u) vb[u] = __ldg(g.pts + min(p0 + u, end - 1));
#pragma unroll
        for (int u = 0; u < S3D_KNN_BATCH; ++u) {
          if (u > 0 && p0 + u >= end) break;
          const float4 v = vb[u];
          const float cd = dist2_pcl(qv.x, qv.y, qv.z, v.x, v.y, v.z);
          if (!(cd == cd)) continue;  // NaN never enters
          const uint64_t ck = ((uint64_t)__float_as_uint(cd) << 32) | (uint64_t)__float_as_uint(v.w);
          if (cnt < kk) {
            if (ck <= bound) {  // fill phase: append, and heapify once when the k-th candidate arrives
              h[cnt * kKnnThreads] = ck; ++cnt; KSTAT(5, 1);
              if (cnt == kk) {
                for (int i = kk / 2 - 1; i >= 0; --i) heap_sift_down(h, kk, h[i * kKnnThreads], i);
                tau = h[0]; tau_d2 = __uint_as_float((uint32_t)(tau >> 32));
              }
            }
          } else if (ck < tau) {
            heap_sift_down(h, kk, ck); KSTAT(6, 1);
            tau = h[0]; tau_d2 = __uint_as_float((uint32_t)(tau >> 32));
          }
        }
      }
    }
    const bool full = cnt == kk;
    if ((full && tau_d2 <= g2) || top) break;
    if (full) bound = tau;
  }
  return cnt;
}

// start level of a query: smallest L whose parent cell (level L+1) already holds >= 20 points (swept 8..40 on B200)
__device__ __forceinline__ int knn_start_level(const GridView& g, float ux, float uy, float uz) {
  const int c0x = (int)floorf(ux), c0y = (int)floorf(uy), c0z = (int)floorf(uz);
  for (int lv = 1; lv < g.nlev; ++lv) {
    uint32_t b, e;
    KSTAT(7, 1);
#ifndef S3D_KNN_START
#define S3D_KNN_START 20u
#endif
    if (cell_range(g.table, g.cap, g.nlev, lv, c0x >> lv, c0y >> lv, c0z >> lv, b, e) && (e - b) >= S3D_KNN_START) return lv - 1;
  }
  return g.nlev - 1;
}

// heapsort (ascending (d2, idx) = FLANN's sorted result order), moments exactly as PCL (float products accumulated in
// double, in neighbour order; A.3 step 2), smallest eigenvector; optional neighbour lists for the stage API
__device__ __forceinline__ void knn_finish(const SlotInfo& si, const float4* __restrict__ cloud, uint64_t* h, int cnt, int k, uint32_t r, uint32_t q_orig,
                                           double4* __restrict__ normals, uint32_t* __restrict__ knn_index, float* __restrict__ knn_dist2) {
  for (int m = cnt - 1; m > 0; --m) {
    const uint64_t last = h[m * kKnnThreads];
    h[m * kKnnThreads] = h[0];
    heap_sift_down(h, m, last);
  }
  double mean[3] = {0, 0, 0}, cov[6] = {0, 0, 0, 0, 0, 0};  // cov: 00,10,11,20,21,22
  for (int j = 0; j < cnt; ++j) {
    const uint64_t key = h[j * kKnnThreads];
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

// ---- kernel A: one thread per query, every lane walks its own cells (the round-1 kernel; kept as the A/B reference and
// for S3D_KNN_TILE=0) ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kKnnThreads, S3D_KNN_MINBLOCKS) knn_cov_kernel(const SlotInfo* __restrict__ slots, const float4* __restrict__ work,
                                                              double4* __restrict__ normals, int k, uint32_t* __restrict__ knn_index,
                                                              float* __restrict__ knn_dist2) {
  extern __shared__ uint64_t heap_smem[];  // k * kKnnThreads keys
  const SlotInfo& si = slots[blockIdx.y];
  const uint32_t n = si.n_pts;
  const uint32_t r = blockIdx.x * kKnnThreads + threadIdx.x;
  if (r >= n) return;
  const GridView g = make_grid_view(si);
  if (g.cap == 0) return;  // grid build overflowed its arena: the host re-runs the batch
  const int kk = k < (int)n ? k : (int)n;  // FLANN clamps k to the cloud size
  uint64_t* h = heap_smem + threadIdx.x;
  const float4 qv = g.pts[r];
  const float ux = clamp_coord(grid_coord(qv.x, g.ox, g.inv_h0));
  const float uy = clamp_coord(grid_coord(qv.y, g.oy, g.inv_h0));
  const float uz = clamp_coord(grid_coord(qv.z, g.oz, g.inv_h0));
  KSTAT(0, 1);
  const int cnt = thread_walk(g, qv, ux, uy, uz, knn_start_level(g, ux, uy, uz), KMAX, h, kk);
  if (cnt < kk) for (int i = cnt / 2 - 1; i >= 0; --i) heap_sift_down(h, cnt, h[i * kKnnThreads], i);  // top level ended before the list filled
  knn_finish(si, work + si.off, h, cnt, k, r, __float_as_uint(qv.w), normals, knn_index, knn_dist2);
}

// ---- kernel B: warp-cooperative tile search ------------------------------------------------------------------------
// A warp owns 32 Morton-adjacent queries and scans ONE candidate set for all of them:
//   * level l = median of the lanes' start levels, raised until the cell box of the queries plus kTileRings cells of margin
//     has at most kTileMaxCells cells;
//   * ring 0 = the cells of the queries' bounding box, ring r = the shell r cells further out.  The lanes probe the hash for
//     32 cells at a time; every occupied cell is then handled by the whole warp: each lane bounds its distance to the cell,
//     and the cell is skipped when no lane can still improve (one ballot); otherwise its points are fetched with one
//     coalesced 16-byte load per lane, staged in shared memory and read back as broadcasts, so all 32 lanes run the same
//     candidate loop (the per-thread walk of kernel A keeps 13 of 32 lanes busy: profiles/r01g_summary.md);
//   * a candidate that beats a lane's k-th best goes to a small per-lane buffer; when any buffer fills, all lanes merge
//     theirs into their heaps together, so the divergent heap work is paid once per flush, not once per candidate;
//   * after ring r a lane is done when its k-th distance is inside the scanned box (distance to the nearest box face that
//     is not a face of the grid itself, with the slack of nn_search.cuh); pruned cells cannot hold anything closer, so they
//     count as scanned.  Lanes still open after the last ring (3 % of the warps on a LiDAR scan) finish with the per-thread
//     walk one level up, with their k-th best as the admission bound.
// The result is the exact k-NN set either way: which cells are scanned only changes the work, never the answer.
#ifndef S3D_KNN_RINGS
#define S3D_KNN_RINGS 2
#endif
#ifndef S3D_KNN_TILE_MINBLOCKS
#define S3D_KNN_TILE_MINBLOCKS 6
#endif
constexpr int kTileRings = S3D_KNN_RINGS;
constexpr int kTileMaxCells = 512;
constexpr int kPushBuf = 8;  // buffered candidates per lane
constexpr unsigned kFullMask = 0xFFFFFFFFu;

__global__ void __launch_bounds__(kKnnThreads, S3D_KNN_TILE_MINBLOCKS) knn_cov_tile_kernel(const SlotInfo* __restrict__ slots, const float4* __restrict__ work,
                                                              double4* __restrict__ normals, int k, uint32_t* __restrict__ knn_index,
                                                              float* __restrict__ knn_dist2) {
  extern __shared__ uint64_t heap_smem[];  // k * kKnnThreads heap keys, then kPushBuf * kKnnThreads buffered keys
  __shared__ float4 stage_smem[kKnnThreads];
  const SlotInfo& si = slots[blockIdx.y];
  const uint32_t n = si.n_pts;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t r_raw = blockIdx.x * kKnnThreads + threadIdx.x;
  if (r_raw - lane >= n) return;  // the whole warp is past the end of the cloud
  const bool valid = r_raw < n;
  const uint32_t r = valid ? r_raw : n - 1;  // lanes past the end shadow the last query and write nothing
  const GridView g = make_grid_view(si);
  if (g.cap == 0) return;  // grid build overflowed its arena: the host re-runs the batch
  const int kk = k < (int)n ? k : (int)n;  // FLANN clamps k to the cloud size
  uint64_t* h = heap_smem + threadIdx.x;
  uint64_t* buf = heap_smem + (size_t)k * kKnnThreads + threadIdx.x;
  float4* st = stage_smem + (threadIdx.x & ~31u);

  const float4 qv = g.pts[r];
  const float ux = clamp_coord(grid_coord(qv.x, g.ox, g.inv_h0));
  const float uy = clamp_coord(grid_coord(qv.y, g.oy, g.inv_h0));
  const float uz = clamp_coord(grid_coord(qv.z, g.oz, g.inv_h0));
  KSTAT(0, valid ? 1 : 0);

  // ---- level and cell box of the warp ---------------------------------------------------------------------------------
  const int Lq = knn_start_level(g, ux, uy, uz);
  int l = g.nlev - 1;
  for (int v = 0; v < g.nlev - 1; ++v)
    if (__popc(__ballot_sync(kFullMask, Lq <= v)) >= 16) { l = v; break; }
  const int c0x = (int)floorf(ux), c0y = (int)floorf(uy), c0z = (int)floorf(uz);
  const int lo0x = __reduce_min_sync(kFullMask, c0x), hi0x = __reduce_max_sync(kFullMask, c0x);
  const int lo0y = __reduce_min_sync(kFullMask, c0y), hi0y = __reduce_max_sync(kFullMask, c0y);
  const int lo0z = __reduce_min_sync(kFullMask, c0z), hi0z = __reduce_max_sync(kFullMask, c0z);
  for (; l < g.nlev - 1; ++l) {
    const long long cells = (long long)((hi0x >> l) - (lo0x >> l) + 1 + 2 * kTileRings) * ((hi0y >> l) - (lo0y >> l) + 1 + 2 * kTileRings) *
                            ((hi0z >> l) - (lo0z >> l) + 1 + 2 * kTileRings);
    if (cells <= kTileMaxCells) break;
  }
  const int lox = lo0x >> l, loy = lo0y >> l, loz = lo0z >> l;
  const int sx = (hi0x >> l) - lox + 1, sy = (hi0y >> l) - loy + 1, sz = (hi0z >> l) - loz + 1;  // cells of the queries' box per axis
  const int dim = 1 << (g.nlev - l);
  const float sc = 1.0f / (float)(1 << l);  // exact power of two
  const float vx = ux * sc, vy = uy * sc, vz = uz * sc;
  const float hl = g.h0 * (float)(1 << l) * 0.9999f;

  for (int j = 0; j < kk; ++j) h[j * kKnnThreads] = KMAX;  // k sentinels: the root is a real key once k candidates came in
  uint64_t tau = KMAX;
  float tau_f = INFINITY;  // admission pre-filter: d2 <= tau_f (ties are settled on the 64-bit key when the buffer is merged)
  int bcnt = 0;
  bool done = !valid;

#define S3D_KNN_FLUSH()                                                   \
  do {                                                                    \
    for (int _i = 0; _i < bcnt; ++_i) {                                   \
      const uint64_t _ck = buf[_i * kKnnThreads];                         \
      if (_ck < tau) { heap_sift_down(h, kk, _ck); tau = h[0]; KSTAT(6, 1); } \
    }                                                                     \
    bcnt = 0;                                                             \
    if (tau != KMAX) tau_f = __uint_as_float((uint32_t)(tau >> 32));      \
  } while (0)

  for (int ring = 0; ring <= kTileRings; ++ring) {
    const int bx0 = lox - ring, by0 = loy - ring, bz0 = loz - ring;
    const int nx = sx + 2 * ring, ny = sy + 2 * ring, nz = sz + 2 * ring;
    const int total = nx * ny * nz;
    KSTAT(1, lane == 0 ? 1 : 0);
    for (int t0 = 0; t0 < total; t0 += 32) {
      const int t = t0 + (int)lane;
      bool found = false;
      uint32_t cb = 0, ce = 0;
      int ccx = 0, ccy = 0, ccz = 0;
      if (t < total) {
        const int ix = t % nx, iy = (t / nx) % ny, iz = t / (nx * ny);
        const bool inner = ring > 0 && ix > 0 && ix < nx - 1 && iy > 0 && iy < ny - 1 && iz > 0 && iz < nz - 1;  // scanned by an earlier ring
        if (!inner) {
          ccx = bx0 + ix; ccy = by0 + iy; ccz = bz0 + iz;
          KSTAT(2, 1);
          found = cell_range(g.table, g.cap, g.nlev, l, ccx, ccy, ccz, cb, ce);
        }
      }
      uint32_t mask = __ballot_sync(kFullMask, found);
      while (mask) {
        const int j = __ffs(mask) - 1;
        mask &= mask - 1;
        const uint32_t b = __shfl_sync(kFullMask, cb, j), e = __shfl_sync(kFullMask, ce, j);
        const float fx = (float)__shfl_sync(kFullMask, ccx, j), fy = (float)__shfl_sync(kFullMask, ccy, j), fz = (float)__shfl_sync(kFullMask, ccz, j);
        // lower bound of this lane's distance to the cell (axis gaps in cells, shrunk by the float slack as in scan_block)
        const float gx = fmaxf(fmaxf(fx - vx, vx - (fx + 1.f)), 0.f), gy = fmaxf(fmaxf(fy - vy, vy - (fy + 1.f)), 0.f), gz = fmaxf(fmaxf(fz - vz, vz - (fz + 1.f)), 0.f);
        const float ax = fmaxf(gx * hl - g.margin, 0.f), ay = fmaxf(gy * hl - g.margin, 0.f), az = fmaxf(gz * hl - g.margin, 0.f);
        const float cell_lb = (ax * ax + ay * ay + az * az) * 0.99999f;
        if (!__any_sync(kFullMask, !done && !(cell_lb > tau_f))) { KSTAT(3, lane == 0 ? 1 : 0); continue; }  // no lane can improve from this cell
        KSTAT(4, lane == 0 ? e - b : 0);
        for (uint32_t p0 = b; p0 < e; p0 += 32) {
          if (p0 + lane < e) st[lane] = __ldg(g.pts + p0 + lane);
          __syncwarp();
          const int m = (int)min(32u, e - p0);
          for (int j0 = 0; j0 < m; j0 += 4) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (j0 + u < m) {
                const float4 v = st[j0 + u];
                const float cd = dist2_pcl(qv.x, qv.y, qv.z, v.x, v.y, v.z);
                if (cd <= tau_f) {  // NaN never enters
                  buf[bcnt * kKnnThreads] = ((uint64_t)__float_as_uint(cd) << 32) | (uint64_t)__float_as_uint(v.w);
                  ++bcnt; KSTAT(5, 1);
                }
              }
            }
            if (__any_sync(kFullMask, bcnt > kPushBuf - 4)) S3D_KNN_FLUSH();
          }
          __syncwarp();
        }
      }
    }
    S3D_KNN_FLUSH();
    // coverage of the scanned box: distance to its nearest face that is not a face of the grid itself
    float d = INFINITY;
    if (bx0 > 0) d = fminf(d, vx - (float)bx0);
    if (by0 > 0) d = fminf(d, vy - (float)by0);
    if (bz0 > 0) d = fminf(d, vz - (float)bz0);
    if (bx0 + nx < dim) d = fminf(d, (float)(bx0 + nx) - vx);
    if (by0 + ny < dim) d = fminf(d, (float)(by0 + ny) - vy);
    if (bz0 + nz < dim) d = fminf(d, (float)(bz0 + nz) - vz);
    float g2 = INFINITY;
    if (d < INFINITY) { const float rr = hl * d - g.margin; g2 = rr > 0.f ? rr * rr * 0.999999f : 0.f; }
    if (tau != KMAX && tau_f <= g2) done = true;
    if (__all_sync(kFullMask, done)) break;
  }
#undef S3D_KNN_FLUSH
  int cnt = kk;
  if (!done) {  // still open after the last ring: per-thread walk one level up, bounded by the k-th best found so far
    cnt = thread_walk(g, qv, ux, uy, uz, min(l + 1, g.nlev - 1), tau, h, kk);
    if (cnt < kk) for (int i = cnt / 2 - 1; i >= 0; --i) heap_sift_down(h, cnt, h[i * kKnnThreads], i);
  }
  if (!valid) return;
  knn_finish(si, work + si.off, h, cnt, k, r, __float_as_uint(qv.w), normals, knn_index, knn_dist2);
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

static bool knn_tile_enabled() {
  static const int v = [] { const char* e = getenv("S3D_KNN_TILE"); return e ? atoi(e) : 1; }();  // 0: kernel A (A/B measurements)
  return v != 0;
}

void run_knn_covariances(Workspace& ws, int k, uint32_t* knn_index, float* knn_dist2) {
  if (ws.n_tiles == 0) return;
  uint32_t max_n = 0;
  for (uint32_t s = 0; s < ws.n_slots; ++s) max_n = std::max(max_n, ws.h_n[s]);
  StageTimer timer(ws, kStageKnn);
  dim3 grid((max_n + kKnnThreads - 1) / kKnnThreads, ws.n_slots);
  if (knn_tile_enabled()) {
    const size_t bytes = sizeof(uint64_t) * (size_t)(k + kPushBuf) * kKnnThreads;
    if (bytes > 40 * 1024) S3D_CUDA(cudaFuncSetAttribute(knn_cov_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    knn_cov_tile_kernel<<<grid, kKnnThreads, bytes, ws.stream>>>(ws.slots.as<SlotInfo>(), ws.work.as<float4>(), ws.normals.as<double4>(), k, knn_index, knn_dist2);
  } else {
    const size_t heap_bytes = sizeof(uint64_t) * k * kKnnThreads;
    if (heap_bytes > 48 * 1024) S3D_CUDA(cudaFuncSetAttribute(knn_cov_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)heap_bytes));
    knn_cov_kernel<<<grid, kKnnThreads, heap_bytes, ws.stream>>>(ws.slots.as<SlotInfo>(), ws.work.as<float4>(), ws.normals.as<double4>(), k, knn_index, knn_dist2);
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
