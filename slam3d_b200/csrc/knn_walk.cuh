// knn_walk.cuh — the per-thread exact k-NN walk of knn_cov_kernel (knn.cu): heap, cell scan, level loop.
//
// Kept in a header so that the search logic — what is scanned, what is pruned, when a result is certified — can also be
// compiled for the host by the CPU test-suite (tests/hostsearch.cpp, with tests/cuda_host_shim.h standing in for the device
// intrinsics) and checked against brute force without a GPU.  The product only ever runs it on the device.
#pragma once

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

// d-ary max-heap of 64-bit keys in shared memory, element j of thread t at h.at(j); children of node i are
// D i + 1 .. D i + D.  With D = 4 a 20-element heap is two levels deep instead of five, and the child loads of a level are
// independent (their latencies overlap), so a sift-down is two short rounds instead of four or five dependent ones
// (B200, 64 pairs per step: see profiles/r01h_summary.md).  Any heap order gives the same k-NN set and the heapsort below
// the same ascending order, so the arity cannot change a result.
#ifndef S3D_KNN_HEAP_ARITY
#define S3D_KNN_HEAP_ARITY 4
#endif
constexpr int kHeapD = S3D_KNN_HEAP_ARITY;
__device__ __forceinline__ int heap_last_parent(int n) { return n >= 2 ? (n - 2) / kHeapD : -1; }

// Where a thread's heap lives.  SmemHeap: element j of thread t at smem[j * kKnnThreads + t] (bank-conflict free, compile-time
// stride) — k up to kMaxKShared.  GlobalHeap: the same layout in a global-memory arena with a run-time stride (the number of
// threads of the launch), for the k beyond what shared memory holds; slower per access, but the walk and its results are the same.
struct SmemHeap {
  uint64_t* base;
  __device__ __forceinline__ uint64_t& at(int j) const { return base[j * kKnnThreads]; }
};
struct GlobalHeap {
  uint64_t* base;
  size_t stride;
  __device__ __forceinline__ uint64_t& at(int j) const { return base[(size_t)j * stride]; }
};

// puts x at node i (whose subtrees are heaps) and restores the heap below it; i = 0 replaces the root
template <typename Heap>
__device__ __forceinline__ void heap_sift_down(const Heap& h, int n, uint64_t x, int i = 0) {
  for (;;) {
    const int c0 = kHeapD * i + 1;
    if (c0 >= n) break;
    int c = c0;
    uint64_t hc = h.at(c0);
#pragma unroll
    for (int j = 1; j < kHeapD; ++j) {
      const int cj = min(c0 + j, n - 1);  // past the end: the last element again (a real child, compared twice)
      const uint64_t hj = h.at(cj);
      if (hj > hc) { hc = hj; c = cj; }
    }
    if (hc <= x) break;
    h.at(i) = hc;
    i = c;
  }
  h.at(i) = x;
}

// Replace-the-root for the one shape that matters — a 4-ary heap of exactly 20 keys (PCL's default k): root, nodes 1..4, leaves
// 5..19 (node 4 has three children).  Same result as heap_sift_down(h, 20, x), without its loop, bounds and clamps (the generic
// sift-down was 30 % of the kernel's warp instructions at 6.6 of 32 lanes, profiles/r02_summary.md); returns the new root.
template <typename Heap>
__device__ __forceinline__ uint64_t heap_replace_root_20(const Heap& h, uint64_t x) {
  const uint64_t a1 = h.at(1), a2 = h.at(2), a3 = h.at(3), a4 = h.at(4);
  int c = 1;
  uint64_t hc = a1;
  if (a2 > hc) { hc = a2; c = 2; }
  if (a3 > hc) { hc = a3; c = 3; }
  if (a4 > hc) { hc = a4; c = 4; }
  if (hc <= x) { h.at(0) = x; return x; }
  h.at(0) = hc;
  const int c0 = 4 * c + 1;  // 5, 9, 13 or 17
  const uint64_t b1 = h.at(c0), b2 = h.at(c0 + 1), b3 = h.at(c0 + 2), b4 = c < 4 ? h.at(c0 + 3) : 0ull;  // node 20 does not exist
  int d = c0;
  uint64_t hd = b1;
  if (b2 > hd) { hd = b2; d = c0 + 1; }
  if (b3 > hd) { hd = b3; d = c0 + 2; }
  if (b4 > hd) { hd = b4; d = c0 + 3; }
  if (hd <= x) { h.at(c) = x; } else { h.at(c) = hd; h.at(d) = x; }
  return hc;
}

// ---- per-thread walk ---------------------------------------------------------------------------------------------
// One query, one thread: scans the 27-block around the query at level L and widens by doubling the cell size until the
// k-th distance is certified by the block's coverage radius (nn_search.cuh).  Candidates are (d2, original index)
// packed into one 64-bit key — d2 >= +0, so unsigned key order is the lexicographic (d2, idx) order of the parity
// contract — and the k best live in a per-thread max-heap in shared memory (bank-conflict free: element j of thread t
// at [j][t]).  A cell is skipped when the lower bound of its distance exceeds the current k-th best.  `bound`: inclusive
// admission bound carried over from an earlier scan (KMAX: none).  Returns the number of neighbours in the heap.
constexpr uint64_t KMAX = 0xFFFFFFFFFFFFFFFFull;
#ifndef S3D_KNN_BATCH
#define S3D_KNN_BATCH 4
#endif

// (Round 2 also tried admission for the whole batch first and the heap insertions afterwards, one per turn of a loop, so that the
// lanes with something to insert do it together — same results; 9.70 ms against 9.26 for batches of 4, 8.95 for batches of 8: the
// extra selects and spills cost what the better lane use saves.  Not kept.)
// scans the points [begin, end) of one cell for the query qv: fill phase (append, heapify once when the k-th candidate
// arrives), then replace-the-root insertions.  S3D_KNN_BATCH points are fetched before the first of them is examined, so
// their load latencies overlap (B200, 32 pairs per step: 4.41 / 4.30 / 4.06 ms for batches of 1 / 2 / 4).
template <typename Heap>
__device__ __forceinline__ void knn_scan_range(const GridView& g, const float4 qv, uint32_t begin, uint32_t end, uint64_t bound, const Heap& h, int kk,
                                               int& cnt, uint64_t& tau, float& tau_d2) {
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
        if (ck <= bound) {
          h.at(cnt) = ck; ++cnt; KSTAT(5, 1);
          if (cnt == kk) {
            for (int i = heap_last_parent(kk); i >= 0; --i) heap_sift_down(h, kk, h.at(i), i);
            tau = h.at(0); tau_d2 = __uint_as_float((uint32_t)(tau >> 32));
          }
        }
      } else if (ck < tau) {
        KSTAT(6, 1);
        if (kHeapD == 4 && kk == 20) tau = heap_replace_root_20(h, ck);
        else { heap_sift_down(h, kk, ck); tau = h.at(0); }
        tau_d2 = __uint_as_float((uint32_t)(tau >> 32));
      }
    }
  }
}

// (Tried and dropped, B200: noting the occupied cells of the 27-block in a first lockstep pass and letting every lane scan its
// noted cells back to back in a second pass — the variant that speeds up the 1-NN walk of nn_search.cuh by 4 % — costs the kNN
// kernel 27 %: while the heap is filling nothing can be pruned, so all 27 cells are probed, and the list takes shared memory.)
template <typename Heap>
__device__ __forceinline__ int thread_walk(const GridView& g, const float4 qv, float ux, float uy, float uz, int L, uint64_t bound, const Heap& h, int kk) {
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
S3D_CELL_LOOP_PRAGMA
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
      S3D_SCAN_TRACE(L, i, end - begin);
      knn_scan_range(g, qv, begin, end, bound, h, kk, cnt, tau, tau_d2);
    }
    const bool full = cnt == kk;
    if ((full && tau_d2 <= g2) || top) break;
    if (full) bound = tau;
  }
  return cnt;
}

// (Tried and dropped in round 2, B200: a two-pass walk without a heap — pass 1 histograms the squared distances of the points
// within the block's certified radius into 32 bins of [0, g2], the bin in which the count reaches k bounds the k-th distance,
// pass 2 collects everything up to that bin and drops the surplus.  Bit-exact on the host and on the device, 0.14 % of the queries
// fall back to the heap walk — and 2.1x SLOWER: it executes as many warp instructions as the heap walk (477 M against 429 M per
// launch: 129 + ~50 points scanned per query instead of 87, no pruning while counting), at 93 registers and 40 KB of shared
// memory only 18 % of the warp slots are occupied, and with no heap work between the point loads 34 % of the stall samples sit on
// the first use of a loaded point.  The heap is also what hides the load latency.  profiles/r02_summary.md.)

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

}  // namespace s3d
