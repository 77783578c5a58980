// nn_search.cuh — exact nearest-neighbour queries on the multi-resolution voxel hash (grid.cu).
//
// Parity contract: the result equals brute force with flann::L2_Simple<float> distances
// ((dx*dx + dy*dy) + dz*dz, unfused) and lexicographic (d2, original index) order, i.e. exact search with ties
// going to the lowest index.  Exactness argument: a scan of the 3x3x3 block of level-L cells around the query
// sees every point closer than  g = h_L * (1 + o) - slack, o = distance (in cells) from the query to the nearest
// face of its own cell; so a best candidate with d2 <= g^2 is the global optimum.  Otherwise the cell size doubles.
// The top level has 2 cells per axis, so its block is the whole cloud and the loop always terminates exactly.
#pragma once

#include "common.cuh"

namespace s3d {

struct GridView {
  const HashEntry* table;
  const float4* pts;   // Morton-sorted points of the slot, .w = original index bits
  uint32_t cap, n;
  int nlev;
  float ox, oy, oz, inv_h0, h0, margin;
};

__device__ __forceinline__ GridView make_grid_view(const SlotInfo& si) {
  GridView g;
  g.table = si.table + si.hash_off; g.pts = si.gpts; g.cap = si.hash_cap; g.n = si.n_pts; g.nlev = si.nlev;
  g.ox = si.g_min[0]; g.oy = si.g_min[1]; g.oz = si.g_min[2]; g.inv_h0 = si.inv_h0; g.h0 = si.h0; g.margin = si.margin;
  return g;
}

__device__ __forceinline__ float clamp_coord(float u) { return fminf(fmaxf(u, -1.0e6f), 1.0e6f); }  // NaN -> -1e6

// Cell of the query at level L, its fractional position inside that cell, and the squared radius the 27-block
// around it is guaranteed to cover.
__device__ __forceinline__ float block_guarantee2(const GridView& g, float ux, float uy, float uz, int L, int& cx, int& cy, int& cz,
                                                  float& ax, float& ay, float& az) {
  const float s = 1.0f / (float)(1 << L);  // exact power of two
  const float vx = ux * s, vy = uy * s, vz = uz * s;
  const float fx = floorf(vx), fy = floorf(vy), fz = floorf(vz);
  cx = (int)fx; cy = (int)fy; cz = (int)fz;
  ax = vx - fx; ay = vy - fy; az = vz - fz;
  const float o = fminf(fminf(fminf(ax, 1.f - ax), fminf(ay, 1.f - ay)), fminf(az, 1.f - az));
  const float r = g.h0 * (float)(1 << L) * (1.f + o) * 0.9999f - g.margin;
  return r > 0.f ? r * r * 0.999999f : 0.f;
}
__device__ __forceinline__ float block_guarantee2(const GridView& g, float ux, float uy, float uz, int L, int& cx, int& cy, int& cz) {
  float ax, ay, az;
  return block_guarantee2(g, ux, uy, uz, L, cx, cy, cz, ax, ay, az);
}

// lb2: lower bound of the squared distance from the query to every reference point OTHER than the winner
// (second-best scanned candidate, pruned cells, and the coverage radius of the last block).
struct NNResult { float d2; uint32_t idx; uint32_t pos; float lb2; };

__device__ __forceinline__ bool cand_less(float d2, uint32_t idx, float bd2, uint32_t bidx) { return d2 < bd2 || (d2 == bd2 && idx < bidx); }

// Visit order of the 27 cells of a block: own cell, 6 face neighbours, 12 edge neighbours, 8 corners (index = dx + 3 dy + 9 dz).
// Near cells first tightens the pruning bound early, so most edge/corner cells are skipped without a hash probe.
// {13, 12,14,10,16,4,22, 9,11,15,17,3,5,21,23,1,7,19,25, 0,2,6,8,18,20,24,26} packed 5 bits per entry, 12 entries per word
// (registers, not constant memory: the index differs per lane).
__device__ __forceinline__ int cell_order(int i) {
  const unsigned long long w = i < 12 ? 0x1c5eb4d8905398dull : (i < 24 ? 0x920c2066670dea5ull : 0x6b14ull);
  const int j = i < 12 ? i : (i < 24 ? i - 12 : i - 24);
  return (int)((w >> (5 * j)) & 31ull);
}

// thread-per-query scan of the 27-block at level L.  All lanes of a warp walk the cells in the same (near-first) order,
// so the hash probes of a cell are issued together; a flattened per-lane iterator that lets lanes drift apart was
// measured 2x slower because every step then waits for some lane's probe.  A cell is skipped when the lower bound of
// its distance (gap along each axis, shrunk by the float slack) already exceeds the best candidate — with a warm start
// that leaves 1-4 of the 27 cells.  (ax,ay,az) = position of the query inside its cell in [0,1); prune = false at the
// top level, where the block is anchored at cell 0 instead of the query.
// the 27-slot loop of the walks: rolled by default (measured in round 2: see profiles/r02_summary.md); -DS3D_CELL_UNROLL=27 lets
// the compiler fold the slot decoding and the per-axis selects of every slot into constants
#ifndef S3D_CELL_UNROLL
#define S3D_CELL_UNROLL 1
#endif
#define S3D_PRAGMA_(x) _Pragma(#x)
#define S3D_PRAGMA(x) S3D_PRAGMA_(x)
#define S3D_CELL_LOOP_PRAGMA S3D_PRAGMA(unroll S3D_CELL_UNROLL)
constexpr int kNNGatherCap = 8;  // noted cells per thread in the gathering variant (8 bytes each)

// host-side work statistics (tests/hostsearch.cpp defines it to record which cell slot scanned how many points); nothing on the device
#ifndef S3D_SCAN_TRACE
#define S3D_SCAN_TRACE(level, slot, count)
#endif

__device__ __forceinline__ void scan_range(const GridView& g, uint32_t begin, uint32_t end, float qx, float qy, float qz, NNResult& best) {
  // (fetching 2 or 4 points before examining the first was measured SLOWER for this 1-NN walk on B200: gicp_iter 4.55 ->
  // 5.55 ms per 32 pairs — most cells are left after one or two points; the kNN kernel, which visits every point, gains 8 %)
  for (uint32_t p = begin; p < end; ++p) {
    const float4 v = __ldg(g.pts + p);
    const float d2 = dist2_pcl(qx, qy, qz, v.x, v.y, v.z);
    const uint32_t id = __float_as_uint(v.w);
    if (cand_less(d2, id, best.d2, best.idx)) { best.lb2 = fminf(best.lb2, best.d2); best.d2 = d2; best.idx = id; best.pos = p; }
    else if (id != best.idx) best.lb2 = fminf(best.lb2, d2);  // the winner itself is met again when a level is rescanned
  }
}

// kGather: the lockstep pass scans only the own cell and notes the other surviving cells in `lst` (entry j of a thread at
// lst[j * blockDim.x]); a second pass then lets every lane scan ITS noted cells back to back, in the same near-first order and
// with the same pruning rule evaluated at scan time.  The winner is the exact nearest neighbour either way.  lb2 stays a
// valid lower bound for every other point but need not be the single-pass value: when more than kNNGatherCap cells survive
// the first pass (own cell empty), the surplus is scanned before the noted ones, so a cell may be scanned where the single
// pass would have pruned it, or the other way round (tests/test_hostsearch.py checks both walks on the host).
template <bool kGather>
__device__ __forceinline__ void scan_block(const GridView& g, int L, int cx, int cy, int cz, float ax, float ay, float az, bool prune,
                                           float qx, float qy, float qz, NNResult& best, uint2* lst = nullptr) {
  const int dim = 1 << (g.nlev - L);
  const float hl = g.h0 * (float)(1 << L) * 0.9999f;
  // Morton bits of the three cell coordinates per axis, spread once per block instead of once per cell
  const uint32_t sx0 = spread3((uint32_t)(cx - 1)), sx1 = spread3((uint32_t)cx), sx2 = spread3((uint32_t)(cx + 1));
  const uint32_t sy0 = spread3((uint32_t)(cy - 1)) << 1, sy1 = spread3((uint32_t)cy) << 1, sy2 = spread3((uint32_t)(cy + 1)) << 1;
  const uint32_t sz0 = spread3((uint32_t)(cz - 1)) << 2, sz1 = spread3((uint32_t)cz) << 2, sz2 = spread3((uint32_t)(cz + 1)) << 2;
  // squared axis gaps to the lower neighbour / upper neighbour cell (0 for the own cell), shrunk by the float slack
  float glx = 0.f, gux = 0.f, gly = 0.f, guy = 0.f, glz = 0.f, guz = 0.f;
  if (prune) {
    const float a0 = fmaxf(ax * hl - g.margin, 0.f), a1 = fmaxf((1.f - ax) * hl - g.margin, 0.f);
    const float b0 = fmaxf(ay * hl - g.margin, 0.f), b1 = fmaxf((1.f - ay) * hl - g.margin, 0.f);
    const float c0 = fmaxf(az * hl - g.margin, 0.f), c1 = fmaxf((1.f - az) * hl - g.margin, 0.f);
    glx = a0 * a0; gux = a1 * a1; gly = b0 * b0; guy = b1 * b1; glz = c0 * c0; guz = c1 * c1;
  }
  uint32_t found = 0;
  int n_list = 0;
S3D_CELL_LOOP_PRAGMA
  for (int i = 0; i < 27; ++i) {
    const int c = cell_order(i);
    const int dx = c % 3, dy = (c / 3) % 3, dz = c / 9;
    const int ix = cx + dx - 1, iy = cy + dy - 1, iz = cz + dz - 1;
    if ((unsigned)ix >= (unsigned)dim || (unsigned)iy >= (unsigned)dim || (unsigned)iz >= (unsigned)dim) continue;
    if (prune) {
      const float cell_lb = ((dx == 0 ? glx : (dx == 1 ? 0.f : gux)) + (dy == 0 ? gly : (dy == 1 ? 0.f : guy)) + (dz == 0 ? glz : (dz == 1 ? 0.f : guz))) * 0.99999f;
      if (cell_lb > best.d2) { best.lb2 = fminf(best.lb2, cell_lb); continue; }
    }
    const uint32_t key = (dx == 0 ? sx0 : (dx == 1 ? sx1 : sx2)) | (dy == 0 ? sy0 : (dy == 1 ? sy1 : sy2)) | (dz == 0 ? sz0 : (dz == 1 ? sz1 : sz2));
    uint32_t begin, end;
    if (!cell_range_key(g.table, g.cap, key, L, begin, end)) continue;
    if (!kGather || i == 0 || n_list == kNNGatherCap) { S3D_SCAN_TRACE(L, i, end - begin); scan_range(g, begin, end, qx, qy, qz, best); }
    else { lst[n_list * blockDim.x] = make_uint2(begin, end); found |= 1u << i; ++n_list; }
  }
  if (kGather) {
    for (int j = 0; found; ++j) {
      const int i = __ffs(found) - 1;
      found &= found - 1;
      const uint2 be = lst[j * blockDim.x];
      if (prune) {
        const int c = cell_order(i);
        const int dx = c % 3, dy = (c / 3) % 3, dz = c / 9;
        const float cell_lb = ((dx == 0 ? glx : (dx == 1 ? 0.f : gux)) + (dy == 0 ? gly : (dy == 1 ? 0.f : guy)) + (dz == 0 ? glz : (dz == 1 ? 0.f : guz))) * 0.99999f;
        if (cell_lb > best.d2) { best.lb2 = fminf(best.lb2, cell_lb); continue; }
      }
      S3D_SCAN_TRACE(L, i, be.y - be.x);
      scan_range(g, be.x, be.y, qx, qy, qz, best);
    }
  }
}

// Exact 1-NN.  cutoff2: distances above it are of no interest (search may stop once the block covers that radius);
// hint_pos / hint2_pos: sorted positions of candidates whose distance bounds the search (the previous correspondence of this
// query, the fresh result of the neighbouring query) or kNoIndex.  Hints only steer the level and the pruning; the result is
// the exact nearest neighbour either way.
template <bool kGather = false>
__device__ __forceinline__ NNResult nn_search(const GridView& g, float qx, float qy, float qz, float cutoff2, uint32_t hint_pos,
                                              uint32_t hint2_pos = kNoIndex, uint2* lst = nullptr) {
  NNResult best{INFINITY, kNoIndex, kNoIndex, INFINITY};
  if (g.n == 0 || g.cap == 0) return best;
  const float ux = clamp_coord(grid_coord(qx, g.ox, g.inv_h0));
  const float uy = clamp_coord(grid_coord(qy, g.oy, g.inv_h0));
  const float uz = clamp_coord(grid_coord(qz, g.oz, g.inv_h0));
  int L = 0;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const uint32_t hp = h == 0 ? hint_pos : hint2_pos;
    if (hp < g.n && hp != best.pos) {
      const float4 v = __ldg(g.pts + hp);
      const float d2 = dist2_pcl(qx, qy, qz, v.x, v.y, v.z);
      const uint32_t id = __float_as_uint(v.w);
      if (d2 == d2) {  // not NaN
        if (cand_less(d2, id, best.d2, best.idx)) { best.lb2 = fminf(best.lb2, best.d2); best.d2 = d2; best.idx = id; best.pos = hp; }
        else best.lb2 = fminf(best.lb2, d2);
      }
    }
  }
  if (best.pos != kNoIndex) {
    const float need = fminf(best.d2, cutoff2);
    int cx, cy, cz;
    while (L < g.nlev - 1 && block_guarantee2(g, ux, uy, uz, L, cx, cy, cz) < need) ++L;
  }
  for (;; ++L) {
    int cx, cy, cz;
    float ax, ay, az;
    const float g2 = block_guarantee2(g, ux, uy, uz, L, cx, cy, cz, ax, ay, az);
    const bool top = L >= g.nlev - 1;
    if (top) cx = cy = cz = 0;  // 2 cells per axis: the block around cell 0 is the whole cloud, wherever the query is
    scan_block<kGather>(g, L, cx, cy, cz, ax, ay, az, !top, qx, qy, qz, best, lst);
    if (top) break;                                   // the block was the whole cloud
    if (best.d2 <= g2 || g2 >= cutoff2) { best.lb2 = fminf(best.lb2, g2); break; }  // everything outside the block is farther than sqrt(g2)
  }
  return best;
}

}  // namespace s3d
