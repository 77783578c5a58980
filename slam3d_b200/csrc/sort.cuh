// sort.cuh — batched (segmented), stable LSD radix sort of (uint32 key, uint32 value) pairs, 8-bit digits.
//
// Used twice per cloud: (1) voxel keys of pcl::VoxelGrid (PointCloudSensor.cpp:195-198) and
// (2) Morton keys of the NN grid.  Stability is part of the parity contract: equal voxel keys keep ascending
// input order, which fixes the float summation order of each centroid (A.1 step 8).
//
// One pass = three launches over a host-built tile table (tiles never straddle slots):
//   hist    per-tile 256-bin digit histogram (shared-memory atomics)
//   scan    one warp per (slot, digit): digit-major / tile-minor exclusive scan -> per-(tile,digit) output offsets
//   scatter per-warp match-any ranking -> stable positions, direct scatter
// Memory-bound streaming kernels: 2 x 8 B x n per pass (L2-resident for one scan, HBM for map-sized clouds).
#pragma once

#include "common.cuh"

namespace s3d {

enum CountSel { kCountRaw = 0, kCountPts = 1 };

__device__ __forceinline__ uint32_t slot_count(const SlotInfo& s, int which) { return which == kCountRaw ? s.n_raw : s.n_pts; }

static __global__ void __launch_bounds__(kSortThreads) sort_hist_kernel(const SlotInfo* __restrict__ slots, TileMap tm,
                                                                  const uint32_t* __restrict__ keys, uint32_t* __restrict__ hist,
                                                                  uint32_t* __restrict__ totals, int shift, int which) {
  __shared__ uint32_t sh[256];
  const uint32_t t = blockIdx.x;
  const uint32_t slot = tm.tile_slot[t], first = tm.tile_first[t];
  const uint32_t n = slot_count(slots[slot], which);
  if (first >= n) return;
  sh[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t* k = keys + slots[slot].off;
#pragma unroll
  for (int j = 0; j < kSortTile / kSortThreads; ++j) {
    const uint32_t e = first + j * kSortThreads + threadIdx.x;
    if (e < n) atomicAdd(&sh[(k[e] >> shift) & 255u], 1u);
  }
  __syncthreads();
  const uint32_t c = sh[threadIdx.x];
  hist[(size_t)t * 256 + threadIdx.x] = c;
  if (c) atomicAdd(&totals[(size_t)slot * 256 + threadIdx.x], c);  // per-slot digit totals (integer: order independent)
}

// grid (32, n_slots), 8 warps per CTA, one warp per digit: the warp first sums the totals of all smaller digits, then
// walks the tiles 32 at a time with a shuffle scan, turning hist[tile][digit] into the first output position of that
// (tile, digit).  (A first version with one CTA per slot walking all tiles serially took 65 us per pass on a 2M-point cloud.)
static __global__ void __launch_bounds__(256) sort_scan_kernel(const SlotInfo* __restrict__ slots, const uint32_t* __restrict__ slot_tile_begin,
                                                         uint32_t* __restrict__ hist, const uint32_t* __restrict__ totals, int which) {
  const uint32_t slot = blockIdx.y;
  const uint32_t n = slot_count(slots[slot], which);
  const uint32_t ntiles = (n + kSortTile - 1) / kSortTile;
  if (ntiles == 0) return;
  uint32_t* h = hist + (size_t)slot_tile_begin[slot] * 256;
  const int lane = threadIdx.x & 31;
  const int d = blockIdx.x * 8 + (threadIdx.x >> 5);
  const uint32_t* tot = totals + (size_t)slot * 256;
  uint32_t base = 0;
  for (int i = lane; i < d; i += 32) base += tot[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) base += __shfl_xor_sync(0xFFFFFFFFu, base, o);
  uint32_t running = base;
  for (uint32_t b = 0; b < ntiles; b += 32) {
    const uint32_t t = b + lane;
    const uint32_t v = t < ntiles ? h[(size_t)t * 256 + d] : 0u;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += u; }
    if (t < ntiles) h[(size_t)t * 256 + d] = running + incl - v;
    running += __shfl_sync(0xFFFFFFFFu, incl, 31);
  }
}

// vals_in == nullptr: the payload of element e is e (index inside the slot).
static __global__ void __launch_bounds__(kSortThreads) sort_scatter_kernel(const SlotInfo* __restrict__ slots, TileMap tm,
                                                                     const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                                     uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                                                                     const uint32_t* __restrict__ offsets, int shift, int which) {
  __shared__ uint32_t wcount[8][256];
  const uint32_t t = blockIdx.x;
  const uint32_t slot = tm.tile_slot[t], first = tm.tile_first[t];
  const uint32_t n = slot_count(slots[slot], which);
  if (first >= n) return;
  const uint32_t off = slots[slot].off;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 8 * 256; i += kSortThreads) (&wcount[0][0])[i] = 0;
  __syncthreads();
  uint32_t key[8], val[8];
  bool ok[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const uint32_t e = first + w * 256 + r * 32 + lane;
    ok[r] = e < n;
    key[r] = ok[r] ? keys_in[off + e] : 0u;
    val[r] = ok[r] ? (vals_in ? vals_in[off + e] : e) : 0u;
    if (ok[r]) atomicAdd(&wcount[w][(key[r] >> shift) & 255u], 1u);
  }
  __syncthreads();
  {  // wcount[w][d] <- first output position of warp w's elements with digit d
    const int d = threadIdx.x;
    uint32_t running = offsets[(size_t)t * 256 + d];
#pragma unroll
    for (int i = 0; i < 8; ++i) { const uint32_t c = wcount[i][d]; wcount[i][d] = running; running += c; }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const uint32_t digit = ok[r] ? ((key[r] >> shift) & 255u) : (256u + lane);  // invalid lanes never match a digit
    const uint32_t peers = __match_any_sync(0xFFFFFFFFu, digit);
    const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
    uint32_t pos = 0;
    if (ok[r]) pos = wcount[w][digit] + rank;
    __syncwarp();
    if (ok[r] && rank == 0) wcount[w][digit] += __popc(peers);
    __syncwarp();
    if (ok[r]) { keys_out[off + pos] = key[r]; vals_out[off + pos] = val[r]; }
  }
}

// Sorts bits [0, 8*passes) of keys.  Buffers ping-pong; the result is in (keys[passes & 1], vals[passes & 1]).
// `totals`: passes * n_slots * 256 counters (zeroed here).
inline void radix_sort_segmented(cudaStream_t st, const SlotInfo* slots, uint32_t n_slots, const TileMap& tm,
                                 const uint32_t* slot_tile_begin, uint32_t* keys[2], uint32_t* vals[2], uint32_t* hist, uint32_t* totals,
                                 int passes, int which, uint64_t* launch_counter) {
  if (tm.n_tiles == 0) return;
  cudaMemsetAsync(totals, 0, sizeof(uint32_t) * 256 * size_t(n_slots) * passes, st);
  for (int p = 0; p < passes; ++p) {
    const int in = p & 1, out = in ^ 1;
    uint32_t* tot = totals + size_t(p) * n_slots * 256;
    sort_hist_kernel<<<tm.n_tiles, kSortThreads, 0, st>>>(slots, tm, keys[in], hist, tot, 8 * p, which);
    sort_scan_kernel<<<dim3(32, n_slots), 256, 0, st>>>(slots, slot_tile_begin, hist, tot, which);
    sort_scatter_kernel<<<tm.n_tiles, kSortThreads, 0, st>>>(slots, tm, keys[in], p == 0 ? nullptr : vals[in], keys[out], vals[out], hist,
                                                             8 * p, which);
    *launch_counter += 3;
  }
}

}  // namespace s3d
