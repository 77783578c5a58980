// grid.cu — the nearest-neighbour acceleration structure: a multi-resolution voxel hash over Morton-sorted points.
//
// Replaces the two FLANN kd-trees PCL builds inside icp.align() (PointCloudSensor.cpp:70).
// Only the RESULTS of the searches are part of the parity contract (exact, ties -> lowest index); the structure is
// this project's own design:
//   * every working cloud is sorted by a 30-bit Morton code of its finest-level cell (cell size h0 ~ 3 x voxel leaf);
//   * because of the Morton order, the points of a cell at ANY level L (cell size h0 * 2^L) are one contiguous range,
//     so one sort serves all levels; a per-slot open-addressing hash maps (level, cell) -> [begin, end);
//   * LiDAR density falls off as 1/r^2, so queries pick the level that fits the local density and widen by doubling
//     the cell size (knn.cu), which bounds the work of the sparse far field where a fixed grid would need ~25 rings.
// gpts[] holds the sorted points as float4 with the ORIGINAL index bit-cast into .w, so one 16-byte coalesced load
// yields both the coordinates and the tie-break key.
#include "internal.h"
#include "bbox.cuh"
#include "sort.cuh"

namespace s3d {

// run per slot by the last CTA of the bbox launch (bbox.cuh)
struct GridParams {
  float leaf_hint;
  __device__ void operator()(SlotInfo& si) const {
  float ext = 0.f, amax = 0.f;
  for (int a = 0; a < 3; ++a) {
    si.g_min[a] = ordered_to_float(reinterpret_cast<uint32_t&>(si.g_min[a]));
    si.g_max[a] = ordered_to_float(reinterpret_cast<uint32_t&>(si.g_max[a]));
    ext = fmaxf(ext, si.g_max[a] - si.g_min[a]);
    amax = fmaxf(amax, fmaxf(fabsf(si.g_min[a]), fabsf(si.g_max[a])));
  }
  if (si.n_pts == 0 || !(ext >= 0.f) || !isfinite(ext)) {  // empty cloud or no finite point
    si.h0 = 1.f; si.inv_h0 = 1.f; si.nlev = 1; si.margin = 0.f;
    for (int a = 0; a < 3; ++a) si.g_min[a] = 0.f;
    return;
  }
  const float span = ext * 1.001f + 1e-6f;
#ifndef S3D_H0_FACTOR
#define S3D_H0_FACTOR 3.0f
#endif
  float h0 = leaf_hint > 0.f ? S3D_H0_FACTOR * leaf_hint : span / 1024.f;  // finest cell = 3 voxel leaves (swept 2..4 on B200)
  int nlev = 1;
  while (nlev < kMaxLevels && h0 * (float)(1 << nlev) <= span) ++nlev;
  if (h0 * (float)(1 << nlev) <= span) h0 = span / (float)(1 << nlev);
  si.h0 = h0; si.inv_h0 = 1.0f / h0; si.nlev = nlev;
  si.margin = 1e-4f * h0 + 16.f * 1.1920929e-7f * (amax + ext);
  }
};

__device__ __forceinline__ int cell_of(float u, int dim) {
  // u >= 0 inside the bbox; NaN -> 0; clamp keeps non-finite / out-of-box inputs inside the key range
  const float f = floorf(fminf(fmaxf(u, 0.f), (float)(dim - 1)));
  return (int)f;
}

// + the digit totals of all four sort passes (sort.cuh)
__global__ void __launch_bounds__(kSortThreads) grid_keys_kernel(const SlotInfo* __restrict__ slots, TileMap tm, uint32_t n_slots, const float4* __restrict__ work,
                                                                  uint32_t* __restrict__ keys, uint32_t* __restrict__ totals) {
  __shared__ uint32_t sh[kSortPasses][256];
  const uint32_t t = blockIdx.x;
  const uint32_t slot = tm.tile_slot[t], first = tm.tile_first[t];
  const SlotInfo& si = slots[slot];
  if (first >= si.n_pts) return;
#pragma unroll
  for (int p = 0; p < kSortPasses; ++p) sh[p][threadIdx.x] = 0;
  __syncthreads();
  const int dim = 1 << si.nlev;
#pragma unroll
  for (int j = 0; j < kSortTile / kSortThreads; ++j) {
    const uint32_t e = first + j * kSortThreads + threadIdx.x;
    if (e < si.n_pts) {
      const float4 v = work[si.off + e];
      const int cx = cell_of(grid_coord(v.x, si.g_min[0], si.inv_h0), dim);
      const int cy = cell_of(grid_coord(v.y, si.g_min[1], si.inv_h0), dim);
      const int cz = cell_of(grid_coord(v.z, si.g_min[2], si.inv_h0), dim);
      const uint32_t key = morton3(cx, cy, cz);
      keys[si.off + e] = key;
      count_digits(sh, key);
    }
  }
  __syncthreads();
  flush_digits(sh, totals, n_slots, slot);
}

__device__ __forceinline__ int levels_started(const uint32_t* __restrict__ k, uint32_t e, int nlev) {
  if (e == 0) return nlev;
  const uint32_t x = k[e] ^ k[e - 1];
  if (x == 0) return 0;
  const int l = (31 - __clz(x)) / 3 + 1;
  return l < nlev ? l : nlev;
}

// gather into Morton order (+ original index in .w) and count the occupied cells over all levels
__global__ void __launch_bounds__(kSortThreads) grid_gather_kernel(SlotInfo* __restrict__ slots, TileMap tm, const float4* __restrict__ work,
                                                                    const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                                                    float4* __restrict__ gpts) {
  const uint32_t t = blockIdx.x;
  const uint32_t slot = tm.tile_slot[t], first = tm.tile_first[t];
  SlotInfo& si = slots[slot];
  if (first >= si.n_pts) return;
  uint32_t cells = 0;
#pragma unroll
  for (int j = 0; j < kSortTile / kSortThreads; ++j) {
    const uint32_t e = first + j * kSortThreads + threadIdx.x;
    if (e < si.n_pts) {
      const uint32_t src = vals[si.off + e];
      float4 v = work[si.off + src];
      v.w = __uint_as_float(src);
      gpts[si.off + e] = v;
      cells += levels_started(keys + si.off, e, si.nlev);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cells += __shfl_xor_sync(0xFFFFFFFFu, cells, o);
  if ((threadIdx.x & 31) == 0 && cells) atomicAdd(&si.n_cells, cells);
}

// one thread: carve the hash arena (2 entries per occupied cell).  flags[3] always receives the number of entries the
// batch needs, so that an overflow (flags[0] & kErrHashArena) can be answered by one exact re-allocation and a re-run.
__global__ void hash_layout_kernel(SlotInfo* __restrict__ slots, uint32_t n_slots, uint32_t arena_cap, int32_t* __restrict__ flags) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  uint64_t running = 0;
  for (uint32_t s = 0; s < n_slots; ++s) {
    const uint64_t cap = 2ull * slots[s].n_cells + 8;
    slots[s].hash_off = (uint32_t)min(running, (uint64_t)0xFFFFFFFFu); slots[s].hash_cap = (uint32_t)cap;
    running += cap;
  }
  flags[3] = (int32_t)min(running, (uint64_t)0x7FFFFFFF);
  if (running > arena_cap) {
    atomicOr(&flags[0], kErrHashArena);
    for (uint32_t s = 0; s < n_slots; ++s) { slots[s].hash_off = 0; slots[s].hash_cap = 0; }  // every later kernel sees an empty grid
    running = 0;
  }
  flags[2] = (int32_t)running;
}

__global__ void hash_clear_kernel(HashEntry* __restrict__ table, const int32_t* __restrict__ flags) {
  const uint32_t used = (uint32_t)flags[2];
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < used; i += gridDim.x * blockDim.x)
    reinterpret_cast<uint4*>(table)[i] = make_uint4(0u, 0xFFFFFFFFu, 0u, 0u);
}

__global__ void __launch_bounds__(kSortThreads) hash_insert_kernel(const SlotInfo* __restrict__ slots, TileMap tm, const uint32_t* __restrict__ keys,
                                                                    HashEntry* __restrict__ table) {
  const uint32_t t = blockIdx.x;
  const uint32_t slot = tm.tile_slot[t], first = tm.tile_first[t];
  const SlotInfo& si = slots[slot];
  const uint32_t n = si.n_pts;
  if (first >= n || si.hash_cap == 0) return;
  const uint32_t* k = keys + si.off;
  HashEntry* tab = table + si.hash_off;
#pragma unroll 1
  for (int j = 0; j < kSortTile / kSortThreads; ++j) {
    const uint32_t e = first + j * kSortThreads + threadIdx.x;
    if (e >= n) continue;
    const int started = levels_started(k, e, si.nlev);
    const uint32_t key0 = k[e];
    uint32_t lo = e + 1;  // cells nest: the level-L cell ends where the level-(L-1) cell ended, or later
    for (int L = 0; L < started; ++L) {
      const uint32_t ck = key0 >> (3 * L);
      // end = first position whose level-L cell differs (keys ascending => cells ascending).  Cells are short (1.2 points at the
      // finest level, ~4x per level), so gallop from the last known member instead of bisecting [e, n): 2-4 dependent loads
      // instead of 17 per (cell, level) — the bisection made this kernel 80 % of the grid stage (profiles/r02_summary.md).
      uint32_t hi = n;
      for (uint32_t step = 1; lo + step <= n; step <<= 1) {
        const uint32_t p = lo + step - 1;
        if ((k[p] >> (3 * L)) > ck) { hi = p; break; }
        lo = p + 1;
      }
      while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if ((k[mid] >> (3 * L)) > ck) hi = mid; else lo = mid + 1; }
      uint32_t s = hash_slot(ck, (uint32_t)L, si.hash_cap);
      for (;;) {  // all inserted (level, cell) pairs are distinct: claim the first empty entry
        if (atomicCAS(&tab[s].level, 0xFFFFFFFFu, (uint32_t)L) == 0xFFFFFFFFu) { tab[s].key = ck; tab[s].begin = e; tab[s].end = lo; break; }
        if (++s == si.hash_cap) s = 0;
      }
    }
  }
}

void run_grid(Workspace& ws, float leaf_hint) {
  if (ws.n_tiles == 0) return;
  cudaStream_t st = ws.stream;
  StageTimer timer(ws, kStageGrid);
  SlotInfo* slots = ws.slots.as<SlotInfo>();
  TileMap tm{ws.tile_slot.as<uint32_t>(), ws.tile_first.as<uint32_t>(), ws.n_tiles};
  uint32_t* bbox_done = ws.flags.as<uint32_t>() + 10;  // zero between launches (bbox.cuh)
  bbox_kernel<kCountPts><<<ws.n_tiles, kSortThreads, 0, st>>>(slots, tm, ws.work.as<float4>(), ws.n_slots, bbox_done, GridParams{leaf_hint});
  uint32_t* keys[2] = {ws.keys0.as<uint32_t>(), ws.keys1.as<uint32_t>()};
  uint32_t* vals[2] = {ws.vals0.as<uint32_t>(), ws.vals1.as<uint32_t>()};
  const SortState ss = ws.sort_state();
  sort_clear_aux(st, ss, ws.n_slots);
  grid_keys_kernel<<<ws.n_tiles, kSortThreads, 0, st>>>(slots, tm, ws.n_slots, ws.work.as<float4>(), keys[0], ss.aux);
  ws.launches += 2;
  radix_sort_segmented(st, slots, ws.n_slots, tm, ws.slot_tile_begin.as<uint32_t>(), keys, vals, ss, kCountPts, /*digits_done=*/true, &ws.launches);
  grid_gather_kernel<<<ws.n_tiles, kSortThreads, 0, st>>>(slots, tm, ws.work.as<float4>(), keys[0], vals[0], ws.gpts.as<float4>());
  hash_layout_kernel<<<1, 32, 0, st>>>(slots, ws.n_slots, (uint32_t)std::min<size_t>(ws.hash_cap, 0xFFFFFFF0u), ws.flags.as<int32_t>());
  hash_clear_kernel<<<ws.n_sms * 4, 256, 0, st>>>(ws.hash.as<HashEntry>(), ws.flags.as<int32_t>());
  hash_insert_kernel<<<ws.n_tiles, kSortThreads, 0, st>>>(slots, tm, keys[0], ws.hash.as<HashEntry>());
  ws.launches += 4;
  S3D_CUDA(cudaGetLastError());
}

}  // namespace s3d
