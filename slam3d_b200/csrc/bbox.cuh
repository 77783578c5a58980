// bbox.cuh — bounding box over the finite points of every cloud of a batch, with the per-cloud parameter step that needs it folded
// into the same launch: the CTA that finishes last (a ticket) runs `finish` for every slot.  Used by the voxel filter (raw cloud ->
// bb_*, then A.1 steps 1/3/4 of pcl::VoxelGrid) and by the NN grid (working cloud -> g_*, then the cell size / level count).
// (Round 1 ran the parameter step as its own one-CTA kernel: 2.5 us + a launch gap, twice per batch.)
#pragma once

#include "common.cuh"
#include "sort.cuh"

namespace s3d {

// done: a counter that is zero at launch; the last CTA leaves it zero again.
template <int kWhich, typename Finish>
__global__ void __launch_bounds__(kSortThreads) bbox_kernel(SlotInfo* __restrict__ slots, TileMap tm, const float4* __restrict__ work, uint32_t n_slots,
                                                             uint32_t* __restrict__ done, Finish finish) {
  const uint32_t t = blockIdx.x;
  const uint32_t slot = tm.tile_slot[t], first = tm.tile_first[t];
  SlotInfo& si = slots[slot];
  const uint32_t n = slot_count(si, kWhich);
  __shared__ float smn[8][3], smx[8][3];
  __shared__ uint32_t scnt[8];
  __shared__ bool s_last;
  if (first < n) {
    const float4* p = kWhich == kCountRaw ? si.raw : work + si.off;
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    uint32_t cnt = 0;
#pragma unroll
    for (int j = 0; j < kSortTile / kSortThreads; ++j) {
      const uint32_t e = first + j * kSortThreads + threadIdx.x;
      if (e < n) {
        const float4 v = __ldg(p + e);  // pointer from the slot table: a plain load would be a generic LD
        if (finite3(v.x, v.y, v.z)) {
          ++cnt;
          mn[0] = fminf(mn[0], v.x); mn[1] = fminf(mn[1], v.y); mn[2] = fminf(mn[2], v.z);
          mx[0] = fmaxf(mx[0], v.x); mx[1] = fmaxf(mx[1], v.y); mx[2] = fmaxf(mx[2], v.z);
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        mn[a] = fminf(mn[a], __shfl_xor_sync(0xFFFFFFFFu, mn[a], o));
        mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xFFFFFFFFu, mx[a], o));
      }
      cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
    }
    // one set of atomics per CTA (8192 warps hammering 7 addresses cost 40 us on a 2M-point cloud)
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { for (int a = 0; a < 3; ++a) { smn[w][a] = mn[a]; smx[w][a] = mx[a]; } scnt[w] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int i = 1; i < 8; ++i) {
        for (int a = 0; a < 3; ++a) { mn[a] = fminf(mn[a], smn[i][a]); mx[a] = fmaxf(mx[a], smx[i][a]); }
        cnt += scnt[i];
      }
      if (cnt) {
        uint32_t* dmin = reinterpret_cast<uint32_t*>(kWhich == kCountRaw ? si.bb_min : si.g_min);
        uint32_t* dmax = reinterpret_cast<uint32_t*>(kWhich == kCountRaw ? si.bb_max : si.g_max);
#pragma unroll
        for (int a = 0; a < 3; ++a) { atomicMin(&dmin[a], float_to_ordered(mn[a])); atomicMax(&dmax[a], float_to_ordered(mx[a])); }
        if (kWhich == kCountRaw) atomicAdd(&si.n_finite, cnt);
      }
    }
  }
  if (threadIdx.x == 0) {
    __threadfence();  // this CTA's atomics before its ticket
    s_last = atomicAdd(done, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();  // every other CTA's atomics are visible; drops this SM's L1 lines of the slot table
  for (uint32_t s = threadIdx.x; s < n_slots; s += kSortThreads) finish(slots[s]);
  if (threadIdx.x == 0) *done = 0;
}

}  // namespace s3d
