// common.cuh — shared device helpers and batch descriptors of the sm_100a scan-matching kernels.
//
// Data model (DESIGN.md "HBM layout"): a launch processes a BATCH of registrations.  Pair p owns two cloud
// "slots": slot 2p = slam3d source scan (PCL target, the FIXED cloud B), slot 2p+1 = slam3d target scan (PCL
// source, the MOVING cloud A; PointCloudSensor.cpp:68-69 swap).  All per-point arrays are concatenations over
// slots; slot s owns the index range [off[s], off[s] + cap[s]) in every array, where off/cap come from the raw
// input sizes (known on the host), while the live counts (after voxel filtering) exist only on the device.
// Kernels are therefore launched on upper-bound grids and read their true extent from device memory, so a
// whole align() runs without a host round trip until the per-iteration convergence flag is read.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace s3d {

constexpr int kSortTile = 2048;      // elements per CTA in the radix sort / segment kernels (256 thr x 8)
constexpr int kSortThreads = 256;
constexpr int kMaxLevels = 10;       // Morton bits per axis (30-bit keys)
constexpr uint32_t kInvalidKey = 0xFFFFFFFFu;
constexpr uint32_t kNoIndex = 0xFFFFFFFFu;
constexpr int kMaxKShared = 200;     // up to here the per-thread heap of the kNN kernel lives in shared memory (k KB per CTA); beyond: global memory
constexpr int kMaxK = 4096;          // sanity limit of correspondence_randomness (heap arena: 8 k bytes per point)

struct HashEntry;

// Per-slot description, device resident.
struct SlotInfo {
  // --- raw input -----------------------------------------------------------------------------------
  const float4* raw;   // device pointer to this slot's n_raw input points (user memory or the H2D staging buffer)
  uint32_t off;        // first index of this slot in all per-point arrays
  uint32_t n_raw;      // raw points
  // --- voxel filter (A.1) --------------------------------------------------------------------------
  float    bb_min[3], bb_max[3];  // over finite points
  uint32_t n_finite;
  float    inv_leaf;
  int32_t  min_b[3];
  uint32_t mul1, mul2;            // divb_mul[1], divb_mul[2]
  int32_t  overflow;              // PCL int32 guard fired: output = input
  uint32_t n_pts;                 // live points of the working (filtered) cloud
  // --- NN grid (my design, not PCL) ------------------------------------------------------------------
  float    g_min[3];              // grid origin = bbox min of the working cloud
  float    g_max[3];
  float    inv_h0;                // 1 / finest cell size
  float    h0;
  int32_t  nlev;                  // levels 0..nlev-1; level L has 2^(nlev-L) cells per axis
  float    margin;                // absolute slack for float cell assignment
  uint32_t hash_off;              // first entry of this slot's table in the hash arena
  uint32_t hash_cap;              // entries
  uint32_t n_cells;               // occupied cells over all levels
  // --- where the prepared cloud lives (batch workspace arrays + off, or the buffers of an s3d_prepared_cloud) ---------
  const float4*    gpts;          // Morton-sorted points, .w = original index bits
  const double4*   normals;       // unit normal of the regularised covariance per sorted point
  const HashEntry* table;         // hash arena base; this slot's entries start at hash_off
};

// Host-built tile table: tile t of a launch belongs to slot tile_slot[t] and covers elements
// [tile_first[t], tile_first[t] + kSortTile) of that slot.
struct TileMap {
  const uint32_t* tile_slot;
  const uint32_t* tile_first;
  uint32_t n_tiles;
};

// ---- float helpers with PCL's operation order; intrinsics are never contracted into FMAs -------------
__device__ __forceinline__ float dist2_pcl(float ax, float ay, float az, float bx, float by, float bz) {
  // flann::L2_Simple<float>: ((dx*dx) + dy*dy) + dz*dz      
  const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// pcl::transformPointCloud float se3 form: x*c0 + (y*c1 + (z*c2 + c3)).   T column-major.
__device__ __forceinline__ float3 transform_se3(const float* T, float x, float y, float z) {
  float3 o;
  o.x = __fadd_rn(__fmul_rn(x, T[0]), __fadd_rn(__fmul_rn(y, T[4]), __fadd_rn(__fmul_rn(z, T[8]), T[12])));
  o.y = __fadd_rn(__fmul_rn(x, T[1]), __fadd_rn(__fmul_rn(y, T[5]), __fadd_rn(__fmul_rn(z, T[9]), T[13])));
  o.z = __fadd_rn(__fmul_rn(x, T[2]), __fadd_rn(__fmul_rn(y, T[6]), __fadd_rn(__fmul_rn(z, T[10]), T[14])));
  return o;
}

// Eigen Matrix4f * Vector4f (w = 1): ((c0*x + c1*y) + c2*z) + c3.   
__device__ __forceinline__ float3 transform_mv(const float* T, float x, float y, float z) {
  float3 o;
  o.x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[0], x), __fmul_rn(T[4], y)), __fmul_rn(T[8], z)), T[12]);
  o.y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[1], x), __fmul_rn(T[5], y)), __fmul_rn(T[9], z)), T[13]);
  o.z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[2], x), __fmul_rn(T[6], y)), __fmul_rn(T[10], z)), T[14]);
  return o;
}

__device__ __forceinline__ bool finite3(float x, float y, float z) { return isfinite(x) && isfinite(y) && isfinite(z); }

// order-preserving float <-> uint mapping for atomicMin/atomicMax
__device__ __forceinline__ uint32_t float_to_ordered(float f) { uint32_t u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float ordered_to_float(uint32_t u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u); }

// ---- Morton codes (10 bits per axis) ---------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t spread3(uint32_t v) {
  v &= 0x3FFu;
  v = (v | (v << 16)) & 0x030000FFu;
  v = (v | (v << 8)) & 0x0300F00Fu;
  v = (v | (v << 4)) & 0x030C30C3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}
__host__ __device__ __forceinline__ uint32_t morton3(uint32_t x, uint32_t y, uint32_t z) { return spread3(x) | (spread3(y) << 1) | (spread3(z) << 2); }

// ---- hash of (level, cell) -> [begin, end) in the Morton-sorted point array -----------------------------
struct __align__(16) HashEntry { uint32_t key; uint32_t level; uint32_t begin; uint32_t end; };  // level == 0xFFFFFFFF: empty

__device__ __forceinline__ uint32_t hash_slot(uint32_t key, uint32_t level, uint32_t cap) {
  uint32_t h = key * 0x9E3779B1u + level * 0x85EBCA6Bu;
  h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12;
  return __umulhi(h, cap);  // uniform in [0, cap)
}

// cell with Morton key `key` at `level` (the caller has checked that it lies inside the grid); false when empty
__device__ __forceinline__ bool cell_range_key(const HashEntry* __restrict__ table, uint32_t cap, uint32_t key, int level, uint32_t& begin, uint32_t& end) {
  uint32_t s = hash_slot(key, (uint32_t)level, cap);
  for (;;) {
    const uint4 e = __ldg(reinterpret_cast<const uint4*>(table + s));
    if (e.y == 0xFFFFFFFFu) return false;
    if (e.x == key && e.y == (uint32_t)level) { begin = e.z; end = e.w; return true; }
    if (++s == cap) s = 0;
  }
}

// cell (cx,cy,cz) at `level`; returns false when the cell is empty or outside the grid
__device__ __forceinline__ bool cell_range(const HashEntry* __restrict__ table, uint32_t cap, int nlev, int level,
                                           int cx, int cy, int cz, uint32_t& begin, uint32_t& end) {
  const int dim = 1 << (nlev - level);
  if ((unsigned)cx >= (unsigned)dim || (unsigned)cy >= (unsigned)dim || (unsigned)cz >= (unsigned)dim) return false;
  const uint32_t key = morton3(cx, cy, cz);
  uint32_t s = hash_slot(key, level, cap);
  for (;;) {
    const uint4 e = __ldg(reinterpret_cast<const uint4*>(table + s));
    if (e.y == 0xFFFFFFFFu) return false;
    if (e.x == key && e.y == (uint32_t)level) { begin = e.z; end = e.w; return true; }
    if (++s == cap) s = 0;
  }
}

// Continuous cell coordinate of a point at level 0 (same expression in key generation and in every query).
__device__ __forceinline__ float grid_coord(float v, float origin, float inv_h0) { return __fmul_rn(__fsub_rn(v, origin), inv_h0); }

// Look-back state of a workspace (buffers are owned by the caller).  `aux`: kSortPasses * n_slots * 256 digit totals followed by
// kSortPasses tickets; `status`: n_tiles * 256 words, zero when (re)allocated; `epoch`: bumped once per pass, never reused.
struct SortState {
  uint32_t* aux;
  uint64_t* status;
  size_t status_words;  // capacity of `status`
  uint32_t* epoch;
  int32_t* flags;
};

}  // namespace s3d
