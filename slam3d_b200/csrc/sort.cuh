// sort.cuh — batched (segmented), stable LSD radix sort of (uint32 key, uint32 value) pairs, 8-bit digits.
//
// Used twice per cloud: (1) voxel keys of pcl::VoxelGrid (PointCloudSensor.cpp:195-198) and
// (2) Morton keys of the NN grid.  Stability is part of the parity contract: equal voxel keys keep ascending
// input order, which fixes the float summation order of each centroid (A.1 step 8).
//
// One sort = 1 + passes launches over a host-built tile table (tiles never straddle slots):
//   digits  (sort_digits_kernel, or fused into the key generation of the caller) per-slot digit totals of ALL passes in one read
//   pass    (sort_pass_kernel) one sweep per digit: a CTA ranks its tile (per-warp match-any ranking, stable), publishes the
//           tile's digit counts and gets its output offsets by decoupled look-back over the earlier tiles of its slot — no
//           separate histogram / scan kernels and no second read of the keys.  Tiles take a ticket (atomic counter) so that a
//           tile only ever waits for tiles that are already running; status words carry an epoch, so nothing is cleared
//           between passes or sorts.  A bounded spin turns a scheduler fault into kErrSortStall instead of a hang.
// Memory-bound streaming kernels: 2 x 8 B x n per pass (L2-resident for one scan, HBM for map-sized clouds).
// (Round 1 ran three kernels per pass — histogram, scan, scatter: 12 launches and ~40 us per pass on 2M points.)
#pragma once

#include "common.cuh"

namespace s3d {

enum CountSel { kCountRaw = 0, kCountPts = 1 };

__device__ __forceinline__ uint32_t slot_count(const SlotInfo& s, int which) { return which == kCountRaw ? s.n_raw : s.n_pts; }

constexpr int kSortPasses = 4;
constexpr uint32_t kLbAggregate = 1u, kLbInclusive = 2u;   // status word: epoch << 32 | flag << 30 | count (slots hold < 2^30 points)
constexpr uint32_t kLbValueMask = 0x3FFFFFFFu;
constexpr uint32_t kLbSpinLimit = 1u << 24;                 // ~10 s of polling: a fault, not a wait

__device__ __forceinline__ uint64_t lb_load(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void lb_store(uint64_t* p, uint64_t v) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }

// Adds one tile's keys to the digit counters of all passes (shared memory: sh[kSortPasses][256], zeroed by the caller).
__device__ __forceinline__ void count_digits(uint32_t (*sh)[256], uint32_t key) {
#pragma unroll
  for (int p = 0; p < kSortPasses; ++p) atomicAdd(&sh[p][(key >> (8 * p)) & 255u], 1u);
}
// totals: [kSortPasses][n_slots][256]
__device__ __forceinline__ void flush_digits(uint32_t (*sh)[256], uint32_t* __restrict__ totals, uint32_t n_slots, uint32_t slot) {
#pragma unroll
  for (int p = 0; p < kSortPasses; ++p) {
    const uint32_t c = sh[p][threadIdx.x];
    if (c) atomicAdd(&totals[(size_t(p) * n_slots + slot) * 256 + threadIdx.x], c);  // integer: order independent
  }
}

static __global__ void __launch_bounds__(kSortThreads) sort_digits_kernel(const SlotInfo* __restrict__ slots, TileMap tm, uint32_t n_slots,
                                                                    const uint32_t* __restrict__ keys, uint32_t* __restrict__ totals, int which) {
  __shared__ uint32_t sh[kSortPasses][256];
  const uint32_t t = blockIdx.x;
  const uint32_t slot = tm.tile_slot[t], first = tm.tile_first[t];
  const uint32_t n = slot_count(slots[slot], which);
  if (first >= n) return;
#pragma unroll
  for (int p = 0; p < kSortPasses; ++p) sh[p][threadIdx.x] = 0;
  __syncthreads();
  const uint32_t* k = keys + slots[slot].off;
#pragma unroll
  for (int j = 0; j < kSortTile / kSortThreads; ++j) {
    const uint32_t e = first + j * kSortThreads + threadIdx.x;
    if (e < n) count_digits(sh, k[e]);
  }
  __syncthreads();
  flush_digits(sh, totals, n_slots, slot);
}

// vals_in == nullptr: the payload of element e is e (index inside the slot).
// totals: this pass's [n_slots][256]; status: [n_tiles][256] look-back words; ticket: this pass's tile counter (zero at launch).
static __global__ void __launch_bounds__(kSortThreads) sort_pass_kernel(const SlotInfo* __restrict__ slots, TileMap tm,
                                                                  const uint32_t* __restrict__ slot_tile_begin,
                                                                  const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                                  uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                                                                  const uint32_t* __restrict__ totals, uint64_t* __restrict__ status,
                                                                  uint32_t* __restrict__ ticket, int32_t* __restrict__ flags, uint32_t epoch, int shift, int which) {
  __shared__ uint32_t wcount[8][256];
  __shared__ uint32_t s_first[256];
  __shared__ uint32_t s_wsum[8];
  __shared__ uint32_t s_tile;
  if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
  for (int i = threadIdx.x; i < 8 * 256; i += kSortThreads) (&wcount[0][0])[i] = 0;
  __syncthreads();
  const uint32_t t = s_tile;
  const uint32_t slot = tm.tile_slot[t], first = tm.tile_first[t];
  const uint32_t n = slot_count(slots[slot], which);
  if (first >= n || (which == kCountRaw && slots[slot].overflow)) return;  // dead tiles are a suffix of their slot (or all of it): nobody looks back at them
  const uint32_t off = slots[slot].off;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
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
  {
    const int d = threadIdx.x;
    // wcount[i][d] <- position of warp i's first element with digit d inside the tile's digit-d run; run = the tile's count
    uint32_t run = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { const uint32_t c = wcount[i][d]; wcount[i][d] = run; run += c; }
    // publish, then look back over the earlier tiles of this slot
    const uint64_t tag = uint64_t(epoch) << 32;
    uint64_t* mine = status + size_t(t) * 256 + d;
    uint32_t excl = 0;
    if (t == slot_tile_begin[slot]) {
      lb_store(mine, tag | (uint64_t(kLbInclusive) << 30) | run);
    } else {
      lb_store(mine, tag | (uint64_t(kLbAggregate) << 30) | run);
      for (uint32_t q = t - 1;; --q) {
        const uint64_t* theirs = status + size_t(q) * 256 + d;
        uint64_t v = lb_load(theirs);
        uint32_t spin = 0;
        while (uint32_t(v >> 32) != epoch) {
          if (++spin > kLbSpinLimit) { atomicOr(&flags[0], kErrSortStall); v = tag | (uint64_t(kLbInclusive) << 30); break; }
          __nanosleep(20);
          v = lb_load(theirs);
        }
        excl += uint32_t(v) & kLbValueMask;
        if ((uint32_t(v) >> 30) == kLbInclusive) break;
      }
      lb_store(mine, tag | (uint64_t(kLbInclusive) << 30) | (excl + run));
    }
    // first position of digit d in the slot: exclusive scan of the slot's digit totals
    const uint32_t tot = totals[size_t(slot) * 256 + d];
    uint32_t incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += u; }
    if (lane == 31) s_wsum[w] = incl;
    __syncthreads();
    uint32_t base = incl - tot;
    for (int i = 0; i < w; ++i) base += s_wsum[i];
    s_first[d] = base + excl;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const uint32_t digit = ok[r] ? ((key[r] >> shift) & 255u) : (256u + lane);  // invalid lanes never match a digit
    const uint32_t peers = __match_any_sync(0xFFFFFFFFu, digit);
    const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
    uint32_t pos = 0;
    if (ok[r]) pos = s_first[digit] + wcount[w][digit] + rank;
    __syncwarp();
    if (ok[r] && rank == 0) wcount[w][digit] += __popc(peers);
    __syncwarp();
    if (ok[r]) { keys_out[off + pos] = key[r]; vals_out[off + pos] = val[r]; }
  }
}

inline size_t sort_aux_bytes(uint32_t n_slots) { return sizeof(uint32_t) * (size_t(kSortPasses) * 256 * n_slots + kSortPasses); }

inline void sort_clear_aux(cudaStream_t st, const SortState& ss, uint32_t n_slots) { cudaMemsetAsync(ss.aux, 0, sort_aux_bytes(n_slots), st); }

// Sorts all 32 key bits.  Buffers ping-pong; the result is back in (keys[0], vals[0]).  digits_done: the caller cleared `aux`
// (sort_clear_aux) and counted the digits while it generated the keys (count_digits / flush_digits).
inline void radix_sort_segmented(cudaStream_t st, const SlotInfo* slots, uint32_t n_slots, const TileMap& tm, const uint32_t* slot_tile_begin,
                                 uint32_t* keys[2], uint32_t* vals[2], const SortState& ss, int which, bool digits_done, uint64_t* launch_counter) {
  if (tm.n_tiles == 0) return;
  if (!digits_done) {
    sort_clear_aux(st, ss, n_slots);
    sort_digits_kernel<<<tm.n_tiles, kSortThreads, 0, st>>>(slots, tm, n_slots, keys[0], ss.aux, which);
    ++*launch_counter;
  }
  uint32_t* tickets = ss.aux + size_t(kSortPasses) * 256 * n_slots;
  for (int p = 0; p < kSortPasses; ++p) {
    const int in = p & 1, out = in ^ 1;
    if (++*ss.epoch == 0) {  // wrapped: stale words could match again
      cudaMemsetAsync(ss.status, 0, sizeof(uint64_t) * ss.status_words, st);
      *ss.epoch = 1;
    }
    sort_pass_kernel<<<tm.n_tiles, kSortThreads, 0, st>>>(slots, tm, slot_tile_begin, keys[in], p == 0 ? nullptr : vals[in], keys[out], vals[out],
                                                          ss.aux + size_t(p) * n_slots * 256, ss.status, tickets + p, ss.flags, *ss.epoch, 8 * p, which);
    ++*launch_counter;
  }
}

}  // namespace s3d
