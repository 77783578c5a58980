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

// Look-back over the earlier tiles of a slot, one thread per counter: status words `stride` apart, from `q` down to `begin`
// (whose word is always inclusive).  Loads are issued kLbBatch at a time so that a long walk costs one L2 round trip per batch.
constexpr int kLbBatch = 4;
__device__ __forceinline__ uint32_t lb_walk(const uint64_t* __restrict__ status, long long q, long long begin, long long step, uint32_t epoch,
                                            int32_t* __restrict__ flags) {
  uint32_t excl = 0;
  bool done = false;
  while (!done) {
    uint64_t v[kLbBatch];
#pragma unroll
    for (int i = 0; i < kLbBatch; ++i) v[i] = (q - i * step >= begin) ? lb_load(status + size_t(q - i * step) * 256) : 0ull;
#pragma unroll
    for (int i = 0; i < kLbBatch; ++i) {
      if (done || q - i * step < begin) continue;
      uint32_t spin = 0;
      while (uint32_t(v[i] >> 32) != epoch) {
        if (++spin > kLbSpinLimit) { atomicOr(&flags[0], kErrSortStall); v[i] = (uint64_t(epoch) << 32) | (uint64_t(kLbInclusive) << 30); break; }
        __nanosleep(20);
        v[i] = lb_load(status + size_t(q - i * step) * 256);
      }
      excl += uint32_t(v[i]) & kLbValueMask;
      if ((uint32_t(v[i]) >> 30) == kLbInclusive) done = true;
    }
    q -= kLbBatch * step;
  }
  return excl;
}

// One digit pass.  A CTA sorts kSortMul consecutive table tiles of one slot (kSortBig keys): the launch has one CTA per table
// tile, and the CTAs whose tile is not the first of its group leave at once (no second tile table).
//   1. keys -> registers, per-warp digit counts (shared-memory atomics);
//   2. per digit: the tile's count is published (aggregate), its start inside the tile comes from a block scan;
//   3. stable ranks (warp order, then round order, then match-any inside the round) -> the keys are laid out in digit order in
//      shared memory; this needs no global information, so the earlier tiles get time to finish;
//   4. look-back -> global start of each digit run; the runs are written out with consecutive threads on consecutive
//      addresses (a direct scatter from registers costs one 32-byte sector per 4-byte key in the low-digit passes);
//   5. the payload takes the same route through the same shared buffer.
// vals_in == nullptr: the payload of element e is e (index inside the slot).
// totals: this pass's [n_slots][256]; status: [n_tiles][256] look-back words; ticket: this pass's tile counter (zero at launch).
#ifndef S3D_SORT_MUL
#define S3D_SORT_MUL 1
#endif
#ifndef S3D_SORT_MINB
#define S3D_SORT_MINB 6
#endif
#ifndef S3D_SORT_EARLY_VALS
#define S3D_SORT_EARLY_VALS 1
#endif
constexpr int kSortMul = S3D_SORT_MUL;
constexpr int kSortItems = kSortMul * kSortTile / kSortThreads;  // keys per thread
constexpr int kSortBig = kSortMul * kSortTile;
static __global__ void __launch_bounds__(kSortThreads, S3D_SORT_MINB) sort_pass_kernel(const SlotInfo* __restrict__ slots, TileMap tm,
                                                                  const uint32_t* __restrict__ slot_tile_begin,
                                                                  const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                                  uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                                                                  const uint32_t* __restrict__ totals, uint64_t* __restrict__ status,
                                                                  uint32_t* __restrict__ ticket, int32_t* __restrict__ flags, uint32_t epoch, int shift, int which) {
  __shared__ uint32_t wcount[8][256];
  __shared__ uint32_t s_start[256];   // first position of digit d inside the tile's sorted order
  __shared__ uint32_t s_delta[256];   // global position of the digit's first key minus s_start[d]
  __shared__ uint32_t s_buf[kSortBig];
  __shared__ uint32_t s_wsum[8];
  __shared__ uint32_t s_tile;
  if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
  for (int i = threadIdx.x; i < 8 * 256; i += kSortThreads) (&wcount[0][0])[i] = 0;
  __syncthreads();
  const uint32_t t = s_tile;
  const uint32_t slot = tm.tile_slot[t], first = tm.tile_first[t];
  const uint32_t n = slot_count(slots[slot], which);
  // dead tiles are a suffix of their slot (or all of it): nobody looks back at them
  if (first >= n || (first / kSortTile) % kSortMul != 0 || (which == kCountRaw && slots[slot].overflow)) return;
  const uint32_t off = slots[slot].off;
  const uint32_t live = min(n - first, (uint32_t)kSortBig);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  constexpr int kPerWarp = kSortBig / 8;
  uint32_t key[kSortItems];
  uint32_t lpos[kSortItems];
#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {
    const uint32_t i = w * kPerWarp + r * 32 + lane;
    key[r] = i < live ? keys_in[off + first + i] : 0u;
    if (i < live) atomicAdd(&wcount[w][(key[r] >> shift) & 255u], 1u);
  }
#if S3D_SORT_EARLY_VALS
  uint32_t val[kSortItems];  // in flight while the tile is ranked and the look-back runs
#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {
    const uint32_t i = w * kPerWarp + r * 32 + lane;
    val[r] = i < live ? (vals_in ? vals_in[off + first + i] : first + i) : 0u;
  }
#endif
  __syncthreads();
  const int d = threadIdx.x;
  uint32_t run = 0;
  {
    // wcount[i][d] <- position of warp i's first key with digit d inside the tile's digit-d run; run = the tile's count
#pragma unroll
    for (int i = 0; i < 8; ++i) { const uint32_t c = wcount[i][d]; wcount[i][d] = run; run += c; }
    const uint64_t tag = uint64_t(epoch) << 32;
    const bool head = t == slot_tile_begin[slot];
    lb_store(status + size_t(t) * 256 + d, tag | (uint64_t(head ? kLbInclusive : kLbAggregate) << 30) | run);
    uint32_t incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += u; }
    if (lane == 31) s_wsum[w] = incl;
    __syncthreads();
    uint32_t start = incl - run;
    for (int i = 0; i < w; ++i) start += s_wsum[i];
    s_start[d] = start;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {
    const uint32_t i = w * kPerWarp + r * 32 + lane;
    const bool ok = i < live;
    const uint32_t digit = ok ? ((key[r] >> shift) & 255u) : (256u + lane);  // invalid lanes never match a digit
    const uint32_t peers = __match_any_sync(0xFFFFFFFFu, digit);
    const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
    uint32_t pos = 0;
    if (ok) pos = s_start[digit] + wcount[w][digit] + rank;
    __syncwarp();
    if (ok && rank == 0) wcount[w][digit] += __popc(peers);
    __syncwarp();
    lpos[r] = pos;
    if (ok) s_buf[pos] = key[r];
  }
  {
    // global start of the digit run: the slot's digit base (exclusive scan of its totals) + the earlier tiles' counts (look-back)
    const uint32_t tot = totals[size_t(slot) * 256 + d];
    uint32_t incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += u; }
    __syncthreads();  // s_buf complete; s_wsum free again
    if (lane == 31) s_wsum[w] = incl;
    uint32_t excl = 0;
    if (t != slot_tile_begin[slot]) {
      excl = lb_walk(status + d, (long long)t - kSortMul, (long long)slot_tile_begin[slot], kSortMul, epoch, flags);
      lb_store(status + size_t(t) * 256 + d, (uint64_t(epoch) << 32) | (uint64_t(kLbInclusive) << 30) | (excl + run));
    }
    __syncthreads();
    uint32_t base = incl - tot;
    for (int i = 0; i < w; ++i) base += s_wsum[i];
    s_delta[d] = base + excl - s_start[d];
  }
  __syncthreads();
  uint32_t gpos[kSortItems];
#pragma unroll
  for (int j = 0; j < kSortItems; ++j) {
    const uint32_t i = j * kSortThreads + threadIdx.x;
    if (i < live) {
      const uint32_t kk = s_buf[i];
      gpos[j] = off + s_delta[(kk >> shift) & 255u] + i;
      keys_out[gpos[j]] = kk;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {
    const uint32_t i = w * kPerWarp + r * 32 + lane;
#if S3D_SORT_EARLY_VALS
    if (i < live) s_buf[lpos[r]] = val[r];
#else
    if (i < live) s_buf[lpos[r]] = vals_in ? vals_in[off + first + i] : first + i;
#endif
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kSortItems; ++j) {
    const uint32_t i = j * kSortThreads + threadIdx.x;
    if (i < live) vals_out[gpos[j]] = s_buf[i];
  }
}

constexpr int kSortTickets = kSortPasses + 4;  // one per pass + spares for the look-back kernels of the callers
inline size_t sort_aux_bytes(uint32_t n_slots) { return sizeof(uint32_t) * (size_t(kSortPasses) * 256 * n_slots + kSortTickets); }
inline uint32_t* sort_ticket(const SortState& ss, uint32_t n_slots, int i) { return ss.aux + size_t(kSortPasses) * 256 * n_slots + i; }
// a fresh tag for the status words (host side; a wrapped counter clears them first)
inline uint32_t sort_next_epoch(cudaStream_t st, const SortState& ss) {
  if (++*ss.epoch == 0) {
    cudaMemsetAsync(ss.status, 0, sizeof(uint64_t) * ss.status_words, st);
    *ss.epoch = 1;
  }
  return *ss.epoch;
}

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
  for (int p = 0; p < kSortPasses; ++p) {
    const int in = p & 1, out = in ^ 1;
    const uint32_t epoch = sort_next_epoch(st, ss);
    sort_pass_kernel<<<tm.n_tiles, kSortThreads, 0, st>>>(slots, tm, slot_tile_begin, keys[in], p == 0 ? nullptr : vals[in], keys[out], vals[out],
                                                          ss.aux + size_t(p) * n_slots * 256, ss.status, sort_ticket(ss, n_slots, p), ss.flags, epoch, 8 * p, which);
    ++*launch_counter;
  }
}

}  // namespace s3d
