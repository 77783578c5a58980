// voxel.cu — batch set-up and the VoxelGrid down-sampling kernels.
//
// Replaces pcl::VoxelGrid<PointXYZ>::filter as called from PointCloudSensor::downsample
// (slam3d/sensor/pcl/PointCloudSensor.cpp:190-201) and twice per align() (:127-131).  Semantics: SURVEY A.1.
//   bbox      min/max over finite points (ordered-uint atomics; min/max are order independent => deterministic)
//   params    per-slot scalars of A.1 steps 1,3,4 incl. the int32 overflow guard (one thread per slot)
//   keys      A.1 step 5, float ops in PCL's order via __fmul_rn/__fsub_rn (no FMA contraction)
//   sort      stable segmented radix sort (sort.cuh)                    A.1 step 6
//   heads     run starts -> per-tile counts -> per-slot scan           A.1 step 7
//   centroid  one thread per run sums its points in ascending input order in float, / float(count)   A.1 steps 8,9
// All kernels are streaming and HBM/L2-bandwidth bound: 16 B/point in, 8 B/point keys, 16 B/voxel out.
#include "internal.h"
#include "sort.cuh"

namespace s3d {

// ------------------------------------------------------------------------------------------------------------
void Workspace::init(int dev) {
  device = dev;
  S3D_CUDA(cudaSetDevice(dev));
  S3D_CUDA(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev));
  S3D_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  S3D_CUDA(cudaEventCreateWithFlags(&sync_event, cudaEventBlockingSync | cudaEventDisableTiming));
  S3D_CUDA(cudaEventCreateWithFlags(&input_event, cudaEventDisableTiming));
  flags.reserve(64);
  h_small.reserve(4096);
}

// Single calls spin on the stream (lowest latency).  The chunks of a batch call run on several host threads per device — and
// under torchrun on several ranks per box — so they sleep on a blocking event instead: 8 ranks x 6 spinning threads on a
// 32-core host was the 8-GPU scaling loss of round 1.
void Workspace::sync() {
  if (blocking_sync && sync_event) {
    S3D_CUDA(cudaEventRecord(sync_event, stream));
    S3D_CUDA(cudaEventSynchronize(sync_event));
  } else {
    S3D_CUDA(cudaStreamSynchronize(stream));
  }
}

// device inputs produced on the caller's stream (s3d_set_input_stream): order this workspace's stream after it
void Workspace::order_after_input() {
  if (!wait_input) return;
  S3D_CUDA(cudaEventRecord(input_event, input_stream));
  S3D_CUDA(cudaStreamWaitEvent(stream, input_event, 0));
}

cudaEvent_t Workspace::get_event() {
  if (!event_pool.empty()) { cudaEvent_t e = event_pool.back(); event_pool.pop_back(); return e; }
  cudaEvent_t e;
  S3D_CUDA(cudaEventCreate(&e));
  return e;
}

void Workspace::collect_spans() {
  for (Span& sp : spans) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) { stage_ms[sp.stage] += ms; stage_launches[sp.stage] += sp.n_launch; }
    else cudaGetLastError();
    event_pool.push_back(sp.a); event_pool.push_back(sp.b);
  }
  spans.clear();
}

void Workspace::destroy() {
  cudaSetDevice(device);
  if (stream) cudaStreamSynchronize(stream);
  collect_spans();
  for (cudaEvent_t e : event_pool) cudaEventDestroy(e);
  event_pool.clear();
  if (sync_event) cudaEventDestroy(sync_event);
  if (input_event) cudaEventDestroy(input_event);
  sync_event = input_event = nullptr;
  for (auto& g : loop_graphs) cudaGraphExecDestroy(g.second);
  for (cudaGraph_t g : loop_graph_defs) cudaGraphDestroy(g);
  loop_graphs.clear(); loop_graph_defs.clear();

  DevBuf* bufs[] = {&slots, &pairs, &raw_stage, &work, &gpts, &keys0, &keys1, &vals0, &vals1, &hist, &sort_totals, &long_runs, &tile_slot, &tile_first,
                    &slot_tile_begin, &tile_heads, &hash, &normals, &moved, &prev_nn, &sec_lb, &moments, &eval_part, &corr, &mahal, &iter_tile_pair, &iter_tile_first,
                    &fit_partial, &flags, &accu, &accu2, &map_aux, &ndt_pairs, &ndt_leaves, &ndt_hash, &ndt_part, &gicp_args, &gicp_sched, &knn_arena};
  for (DevBuf* b : bufs) b->release();
  h_slots.release(); h_pairs.release(); h_small.release(); h_tiles.release(); h_bounce.release();
  if (stream) cudaStreamDestroy(stream);
  stream = nullptr;
}

// ------------------------------------------------------------------------------------------------------------
// Batch set-up: slot table, tile table, input staging.  clouds[2p] = slam3d source of pair p, clouds[2p+1] = target
// (or any list of clouds for the stage-level entry points).
void setup_batch(Workspace& ws, const std::vector<const float*>& clouds, const std::vector<uint64_t>& sizes, uint32_t n_pairs) {
  S3D_CUDA(cudaSetDevice(ws.device));
  const uint32_t ns = static_cast<uint32_t>(clouds.size());
  ws.n_slots = ns; ws.n_pairs = n_pairs;
  ws.h_off.resize(ns); ws.h_n.resize(ns);
  uint64_t total = 0; uint32_t n_tiles = 0;
  for (uint32_t s = 0; s < ns; ++s) {
    ws.h_off[s] = static_cast<uint32_t>(total);
    ws.h_n[s] = static_cast<uint32_t>(sizes[s]);
    total += (sizes[s] + 3) & ~uint64_t(3);  // keep every slot 64-byte aligned
    n_tiles += static_cast<uint32_t>((sizes[s] + kSortTile - 1) / kSortTile);
  }
  if (total >= (1ull << 31)) throw CudaError{"batch too large: more than 2^31 points"};
  ws.total = static_cast<uint32_t>(total); ws.n_tiles = n_tiles;
  const size_t tot = std::max<size_t>(total, 4);
  ws.slots.reserve(sizeof(SlotInfo) * ns);
  ws.work.reserve(16 * tot); ws.gpts.reserve(16 * tot); ws.normals.reserve(32 * tot);
  // hash arena: ~1.2 occupied cells per point is typical for voxel-filtered lidar scans; 3 entries/point leaves 25% head-room
  // at load factor 1/2.  Sparse clouds (every point alone in its cell on many levels) overflow it: the layout kernel
  // then flags kErrHashArena and reports the exact need, and the caller re-runs the batch (with_arena_retry in api.cu).
  {
    const size_t want = std::max(ws.hash_want, 3 * size_t(total) + 64 * size_t(ns));
    if (ws.hash_cap < want) { ws.hash.reserve(sizeof(HashEntry) * want); ws.hash_cap = want; }
  }
  ws.pair_off.resize(n_pairs);
  ws.max_na = 0;
  for (uint32_t p = 0; p < n_pairs; ++p) { ws.pair_off[p] = ws.h_off[2 * p + 1]; ws.max_na = std::max(ws.max_na, ws.h_n[2 * p + 1]); }
  ws.keys0.reserve(4 * tot); ws.keys1.reserve(4 * tot); ws.vals0.reserve(4 * tot); ws.vals1.reserve(4 * tot);
  {
    const size_t words = 256 * size_t(std::max<uint32_t>(n_tiles, 1));
    if (words > ws.status_words) {  // new memory must not hold words that look like a current epoch
      ws.hist.reserve(sizeof(uint64_t) * words);
      ws.status_words = ws.hist.cap / sizeof(uint64_t);
      S3D_CUDA(cudaMemsetAsync(ws.hist.p, 0, ws.hist.cap, ws.stream));
    }
  }
  ws.sort_totals.reserve(sort_aux_bytes(std::max<uint32_t>(ns, 1)));
  ws.long_runs.reserve(sizeof(uint4) * (tot / 64 + ns + 1));
  ws.tile_slot.reserve(4 * std::max<uint32_t>(n_tiles, 1)); ws.tile_first.reserve(4 * std::max<uint32_t>(n_tiles, 1));
  ws.tile_heads.reserve(4 * std::max<uint32_t>(n_tiles, 1));
  ws.slot_tile_begin.reserve(4 * (ns + 1));
  ws.h_slots.reserve(sizeof(SlotInfo) * ns);
  ws.h_tiles.reserve(4 * (2 * size_t(n_tiles) + ns + 1));

  // classify input pointers; host memory (pinned or pageable) and misaligned device memory go through raw_stage
  bool need_stage = false, need_bounce = false;
  std::vector<int> on_device(ns, 0), pageable(ns, 0);
  for (uint32_t s = 0; s < ns; ++s) {
    if (sizes[s] == 0) continue;
    cudaPointerAttributes at{};
    cudaError_t e = cudaPointerGetAttributes(&at, clouds[s]);
    if (e != cudaSuccess) { cudaGetLastError(); at.type = cudaMemoryTypeUnregistered; }
    on_device[s] = (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) && (reinterpret_cast<uintptr_t>(clouds[s]) & 15) == 0;
    pageable[s] = at.type == cudaMemoryTypeUnregistered;
    if (!on_device[s]) need_stage = true;
    if (pageable[s]) need_bounce = true;
  }
  if (need_stage) ws.raw_stage.reserve(16 * tot);
  // Pageable host clouds (a std::vector, the points of a pcl::PointCloud): cudaMemcpyAsync from pageable memory is staged by the
  // driver through one small pinned buffer and blocks the calling thread (~10 GB/s, measured: e2e 1800 against 2430
  // registrations/s from pinned memory).  The chunk's host thread copies them into the workspace's own pinned buffer instead —
  // the host threads of the other chunks do the same in parallel, before they queue for the upload turn — and the DMA runs from there.
  static const bool bounce_enabled = [] { const char* e = getenv("S3D_PINNED_BOUNCE"); return !e || atoi(e) != 0; }();  // 0: A/B measurements
  if (need_bounce && bounce_enabled) {
    ws.h_bounce.reserve(16 * tot);
    for (uint32_t s = 0; s < ns; ++s)
      if (pageable[s]) memcpy(ws.h_bounce.as<char>() + 16 * size_t(ws.h_off[s]), clouds[s], 16 * sizes[s]);
  } else {
    need_bounce = false;
  }

  ws.order_after_input();
  std::unique_lock<std::mutex> gate;  // uploads of concurrent chunks take turns (api.cu, t_gate_uploads)
  if (need_stage && ws.upload_gate) gate = std::unique_lock<std::mutex>(*ws.upload_gate);
  SlotInfo* hs = ws.h_slots.as<SlotInfo>();
  uint32_t* ht = ws.h_tiles.as<uint32_t>();
  uint32_t* h_tile_slot = ht; uint32_t* h_tile_first = ht + n_tiles; uint32_t* h_begin = ht + 2 * size_t(n_tiles);
  uint32_t t = 0;
  for (uint32_t s = 0; s < ns; ++s) {
    SlotInfo& si = hs[s];
    memset(&si, 0, sizeof si);
    si.off = ws.h_off[s]; si.n_raw = ws.h_n[s];
    si.gpts = ws.gpts.as<float4>() + si.off; si.normals = ws.normals.as<double4>() + si.off; si.table = ws.hash.as<HashEntry>();
    for (int a = 0; a < 3; ++a) {  // ordered-uint encodings of +inf / -inf, filled in by the bbox kernel
      reinterpret_cast<uint32_t&>(si.bb_min[a]) = 0xFFFFFFFFu; reinterpret_cast<uint32_t&>(si.bb_max[a]) = 0u;
      reinterpret_cast<uint32_t&>(si.g_min[a]) = 0xFFFFFFFFu; reinterpret_cast<uint32_t&>(si.g_max[a]) = 0u;
    }
    if (sizes[s] == 0) si.raw = nullptr;
    else if (on_device[s]) si.raw = reinterpret_cast<const float4*>(clouds[s]);
    else {
      si.raw = ws.raw_stage.as<float4>() + si.off;
      const void* from = (need_bounce && pageable[s]) ? static_cast<const void*>(ws.h_bounce.as<char>() + 16 * size_t(si.off)) : static_cast<const void*>(clouds[s]);
      S3D_CUDA(cudaMemcpyAsync(ws.raw_stage.as<float4>() + si.off, from, 16 * sizes[s], cudaMemcpyDefault, ws.stream));
      ws.h2d += 16 * sizes[s];
    }
    h_begin[s] = t;
    for (uint64_t f = 0; f < sizes[s]; f += kSortTile) { h_tile_slot[t] = s; h_tile_first[t] = static_cast<uint32_t>(f); ++t; }
  }
  h_begin[ns] = t;
  if (gate.owns_lock()) { ws.sync(); gate.unlock(); }
  S3D_CUDA(cudaMemcpyAsync(ws.slots.p, hs, sizeof(SlotInfo) * ns, cudaMemcpyHostToDevice, ws.stream));
  if (n_tiles) {
    S3D_CUDA(cudaMemcpyAsync(ws.tile_slot.p, h_tile_slot, 4 * size_t(n_tiles), cudaMemcpyHostToDevice, ws.stream));
    S3D_CUDA(cudaMemcpyAsync(ws.tile_first.p, h_tile_first, 4 * size_t(n_tiles), cudaMemcpyHostToDevice, ws.stream));
  }
  S3D_CUDA(cudaMemcpyAsync(ws.slot_tile_begin.p, h_begin, 4 * size_t(ns + 1), cudaMemcpyHostToDevice, ws.stream));
  S3D_CUDA(cudaMemsetAsync(ws.flags.p, 0, 64, ws.stream));
}

// ------------------------------------------------------------------------------------------------------------
// bbox over finite points of the raw cloud (which = kCountRaw -> bb_*) or the working cloud (kCountPts -> g_*)
__global__ void __launch_bounds__(kSortThreads) bbox_kernel(SlotInfo* __restrict__ slots, TileMap tm, const float4* __restrict__ work, int which) {
  const uint32_t t = blockIdx.x;
  const uint32_t slot = tm.tile_slot[t], first = tm.tile_first[t];
  SlotInfo& si = slots[slot];
  const uint32_t n = slot_count(si, which);
  if (first >= n) return;
  const float4* p = which == kCountRaw ? si.raw : work + si.off;
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  uint32_t cnt = 0;
#pragma unroll
  for (int j = 0; j < kSortTile / kSortThreads; ++j) {
    const uint32_t e = first + j * kSortThreads + threadIdx.x;
    if (e < n) {
      const float4 v = p[e];
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
  __shared__ float smn[8][3], smx[8][3];
  __shared__ uint32_t scnt[8];
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { for (int a = 0; a < 3; ++a) { smn[w][a] = mn[a]; smx[w][a] = mx[a]; } scnt[w] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) {
      for (int a = 0; a < 3; ++a) { mn[a] = fminf(mn[a], smn[i][a]); mx[a] = fmaxf(mx[a], smx[i][a]); }
      cnt += scnt[i];
    }
    if (cnt) {
      uint32_t* dmin = reinterpret_cast<uint32_t*>(which == kCountRaw ? si.bb_min : si.g_min);
      uint32_t* dmax = reinterpret_cast<uint32_t*>(which == kCountRaw ? si.bb_max : si.g_max);
#pragma unroll
      for (int a = 0; a < 3; ++a) { atomicMin(&dmin[a], float_to_ordered(mn[a])); atomicMax(&dmax[a], float_to_ordered(mx[a])); }
      if (which == kCountRaw) atomicAdd(&si.n_finite, cnt);
    }
  }
}

// A.1 steps 1, 3, 4 — one thread per slot.  leaf <= 0: no filtering, the working cloud is the raw cloud.
__global__ void voxel_params_kernel(SlotInfo* __restrict__ slots, uint32_t n_slots, float leaf) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots) return;
  SlotInfo& si = slots[s];
  for (int a = 0; a < 3; ++a) {
    si.bb_min[a] = ordered_to_float(reinterpret_cast<uint32_t&>(si.bb_min[a]));
    si.bb_max[a] = ordered_to_float(reinterpret_cast<uint32_t&>(si.bb_max[a]));
  }
  if (!(leaf > 0.f)) { si.overflow = 1; si.n_pts = si.n_raw; return; }  // pass-through (treated like the overflow copy)
  if (si.n_finite == 0) { si.n_pts = 0; si.overflow = 0; return; }
  const float inv = __fdiv_rn(1.0f, leaf);
  si.inv_leaf = inv;
  const long long dx = (long long)__fmul_rn(__fsub_rn(si.bb_max[0], si.bb_min[0]), inv) + 1;
  const long long dy = (long long)__fmul_rn(__fsub_rn(si.bb_max[1], si.bb_min[1]), inv) + 1;
  const long long dz = (long long)__fmul_rn(__fsub_rn(si.bb_max[2], si.bb_min[2]), inv) + 1;
  if (dx * dy * dz > 2147483647ll) { si.overflow = 1; si.n_pts = si.n_raw; return; }  // "output = *input_"
  int div_b[3];
  for (int a = 0; a < 3; ++a) {
    si.min_b[a] = (int)floorf(__fmul_rn(si.bb_min[a], inv));
    const int max_b = (int)floorf(__fmul_rn(si.bb_max[a], inv));
    div_b[a] = max_b - si.min_b[a] + 1;
  }
  si.mul1 = (uint32_t)div_b[0];
  si.mul2 = (uint32_t)div_b[0] * (uint32_t)div_b[1];
  si.overflow = 0;
}

// A.1 step 5
// + the digit totals of all four sort passes (sort.cuh), so that the sort never reads the keys just to count them
__global__ void __launch_bounds__(kSortThreads) voxel_keys_kernel(const SlotInfo* __restrict__ slots, TileMap tm, uint32_t n_slots, uint32_t* __restrict__ keys,
                                                                   uint32_t* __restrict__ totals) {
  __shared__ uint32_t sh[kSortPasses][256];
  const uint32_t t = blockIdx.x;
  const uint32_t slot = tm.tile_slot[t], first = tm.tile_first[t];
  const SlotInfo& si = slots[slot];
  if (si.overflow) return;
#pragma unroll
  for (int p = 0; p < kSortPasses; ++p) sh[p][threadIdx.x] = 0;
  __syncthreads();
  const float inv = si.inv_leaf;
  const float mb0 = (float)si.min_b[0], mb1 = (float)si.min_b[1], mb2 = (float)si.min_b[2];
#pragma unroll
  for (int j = 0; j < kSortTile / kSortThreads; ++j) {
    const uint32_t e = first + j * kSortThreads + threadIdx.x;
    if (e < si.n_raw) {
      const float4 v = si.raw[e];
      uint32_t key = kInvalidKey;
      if (finite3(v.x, v.y, v.z)) {
        const int i0 = (int)__fsub_rn(floorf(__fmul_rn(v.x, inv)), mb0);
        const int i1 = (int)__fsub_rn(floorf(__fmul_rn(v.y, inv)), mb1);
        const int i2 = (int)__fsub_rn(floorf(__fmul_rn(v.z, inv)), mb2);
        key = (uint32_t)i0 + (uint32_t)i1 * si.mul1 + (uint32_t)i2 * si.mul2;
      }
      keys[si.off + e] = key;
      count_digits(sh, key);
    }
  }
  __syncthreads();
  flush_digits(sh, totals, n_slots, slot);
}

// run starts among the first n_finite sorted keys -> count per tile
__global__ void __launch_bounds__(kSortThreads) voxel_heads_kernel(const SlotInfo* __restrict__ slots, TileMap tm,
                                                                    const uint32_t* __restrict__ keys, uint32_t* __restrict__ tile_heads) {
  __shared__ uint32_t wsum[8];
  const uint32_t t = blockIdx.x;
  const uint32_t slot = tm.tile_slot[t], first = tm.tile_first[t];
  const SlotInfo& si = slots[slot];
  if (si.overflow) return;
  const uint32_t n = si.n_finite;
  const uint32_t* k = keys + si.off;
  uint32_t c = 0;
#pragma unroll
  for (int j = 0; j < kSortTile / kSortThreads; ++j) {
    const uint32_t e = first + j * kSortThreads + threadIdx.x;
    if (e < n) c += (e == 0 || k[e] != k[e - 1]) ? 1u : 0u;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) { uint32_t s = 0; for (int i = 0; i < 8; ++i) s += wsum[i]; tile_heads[t] = s; }
}

// exclusive scan of the tile counts of each slot (one warp per slot); n_pts = number of voxels
__global__ void voxel_scan_kernel(SlotInfo* __restrict__ slots, uint32_t n_slots, const uint32_t* __restrict__ slot_tile_begin,
                                  uint32_t* __restrict__ tile_heads) {
  const uint32_t s = blockIdx.x;
  if (s >= n_slots) return;
  SlotInfo& si = slots[s];
  if (si.overflow) return;
  const uint32_t ntiles = (si.n_finite + kSortTile - 1) / kSortTile;
  uint32_t* h = tile_heads + slot_tile_begin[s];
  const int lane = threadIdx.x;
  uint32_t running = 0;
  for (uint32_t b = 0; b < ntiles; b += 32) {
    const uint32_t t = b + lane;
    const uint32_t v = t < ntiles ? h[t] : 0u;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t u = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += u; }
    if (t < ntiles) h[t] = running + incl - v;
    running += __shfl_sync(0xFFFFFFFFu, incl, 31);
  }
  if (lane == 0) si.n_pts = running;
}

// A.1 steps 8-9: float centroid per run (voxel), summed in ascending input order — (((p0 + p1) + p2) + ...) exactly as
// PCL's CentroidPoint — and written in ascending key order.  Two kernels:
//   voxel_centroid_kernel       one thread per run; right for scan-sized clouds (2-3 points per voxel).  Runs longer than
//                               kLongRun are only queued: with dependent index->point loads a thread needs ~0.4 us per point,
//                               and map-sized clouds have voxels with thousands of points (1.16 ms on 2M points, 0.2 m leaf);
//   voxel_long_centroid_kernel  one warp per queued run: 32 lanes fetch 32 points at once, then the warp consumes them in
//                               order through shuffles, so the additions stay strictly sequential but never wait for loads.
constexpr uint32_t kLongRun = 64;

// sorted[e] = raw[vals[e]]: the points in voxel-key order, so that the centroid kernels stream them instead of chasing
// index -> point through two dependent loads per addition
__global__ void __launch_bounds__(kSortThreads) voxel_gather_kernel(const SlotInfo* __restrict__ slots, TileMap tm, const uint32_t* __restrict__ vals,
                                                                     float4* __restrict__ sorted) {
  const uint32_t t = blockIdx.x;
  const uint32_t slot = tm.tile_slot[t], first = tm.tile_first[t];
  const SlotInfo& si = slots[slot];
  if (si.overflow) return;
#pragma unroll
  for (int j = 0; j < kSortTile / kSortThreads; ++j) {
    const uint32_t e = first + j * kSortThreads + threadIdx.x;
    if (e < si.n_finite) sorted[si.off + e] = si.raw[vals[si.off + e]];
  }
}

__global__ void __launch_bounds__(kSortThreads) voxel_centroid_kernel(const SlotInfo* __restrict__ slots, TileMap tm,
                                                                       const uint32_t* __restrict__ keys, const float4* __restrict__ sorted,
                                                                       const uint32_t* __restrict__ tile_heads, float4* __restrict__ work,
                                                                       uint4* __restrict__ long_runs, uint32_t* __restrict__ n_long) {
  __shared__ uint32_t wsum[8];
  const uint32_t t = blockIdx.x;
  const uint32_t slot = tm.tile_slot[t], first = tm.tile_first[t];
  const SlotInfo& si = slots[slot];
  if (si.overflow) return;
  const uint32_t n = si.n_finite;
  if (first >= n) return;
  const uint32_t* k = keys + si.off;
  const float4* pts = sorted + si.off;
  constexpr int kPer = kSortTile / kSortThreads;
  const uint32_t e0 = first + threadIdx.x * kPer;  // kPer consecutive sorted positions per thread
  uint32_t head_mask = 0, cnt = 0;
  uint32_t prev = (e0 > 0 && e0 < n) ? k[e0 - 1] : 0u;
#pragma unroll
  for (int j = 0; j < kPer; ++j) {
    const uint32_t e = e0 + j;
    if (e < n) {
      const uint32_t kk = k[e];
      if (e == 0 || kk != prev) { head_mask |= 1u << j; ++cnt; }
      prev = kk;
    }
  }
  // block exclusive scan of cnt
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { uint32_t u = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += u; }
  if (lane == 31) wsum[w] = incl;
  __syncthreads();
  uint32_t rank = tile_heads[t] + incl - cnt;
  for (int i = 0; i < w; ++i) rank += wsum[i];
  float4* out = work + si.off;
#pragma unroll
  for (int j = 0; j < kPer; ++j) {
    if (!(head_mask & (1u << j))) continue;
    const uint32_t e = e0 + j;
    const uint32_t kk = k[e];
    if (e + kLongRun < n && k[e + kLongRun] == kk) {  // sorted keys: the run has more than kLongRun points
      long_runs[atomicAdd(n_long, 1u)] = make_uint4(slot, e, rank, 0u);
      ++rank;
      continue;
    }
    float sx = 0.f, sy = 0.f, sz = 0.f;
    uint32_t l = e;
    do {
      const float4 p = pts[l];
      sx = __fadd_rn(sx, p.x); sy = __fadd_rn(sy, p.y); sz = __fadd_rn(sz, p.z);
      ++l;
    } while (l < n && k[l] == kk);
    const float c = (float)(l - e);
    out[rank++] = make_float4(__fdiv_rn(sx, c), __fdiv_rn(sy, c), __fdiv_rn(sz, c), 1.0f);
  }
}

__global__ void __launch_bounds__(256) voxel_long_centroid_kernel(const SlotInfo* __restrict__ slots, const uint32_t* __restrict__ keys,
                                                                  const float4* __restrict__ sorted, float4* __restrict__ work,
                                                                  const uint4* __restrict__ long_runs, const uint32_t* __restrict__ n_long) {
  const uint32_t FULL = 0xFFFFFFFFu;
  const int lane = threadIdx.x & 31;
  const uint32_t warps = gridDim.x * (blockDim.x >> 5);
  constexpr int kDepth = 4;  // batches of 32 points in flight per warp
  for (uint32_t q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); q < *n_long; q += warps) {
    const uint4 job = long_runs[q];
    const SlotInfo& si = slots[job.x];
    const uint32_t n = si.n_finite;
    const uint32_t* k = keys + si.off;
    const float4* pts = sorted + si.off;
    const uint32_t kk = k[job.y];
    // length of the run: keys are sorted, so gallop + binary search for the first position with another key
    uint32_t lo = job.y + kLongRun, step = kLongRun, hi;
    for (;;) { hi = lo + step; if (hi >= n) { hi = n; break; } if (k[hi] != kk) break; lo = hi; step <<= 1; }
    while (lo + 1 < hi) { const uint32_t mid = (lo + hi) >> 1; if (k[mid] == kk) lo = mid; else hi = mid; }
    const uint32_t end = hi;  // k[lo] == kk, k[hi] != kk (or hi == n)
    float sx = 0.f, sy = 0.f, sz = 0.f;
    float4 buf[kDepth];
#pragma unroll
    for (int d = 0; d < kDepth; ++d) { const uint32_t e = job.y + d * 32 + lane; buf[d] = e < end ? pts[e] : make_float4(0.f, 0.f, 0.f, 0.f); }
    for (uint32_t b = job.y; b < end; b += 32 * kDepth) {
#pragma unroll
      for (int d = 0; d < kDepth; ++d) {
        const float4 p = buf[d];
        const uint32_t nb = b + (d + kDepth) * 32 + lane;           // refill this slot with the batch kDepth ahead
        buf[d] = nb < end ? pts[nb] : make_float4(0.f, 0.f, 0.f, 0.f);
        const int take = (int)min(32u, end > b + d * 32 ? end - (b + d * 32) : 0u);
        for (int j = 0; j < take; ++j) {
          sx = __fadd_rn(sx, __shfl_sync(FULL, p.x, j));
          sy = __fadd_rn(sy, __shfl_sync(FULL, p.y, j));
          sz = __fadd_rn(sz, __shfl_sync(FULL, p.z, j));
        }
      }
    }
    const float c = (float)(end - job.y);
    if (lane == 0) work[si.off + job.z] = make_float4(__fdiv_rn(sx, c), __fdiv_rn(sy, c), __fdiv_rn(sz, c), 1.0f);
  }
}

// overflow / pass-through slots: working cloud = raw cloud verbatim
__global__ void __launch_bounds__(kSortThreads) voxel_passthrough_kernel(const SlotInfo* __restrict__ slots, TileMap tm, float4* __restrict__ work) {
  const uint32_t t = blockIdx.x;
  const uint32_t slot = tm.tile_slot[t], first = tm.tile_first[t];
  const SlotInfo& si = slots[slot];
  if (!si.overflow) return;
#pragma unroll
  for (int j = 0; j < kSortTile / kSortThreads; ++j) {
    const uint32_t e = first + j * kSortThreads + threadIdx.x;
    if (e < si.n_raw) work[si.off + e] = si.raw[e];
  }
}

void launch_bbox(Workspace& ws, int which) {
  TileMap tm{ws.tile_slot.as<uint32_t>(), ws.tile_first.as<uint32_t>(), ws.n_tiles};
  bbox_kernel<<<ws.n_tiles, kSortThreads, 0, ws.stream>>>(ws.slots.as<SlotInfo>(), tm, ws.work.as<float4>(), which);
  ++ws.launches;
}

void run_voxel(Workspace& ws, float leaf, uint32_t* leaf_keys) {
  if (ws.n_tiles == 0) {  // all clouds empty: n_pts stays 0
    return;
  }
  cudaStream_t st = ws.stream;
  StageTimer timer(ws, kStageVoxel);
  SlotInfo* slots = ws.slots.as<SlotInfo>();
  TileMap tm{ws.tile_slot.as<uint32_t>(), ws.tile_first.as<uint32_t>(), ws.n_tiles};
  launch_bbox(ws, kCountRaw);
  voxel_params_kernel<<<(ws.n_slots + 63) / 64, 64, 0, st>>>(slots, ws.n_slots, leaf);
  ++ws.launches;
  if (leaf > 0.f) {
    uint32_t* keys[2] = {ws.keys0.as<uint32_t>(), ws.keys1.as<uint32_t>()};
    uint32_t* vals[2] = {ws.vals0.as<uint32_t>(), ws.vals1.as<uint32_t>()};
    if (leaf_keys) S3D_CUDA(cudaMemsetAsync(keys[0], 0xFF, 4 * size_t(ws.total), st));  // skipped / overflow points report 0xFFFFFFFF
    const SortState ss = ws.sort_state();
    sort_clear_aux(st, ss, ws.n_slots);
    voxel_keys_kernel<<<ws.n_tiles, kSortThreads, 0, st>>>(slots, tm, ws.n_slots, keys[0], ss.aux);
    ++ws.launches;
    if (leaf_keys) S3D_CUDA(cudaMemcpyAsync(leaf_keys, keys[0], 4 * size_t(ws.total), cudaMemcpyDeviceToDevice, st));
    radix_sort_segmented(st, slots, ws.n_slots, tm, ws.slot_tile_begin.as<uint32_t>(), keys, vals, ss, kCountRaw, /*digits_done=*/true, &ws.launches);
    voxel_heads_kernel<<<ws.n_tiles, kSortThreads, 0, st>>>(slots, tm, keys[0], ws.tile_heads.as<uint32_t>());
    voxel_scan_kernel<<<ws.n_slots, 32, 0, st>>>(slots, ws.n_slots, ws.slot_tile_begin.as<uint32_t>(), ws.tile_heads.as<uint32_t>());
    uint32_t* n_long = ws.flags.as<uint32_t>() + 8;  // flags[8]: number of queued long runs (flags are zeroed by setup_batch)
    float4* sorted = ws.gpts.as<float4>();  // free until the NN grid is built
    voxel_gather_kernel<<<ws.n_tiles, kSortThreads, 0, st>>>(slots, tm, vals[0], sorted);
    voxel_centroid_kernel<<<ws.n_tiles, kSortThreads, 0, st>>>(slots, tm, keys[0], sorted, ws.tile_heads.as<uint32_t>(), ws.work.as<float4>(),
                                                               ws.long_runs.as<uint4>(), n_long);
    voxel_long_centroid_kernel<<<ws.n_sms * 2, 256, 0, st>>>(slots, keys[0], sorted, ws.work.as<float4>(), ws.long_runs.as<uint4>(), n_long);
    ws.launches += 5;
  }
  voxel_passthrough_kernel<<<ws.n_tiles, kSortThreads, 0, st>>>(slots, tm, ws.work.as<float4>());
  ++ws.launches;
  S3D_CUDA(cudaGetLastError());
}

}  // namespace s3d
