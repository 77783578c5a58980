// Host build of the product's search logic (slam3d_b200/csrc/nn_search.cuh, knn_walk.cuh) for the CPU test-suite
// (tests/test_hostsearch.py).  TEST-ONLY: one host thread plays one device thread (tests/cuda_host_shim.h), so what the walk
// scans, prunes and certifies can be checked against brute force without a GPU.  The multi-resolution voxel hash is built here
// by a plain host restatement of grid.cu (same cell size rule, Morton keys, stable order, (level, cell) -> range entries).
#include "cuda_host_shim.h"

#include <algorithm>
#include <numeric>
#include <vector>

// work statistics: every cell scan of the walks reports (level, slot of the 27-block, points in the cell)
struct HsTrace { std::vector<uint32_t> query, visit, slot, count; uint32_t cur_query = 0, cur_visit = 0; int last_level = -1; };
static HsTrace* g_trace = nullptr;
static inline void hs_trace(int level, int slot, uint32_t count) {
  if (!g_trace) return;
  if (g_trace->last_level >= 0 && level != g_trace->last_level) ++g_trace->cur_visit;  // the walk moved on to a coarser level
  g_trace->last_level = level;
  g_trace->query.push_back(g_trace->cur_query); g_trace->visit.push_back(g_trace->cur_visit);
  g_trace->slot.push_back((uint32_t)slot); g_trace->count.push_back(count);
}
static inline void hs_trace_next_query(uint32_t q) { if (g_trace) { g_trace->cur_query = q; g_trace->cur_visit = 0; g_trace->last_level = -1; } }
#define S3D_SCAN_TRACE(level, slot, count) hs_trace((level), (slot), (uint32_t)(count))

#include "../slam3d_b200/csrc/knn_walk.cuh"

namespace {

struct HostGrid {
  std::vector<float4> pts;         // Morton-sorted, .w = original index bits
  std::vector<s3d::HashEntry> table;
  s3d::GridView view;
};

int cell_of(float u, int dim) { return (int)floorf(fminf(fmaxf(u, 0.f), (float)(dim - 1))); }  // grid.cu cell_of

}  // namespace

extern "C" {

// grid.cu: grid_params_kernel, grid_keys_kernel, stable sort, grid_gather_kernel, hash_layout / hash_insert
void* hs_build_grid(const float* xyz, uint64_t n, float leaf_hint) {
  HostGrid* G = new HostGrid();
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (uint64_t i = 0; i < n; ++i)
    for (int a = 0; a < 3; ++a) { lo[a] = fminf(lo[a], xyz[3 * i + a]); hi[a] = fmaxf(hi[a], xyz[3 * i + a]); }
  float ext = 0.f, amax = 0.f;
  for (int a = 0; a < 3; ++a) { ext = fmaxf(ext, hi[a] - lo[a]); amax = fmaxf(amax, fmaxf(fabsf(lo[a]), fabsf(hi[a]))); }
  s3d::GridView& g = G->view;
  const float span = ext * 1.001f + 1e-6f;
  float h0 = leaf_hint > 0.f ? 3.0f * leaf_hint : span / 1024.f;
  int nlev = 1;
  while (nlev < s3d::kMaxLevels && h0 * (float)(1 << nlev) <= span) ++nlev;
  if (h0 * (float)(1 << nlev) <= span) h0 = span / (float)(1 << nlev);
  g.h0 = h0; g.inv_h0 = 1.0f / h0; g.nlev = nlev; g.margin = 1e-4f * h0 + 16.f * 1.1920929e-7f * (amax + ext);
  g.ox = lo[0]; g.oy = lo[1]; g.oz = lo[2]; g.n = (uint32_t)n;
  const int dim = 1 << nlev;
  std::vector<uint32_t> key(n), order(n);
  for (uint64_t i = 0; i < n; ++i)
    key[i] = s3d::morton3(cell_of(s3d::grid_coord(xyz[3 * i], lo[0], g.inv_h0), dim), cell_of(s3d::grid_coord(xyz[3 * i + 1], lo[1], g.inv_h0), dim),
                          cell_of(s3d::grid_coord(xyz[3 * i + 2], lo[2], g.inv_h0), dim));
  std::iota(order.begin(), order.end(), 0u);
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return key[a] < key[b]; });
  std::vector<uint32_t> k(n);
  G->pts.resize(n);
  for (uint64_t e = 0; e < n; ++e) {
    const uint32_t src = order[e];
    k[e] = key[src];
    G->pts[e] = make_float4(xyz[3 * src], xyz[3 * src + 1], xyz[3 * src + 2], __uint_as_float(src));
  }
  auto started = [&](uint64_t e) {
    if (e == 0) return nlev;
    const uint32_t x = k[e] ^ k[e - 1];
    if (x == 0) return 0;
    const int l = (31 - __clz(x)) / 3 + 1;
    return l < nlev ? l : nlev;
  };
  uint64_t n_cells = 0;
  for (uint64_t e = 0; e < n; ++e) n_cells += started(e);
  const uint32_t cap = (uint32_t)(2 * n_cells + 8);
  G->table.assign(cap, s3d::HashEntry{0u, 0xFFFFFFFFu, 0u, 0u});
  for (uint64_t e = 0; e < n; ++e) {
    const int st = started(e);
    for (int L = 0; L < st; ++L) {
      const uint32_t ck = k[e] >> (3 * L);
      uint64_t end = e + 1;
      while (end < n && (k[end] >> (3 * L)) == ck) ++end;
      uint32_t s = s3d::hash_slot(ck, (uint32_t)L, cap);
      while (G->table[s].level != 0xFFFFFFFFu) if (++s == cap) s = 0;
      G->table[s] = s3d::HashEntry{ck, (uint32_t)L, (uint32_t)e, (uint32_t)end};
    }
  }
  g.table = G->table.data(); g.pts = G->pts.data(); g.cap = cap;
  return G;
}

void hs_free_grid(void* grid) { delete static_cast<HostGrid*>(grid); }
int hs_levels(void* grid) { return static_cast<HostGrid*>(grid)->view.nlev; }

// nn_search as gicp_iter_kernel / nn_stage_kernel call it; gather != 0: the two-pass 27-block (scan_block<true>).
// hints: sorted positions (or 0xFFFFFFFF); outputs per query: original index, squared distance, lb2 (bound for every other
// point) and the sorted position of the winner (usable as the next hint)
void hs_nn(void* grid, const float* q, uint64_t nq, float cutoff2, int gather, const uint32_t* hints, uint32_t* idx, float* d2, float* lb2, uint32_t* pos) {
  const s3d::GridView& g = static_cast<HostGrid*>(grid)->view;
  uint2 cells[s3d::kNNGatherCap];
  for (uint64_t i = 0; i < nq; ++i) {
    hs_trace_next_query((uint32_t)i);
    const uint32_t hint = hints ? hints[i] : s3d::kNoIndex;
    const s3d::NNResult r = gather ? s3d::nn_search<true>(g, q[3 * i], q[3 * i + 1], q[3 * i + 2], cutoff2, hint, s3d::kNoIndex, cells)
                                   : s3d::nn_search<false>(g, q[3 * i], q[3 * i + 1], q[3 * i + 2], cutoff2, hint);
    idx[i] = r.idx; d2[i] = r.d2; lb2[i] = r.lb2; pos[i] = r.pos;
  }
}

// the search part of knn_cov_kernel for every point of the grid's own cloud; outputs by ORIGINAL index, ascending (d2, index)
void hs_knn(void* grid, int k, uint32_t* idx, float* d2) {
  const s3d::GridView& g = static_cast<HostGrid*>(grid)->view;
  const uint32_t n = g.n;
  const int kk = k < (int)n ? k : (int)n;
  std::vector<uint64_t> heap((size_t)k * s3d::kKnnThreads);
  const s3d::SmemHeap h{heap.data()};
  for (uint32_t r = 0; r < n; ++r) {
    hs_trace_next_query(r);  // sorted position = thread number
    const float4 qv = g.pts[r];
    const uint32_t q_orig = __float_as_uint(qv.w);
    const float ux = s3d::clamp_coord(s3d::grid_coord(qv.x, g.ox, g.inv_h0));
    const float uy = s3d::clamp_coord(s3d::grid_coord(qv.y, g.oy, g.inv_h0));
    const float uz = s3d::clamp_coord(s3d::grid_coord(qv.z, g.oz, g.inv_h0));
    const int cnt = s3d::thread_walk(g, qv, ux, uy, uz, s3d::knn_start_level(g, ux, uy, uz), s3d::KMAX, h, kk);
    if (cnt < kk) for (int i = s3d::heap_last_parent(cnt); i >= 0; --i) s3d::heap_sift_down(h, cnt, h.at(i), i);
    for (int m = cnt - 1; m > 0; --m) {  // heapsort, as knn_finish
      const uint64_t last = h.at(m);
      h.at(m) = h.at(0);
      s3d::heap_sift_down(h, m, last);
    }
    for (int j = 0; j < k; ++j) {
      const bool have = j < cnt;
      idx[(size_t)q_orig * k + j] = have ? (uint32_t)h.at(j) : s3d::kNoIndex;
      d2[(size_t)q_orig * k + j] = have ? __uint_as_float((uint32_t)(h.at(j) >> 32)) : INFINITY;
    }
  }
}

// work statistics: start recording / fetch the records (query, level visit, cell slot, points scanned) and stop
void hs_trace_begin() { delete g_trace; g_trace = new HsTrace(); }
uint64_t hs_trace_size() { return g_trace ? g_trace->query.size() : 0; }
void hs_trace_end(uint32_t* query, uint32_t* visit, uint32_t* slot, uint32_t* count) {
  if (!g_trace) return;
  const size_t n = g_trace->query.size();
  std::copy_n(g_trace->query.begin(), n, query); std::copy_n(g_trace->visit.begin(), n, visit);
  std::copy_n(g_trace->slot.begin(), n, slot); std::copy_n(g_trace->count.begin(), n, count);
  delete g_trace; g_trace = nullptr;
}

}  // extern "C"
