// insert.cu — K8: batched graph build, restating insertSinglePoint (shard/index/vamana/
// insert.go:16-68) + robustPrune (search.go:106-138) for mini-batches of new points.
//
// Per mini-batch (all on one stream, no host round trip):
//   1. beam search (search.cuh) of every new point against the current graph, emitting the
//      visited (expanded) list in expansion order — greedySearch(vec, 1, L) (insert.go:22);
//   2. prune_new_kernel, one CTA per new point A: stable sort of the visited list by
//      distance (search.go:100), robustPrune(A) with candidate rows staged in shared
//      memory, write A's adjacency row, emit (B <- A) back-edge pairs;
//   3. stable radix sort of the pairs by target B (cub), segment heads;
//   4. backedge_kernel, one CTA per distinct target B: apply its new in-edges in batch
//      order: append, or — when deg(B)+1 > R — candidates = N(B) ∪ {A} by distance from B,
//      stable sort, robustPrune(B) (insert.go:37-65).
// With a mini-batch of 1 this is exactly the reference's sequential (1-worker) schedule
// and reproduces the oracle's graph edge for edge; larger batches search a slightly stale
// snapshot, like the reference's concurrent insert workers (vamana.go:190-195).
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cstdlib>
#include <string>

#include "common.cuh"
#include "index.cuh"

namespace sdb {

namespace {

constexpr int PRUNE_THREADS = 128;
constexpr int PRUNE_GROUPS = PRUNE_THREADS / 8;
constexpr int MAX_CAND = 256;  // visited-list capacity handed to robustPrune

struct StoreView {
  int mode;  // 0 float rows, 1 bit rows, 2 PQ codes (SDC)
  int metric;
  const float* vec; uint32_t vec_pitch; uint32_t dim;
  const uint64_t* bits; uint32_t bits_pitch; uint32_t words;
  const uint8_t* codes; uint32_t codes_pitch; uint32_t pqM, pqK;
  const float* cdist;
  uint32_t row_bytes;  // bytes staged per candidate row
};

struct PruneShared {
  uint32_t id[MAX_CAND + 1];
  float dist[MAX_CAND + 1];
  uint32_t sid[MAX_CAND + 1];   // sorted
  float sdist[MAX_CAND + 1];
  uint8_t removed[MAX_CAND + 1];
  uint32_t edges[64];
  int n;
  int cnt;
};

// distance between two staged rows by an 8-lane group; result valid in the group's lane 0.
// All 32 lanes of the warp must call it together.
template <int METRIC>
__device__ __forceinline__ float group_float_dist(const float* x, const float* y, uint32_t dim, int g) {
  constexpr bool L2 = (METRIC == METRIC_EUCLIDEAN);
  const int trips = dim >> 5;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t = 0; t < trips; ++t) {
    float4 a = *reinterpret_cast<const float4*>(x + 32 * t + 4 * g);
    float4 b = *reinterpret_cast<const float4*>(y + 32 * t + 4 * g);
    trip_accum<L2>(a, b, acc);
  }
  float tail = 0.0f;
  if (g == 0)
    for (uint32_t i = trips << 5; i < dim; ++i) tail = tail_accum<L2>(x[i], y[i], tail);
  return metric_epilogue<METRIC>(group_reduce(acc, tail));
}

__device__ __forceinline__ float row_dist(const StoreView& s, const unsigned char* x, const unsigned char* y, int g) {
  if (s.mode == 0) {
    const float* a = reinterpret_cast<const float*>(x);
    const float* b = reinterpret_cast<const float*>(y);
    switch (s.metric) {
      case METRIC_EUCLIDEAN: return group_float_dist<METRIC_EUCLIDEAN>(a, b, s.dim, g);
      case METRIC_DOT: return group_float_dist<METRIC_DOT>(a, b, s.dim, g);
      case METRIC_COSINE: return group_float_dist<METRIC_COSINE>(a, b, s.dim, g);
      default: {
        // haversine: float64 math in lane 0; keep the warp's shuffle count consistent
        float r = (g == 0) ? haversine_thread(a, b) : 0.0f;
        return r;
      }
    }
  }
  float r = 0.0f;
  if (g == 0) {
    if (s.mode == 1) {
      const uint64_t* a = reinterpret_cast<const uint64_t*>(x);
      const uint64_t* b = reinterpret_cast<const uint64_t*>(y);
      int c = 0, u = 0;
      for (uint32_t w = 0; w < s.words; ++w) {
        if (s.metric == METRIC_JACCARD) { c += __popcll(a[w] & b[w]); u += __popcll(a[w] | b[w]); }
        else c += __popcll(a[w] ^ b[w]);
      }
      r = bits_finish(s.metric, c, u);
    } else {
      // SDC: sum_i centroidDists[i][cx[i]][cy[i]] sequential f32 (product.go:299-303)
      for (uint32_t m = 0; m < s.pqM; ++m) r = __fadd_rn(r, __ldg(s.cdist + (size_t(m) * s.pqK + x[m]) * s.pqK + y[m]));
    }
  }
  return r;
}

__device__ __forceinline__ const unsigned char* global_row(const StoreView& s, uint32_t id) {
  if (s.mode == 0) return reinterpret_cast<const unsigned char*>(s.vec + size_t(id) * s.vec_pitch);
  if (s.mode == 1) return reinterpret_cast<const unsigned char*>(s.bits + size_t(id) * s.bits_pitch);
  return s.codes + size_t(id) * s.codes_pitch;
}

// Stable sort of (id, dist)[0..n) by distance into (sid, sdist): equals the reference's
// insertion sort (distset.go:223-238, strict '<' swaps => stable).
__device__ void stable_sort_by_dist(PruneShared& sh) {
  const int n = sh.n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float di = sh.dist[i];
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      float dj = sh.dist[j];
      rank += (dj < di) || (dj == di && j < i);
    }
    sh.sid[rank] = sh.id[i];
    sh.sdist[rank] = di;
  }
  __syncthreads();
}

// Stage rows of the sorted candidates into shared memory (as many as fit).
__device__ void stage_rows(const StoreView& s, PruneShared& sh, unsigned char* rows, int staged) {
  const uint32_t vec16 = s.row_bytes / 16;
  const uint32_t total = uint32_t(staged) * vec16;
  for (uint32_t t = threadIdx.x; t < total; t += blockDim.x) {
    const uint32_t c = t / vec16, i = t % vec16;
    const uint4* src = reinterpret_cast<const uint4*>(global_row(s, sh.sid[c]));
    reinterpret_cast<uint4*>(rows + size_t(c) * s.row_bytes)[i] = __ldg(src + i);
  }
  __syncthreads();
}

// robustPrune (search.go:106-138) over the sorted candidates in sh.sid/sdist; node = id of
// the node being pruned (skipped if it appears, search.go:116). Fills sh.edges/sh.cnt.
__device__ void robust_prune_cta(const StoreView& s, PruneShared& sh, const unsigned char* rows, int staged,
                                 uint32_t node, int R, float alpha) {
  const int n = sh.n;
  const int lane = threadIdx.x & 31;
  const int g = lane & 7;
  const int grp = threadIdx.x >> 3;  // 0..PRUNE_GROUPS-1
  for (int i = threadIdx.x; i < n; i += blockDim.x) sh.removed[i] = 0;
  if (threadIdx.x == 0) sh.cnt = 0;
  __syncthreads();
  for (int i = 0; i < n; ++i) {
    if (sh.removed[i] || sh.sid[i] == node) continue;  // block-uniform
    __syncthreads();
    if (threadIdx.x == 0) sh.edges[sh.cnt++] = sh.sid[i];
    __syncthreads();
    if (sh.cnt >= R) break;
    const unsigned char* xi = i < staged ? rows + size_t(i) * s.row_bytes : global_row(s, sh.sid[i]);
    for (int j0 = i + 1; j0 < n; j0 += PRUNE_GROUPS) {
      int j = j0 + grp;
      bool act = (j < n) && !sh.removed[j];
      // warp-uniform skip when none of this warp's 4 groups has work
      if (!__any_sync(SDB_FULL, act)) continue;
      int jj = act ? j : i;
      const unsigned char* yj = jj < staged ? rows + size_t(jj) * s.row_bytes : global_row(s, sh.sid[jj]);
      float d = row_dist(s, xi, yj, g);
      if (act && g == 0 && __fmul_rn(alpha, d) < sh.sdist[j]) sh.removed[j] = 1;  // search.go:132
    }
    __syncthreads();
  }
  __syncthreads();
}

struct InsertArgs {
  StoreView s;
  uint32_t* adj; uint32_t* deg; uint32_t R; float alpha;
  const uint32_t* new_ids;   // [m]
  const uint32_t* vis_ids;   // [m][vis_cap] expansion order
  const float* vis_dists;
  const uint32_t* vis_len;
  uint32_t vis_cap;
  uint32_t m;
  uint32_t* pair_key;        // [m*R] target B (INVALID_ID = none)
  uint32_t* pair_val;        // [m*R] pair index (batch order)
  int staged_max;            // rows that fit in shared memory
  uint32_t* error_flag;
};

__global__ void __launch_bounds__(PRUNE_THREADS) prune_new_kernel(InsertArgs a) {
  extern __shared__ __align__(16) unsigned char dyn[];
  __shared__ PruneShared sh;
  const uint32_t b = blockIdx.x;
  const uint32_t A = a.new_ids[b];
  uint32_t n = a.vis_len[b];
  if (n > a.vis_cap || n > MAX_CAND) {
    if (threadIdx.x == 0) atomicExch(a.error_flag, 1u);
    n = min(min(n, a.vis_cap), uint32_t(MAX_CAND));
  }
  if (threadIdx.x == 0) sh.n = int(n);
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
    sh.id[i] = a.vis_ids[size_t(b) * a.vis_cap + i];
    sh.dist[i] = a.vis_dists[size_t(b) * a.vis_cap + i];
  }
  __syncthreads();
  stable_sort_by_dist(sh);
  int staged = min(int(n), a.staged_max);
  stage_rows(a.s, sh, dyn, staged);
  robust_prune_cta(a.s, sh, dyn, staged, A, int(a.R), a.alpha);
  const int cnt = sh.cnt;
  for (uint32_t t = threadIdx.x; t < a.R; t += blockDim.x) {
    uint32_t e = t < uint32_t(cnt) ? sh.edges[t] : INVALID_ID;
    a.adj[size_t(A) * a.R + t] = e;
    a.pair_key[size_t(b) * a.R + t] = e;
    a.pair_val[size_t(b) * a.R + t] = b * a.R + t;
  }
  if (threadIdx.x == 0) a.deg[A] = uint32_t(cnt);
}

__global__ void segment_heads_kernel(const uint32_t* keys, uint32_t n, uint32_t* seg_start, uint32_t* seg_count) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t k = keys[i];
  if (k == INVALID_ID) return;
  if (i == 0 || keys[i - 1] != k) seg_start[atomicAdd(seg_count, 1u)] = i;
}

__global__ void iota_segments_kernel(const uint32_t* keys, uint32_t n, uint32_t* seg_start, uint32_t* seg_count) {
  if (threadIdx.x == 0) {
    uint32_t c = 0;
    for (uint32_t i = 0; i < n; ++i)
      if (keys[i] != INVALID_ID) seg_start[c++] = i;
    *seg_count = c;
  }
}

struct BackArgs {
  StoreView s;
  uint32_t* adj; uint32_t* deg; uint32_t R; float alpha;
  const uint32_t* new_ids;     // [m]
  const uint32_t* keys;        // sorted targets
  const uint32_t* vals;        // sorted pair indices (A index = val / R)
  uint32_t n_pairs;
  const uint32_t* seg_start;
  const uint32_t* seg_count;
  int staged_max;
};

__global__ void __launch_bounds__(PRUNE_THREADS) backedge_kernel(BackArgs a) {
  extern __shared__ __align__(16) unsigned char dyn[];
  __shared__ PruneShared sh;
  __shared__ uint32_t cur[64 + 1];
  const int lane = threadIdx.x & 31;
  const int g = lane & 7;
  const int grp = threadIdx.x >> 3;
  const uint32_t nseg = *a.seg_count;
  for (uint32_t seg = blockIdx.x; seg < nseg; seg += gridDim.x) {
    const uint32_t p0 = a.seg_start[seg];
    const uint32_t B = a.keys[p0];
    __syncthreads();
    // every thread tracks the degree in a register (a shared counter bumped by one thread
    // would race with the other warps' branch on it)
    int cur_n = int(a.deg[B]);
    for (uint32_t t = threadIdx.x; t < a.R; t += blockDim.x) cur[t] = a.adj[size_t(B) * a.R + t];
    __syncthreads();
    const unsigned char* xb = global_row(a.s, B);
    for (uint32_t p = p0; p < a.n_pairs && a.keys[p] == B; ++p) {
      const uint32_t A = a.new_ids[a.vals[p] / a.R];
      if (cur_n + 1 > int(a.R)) {
        // candidateSet.Add(nodeB.neighbours...), Add(vecA): distances from B (insert.go:48-57)
        const int n = cur_n + 1;
        if (threadIdx.x == 0) sh.n = n;
        for (int i = threadIdx.x; i < n; i += blockDim.x) sh.id[i] = i < cur_n ? cur[i] : A;
        __syncthreads();
        for (int j0 = 0; j0 < n; j0 += PRUNE_GROUPS) {
          int j = j0 + grp;
          bool act = j < n;
          if (!__any_sync(SDB_FULL, act)) continue;
          const unsigned char* yj = global_row(a.s, sh.id[act ? j : 0]);
          float d = row_dist(a.s, xb, yj, g);
          if (act && g == 0) sh.dist[j] = d;
        }
        __syncthreads();
        stable_sort_by_dist(sh);  // candidateSet.Sort() (insert.go:58)
        int staged = min(n, a.staged_max);
        stage_rows(a.s, sh, dyn, staged);
        robust_prune_cta(a.s, sh, dyn, staged, B, int(a.R), a.alpha);
        cur_n = sh.cnt;
        for (int t = threadIdx.x; t < cur_n; t += blockDim.x) cur[t] = sh.edges[t];
      } else {
        if (threadIdx.x == 0) cur[cur_n] = A;  // nodeB.AddNeighbour(vecA) (insert.go:62)
        ++cur_n;
      }
      __syncthreads();
    }
    for (uint32_t t = threadIdx.x; t < a.R; t += blockDim.x)
      a.adj[size_t(B) * a.R + t] = t < uint32_t(cur_n) ? cur[t] : INVALID_ID;
    if (threadIdx.x == 0) a.deg[B] = uint32_t(cur_n);
  }
}

// debug aid (SDB_DEBUG_INSERT=1): adjacency rows must hold deg valid ids then padding
__global__ void check_rows_kernel(const uint32_t* adj, const uint32_t* deg, const uint8_t* exists, uint32_t rows,
                                  uint32_t R, uint32_t* report) {
  uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= rows || !exists[id]) return;
  uint32_t d = deg[id];
  bool bad = d > R;
  for (uint32_t t = 0; t < R && !bad; ++t) {
    uint32_t e = adj[size_t(id) * R + t];
    if (t < d) bad = (e == INVALID_ID) || e >= rows || !exists[e] || e == id;
    else bad = e != INVALID_ID;
  }
  if (bad && atomicAdd(report, 1u) == 0) { report[1] = id; report[2] = d; }
}

StoreView make_view(const sdb_index* ix) {
  StoreView s{};
  s.vec = ix->d_vec; s.vec_pitch = ix->vec_pitch; s.dim = ix->p.dim;
  s.bits = ix->d_bits; s.bits_pitch = ix->bits_pitch; s.words = ix->words;
  s.codes = ix->d_codes; s.codes_pitch = ix->codes_pitch; s.pqM = ix->pqM; s.pqK = ix->pqK;
  s.cdist = ix->d_pq_cdist;
  if (ix->p.quantizer == SDB_QUANT_BINARY && ix->bq_fitted) {
    s.mode = 1; s.metric = ix->bq_metric; s.row_bytes = ix->bits_pitch * 8;
  } else if (ix->p.quantizer == SDB_QUANT_PRODUCT && ix->pq_fitted) {
    s.mode = 2; s.metric = ix->store_metric; s.row_bytes = ix->codes_pitch;
  } else {
    s.mode = 0; s.metric = ix->store_metric; s.row_bytes = ix->vec_pitch * 4;
  }
  return s;
}

}  // namespace

int insert_batch_locked(sdb_index* ix, uint64_t n, const uint64_t* ids, const float* vectors) {
  // classify like insertUpdateDelete (vamana.go:149-185): only fresh inserts are handled here
  std::vector<uint32_t> h32(n);
  uint64_t mx = 0;
  for (uint64_t i = 0; i < n; ++i) {
    if (ids[i] == START_ID) return fail(SDB_ERR_RESERVED_ID, "cannot modify point with start id: 1");
    if (ids[i] == 0) return fail(SDB_ERR_RESERVED_ID, "invalid point id: 0");
    if (ids[i] >= (uint64_t(1) << 31) - 1) return fail(SDB_ERR_INVALID, "node id too large for a device index");
    if (ids[i] < ix->rows && ix->h_exists[ids[i]]) return fail(SDB_ERR_STATE, "point already exists (update is not supported by the GPU index yet): " + std::to_string(ids[i]));
    mx = std::max(mx, ids[i]);
    h32[i] = uint32_t(ids[i]);
  }
  if (ix->rows <= START_ID || !ix->h_exists[START_ID]) return fail(SDB_ERR_STATE, "failed to get start point");
  int rc = index_reserve_locked(ix, mx);
  if (rc) return rc;
  cudaStream_t st = ix->stream;
  const uint32_t R = ix->p.degree_bound, L = ix->p.search_size, dim = ix->p.dim;
  const uint32_t vis_cap = MAX_CAND;
  const uint32_t max_batch = std::max<uint32_t>(1, ix->ins_max_batch);

  // device copies of ids and vectors for the whole call (chunked to bound staging)
  const uint64_t chunk_pts = std::max<uint64_t>(max_batch, (uint64_t(512) << 20) / (dim * sizeof(float)));
  sdb::DevBuf<uint32_t> d_ids;
  sdb::DevBuf<float> d_vecs;
  sdb::DevBuf<uint32_t> d_pair_key, d_pair_val, d_pair_key2, d_pair_val2, d_seg, d_misc;
  sdb::DevBuf<unsigned char> d_cubtmp;
  auto release_all = [&]() {
    d_ids.release(); d_vecs.release(); d_pair_key.release(); d_pair_val.release(); d_pair_key2.release();
    d_pair_val2.release(); d_seg.release(); d_misc.release(); d_cubtmp.release();
  };
#define INS_CHECK(expr)            \
  do {                             \
    int _rc = (expr);              \
    if (_rc) { release_all(); return _rc; } \
  } while (0)
#define INS_CUDA(expr)                                              \
  do {                                                              \
    cudaError_t _e = (expr);                                        \
    if (_e != cudaSuccess) { release_all(); return cuda_fail(_e, #expr); } \
  } while (0)

  const size_t max_pairs = size_t(max_batch) * R;
  INS_CHECK(d_pair_key.ensure(max_pairs));
  INS_CHECK(d_pair_val.ensure(max_pairs));
  INS_CHECK(d_pair_key2.ensure(max_pairs));
  INS_CHECK(d_pair_val2.ensure(max_pairs));
  INS_CHECK(d_seg.ensure(max_pairs));
  INS_CHECK(d_misc.ensure(8));
  INS_CHECK(ix->d_vis_ids.ensure(size_t(max_batch) * vis_cap));
  INS_CHECK(ix->d_vis_d.ensure(size_t(max_batch) * vis_cap));
  INS_CHECK(ix->d_vis_len.ensure(max_batch));
  INS_CHECK(ix->d_oid.ensure(max_batch));
  INS_CHECK(ix->d_od.ensure(max_batch));
  INS_CHECK(ix->d_oc.ensure(max_batch));
  size_t cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, d_pair_key.p, d_pair_key2.p, d_pair_val.p, d_pair_val2.p, int(max_pairs), 0, 32, st);
  INS_CHECK(d_cubtmp.ensure(cub_bytes + 16));
  INS_CUDA(cudaMemsetAsync(d_misc.p, 0, 8 * sizeof(uint32_t), st));
  uint32_t* d_err = d_misc.p + 1;
  uint32_t* d_segcount = d_misc.p;

  // shared memory budget for staged candidate rows
  size_t static_smem = sizeof(PruneShared) + 512;
  cudaFuncAttributes fa;
  INS_CUDA(cudaFuncGetAttributes(&fa, prune_new_kernel));
  static_smem = std::max(static_smem, fa.sharedSizeBytes);
  INS_CUDA(cudaFuncGetAttributes(&fa, backedge_kernel));
  static_smem = std::max(static_smem, fa.sharedSizeBytes);

  const bool debug = getenv("SDB_DEBUG_INSERT") != nullptr;
  uint64_t inserted_before = ix->count > 0 ? ix->count - 1 : 0;  // user points already in the graph
  for (uint64_t c0 = 0; c0 < n; c0 += chunk_pts) {
    const uint64_t cn = std::min(chunk_pts, n - c0);
    INS_CHECK(d_ids.ensure(cn));
    INS_CHECK(d_vecs.ensure(size_t(cn) * dim));
    INS_CUDA(cudaMemcpyAsync(d_ids.p, h32.data() + c0, cn * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    INS_CUDA(cudaMemcpyAsync(d_vecs.p, vectors + c0 * dim, size_t(cn) * dim * sizeof(float), cudaMemcpyHostToDevice, st));
    // vecStore.Set for the chunk (insert.go:17): rows become visible, nothing points at them yet
    INS_CHECK(set_rows_device(ix, uint32_t(cn), d_ids.p, d_vecs.p, st));
    StoreView view = make_view(ix);
    // rows staged per CTA: aim for >= 2 CTAs/SM
    size_t budget = std::min<size_t>(ix->smem_optin, size_t(100) << 10);
    size_t dyn_budget = budget > static_smem ? budget - static_smem : 0;
    int staged_max = int(std::min<size_t>(MAX_CAND, dyn_budget / view.row_bytes));
    size_t dyn_smem = size_t(staged_max) * view.row_bytes;
    INS_CUDA(cudaFuncSetAttribute(prune_new_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(dyn_smem)));
    INS_CUDA(cudaFuncSetAttribute(backedge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(dyn_smem)));

    uint64_t done = 0;
    while (done < cn) {
      uint64_t total_in = inserted_before + c0 + done;
      uint64_t want = std::max<uint64_t>(ix->ins_min_batch, total_in / std::max<uint32_t>(1, ix->ins_growth_div));
      uint32_t m = uint32_t(std::min<uint64_t>(std::min<uint64_t>(want, max_batch), cn - done));
      if (m == 0) m = 1;
      const uint32_t* b_ids = d_ids.p + done;
      const float* b_vecs = d_vecs.p + done * dim;
      // 1. greedySearch(vec, 1, L) with the visited list
      INS_CHECK(launch_search(ix, m, b_vecs, 1, L, ix->d_oid.p, ix->d_od.p, ix->d_oc.p, ix->d_vis_ids.p, ix->d_vis_d.p,
                              ix->d_vis_len.p, vis_cap, nullptr, 0, nullptr, st));
      // 2. robustPrune(A) + emit back-edge pairs
      InsertArgs ia{};
      ia.s = view; ia.adj = ix->d_adj; ia.deg = ix->d_deg; ia.R = R; ia.alpha = ix->p.alpha;
      ia.new_ids = b_ids; ia.vis_ids = ix->d_vis_ids.p; ia.vis_dists = ix->d_vis_d.p; ia.vis_len = ix->d_vis_len.p;
      ia.vis_cap = vis_cap; ia.m = m; ia.pair_key = d_pair_key.p; ia.pair_val = d_pair_val.p;
      ia.staged_max = staged_max; ia.error_flag = d_err;
      prune_new_kernel<<<m, PRUNE_THREADS, dyn_smem, st>>>(ia);
      ix->launches++;
      INS_CUDA(cudaGetLastError());
      // 3. group pairs by target (stable => batch order within a target)
      const uint32_t np = m * R;
      const uint32_t* skeys = d_pair_key.p;
      const uint32_t* svals = d_pair_val.p;
      if (m > 1) {
        size_t tb = cub_bytes;
        INS_CUDA(cub::DeviceRadixSort::SortPairs(d_cubtmp.p, tb, d_pair_key.p, d_pair_key2.p, d_pair_val.p, d_pair_val2.p, int(np), 0, 32, st));
        ix->launches += 4;
        skeys = d_pair_key2.p;
        svals = d_pair_val2.p;
      }
      INS_CUDA(cudaMemsetAsync(d_segcount, 0, sizeof(uint32_t), st));
      if (m > 1) {
        segment_heads_kernel<<<(np + 255) / 256, 256, 0, st>>>(skeys, np, d_seg.p, d_segcount);
        ix->launches++;
        INS_CUDA(cudaGetLastError());
      }
      // 4. back-edges
      BackArgs ba{};
      ba.s = view; ba.adj = ix->d_adj; ba.deg = ix->d_deg; ba.R = R; ba.alpha = ix->p.alpha;
      ba.new_ids = b_ids; ba.keys = skeys; ba.vals = svals; ba.n_pairs = np; ba.seg_start = d_seg.p;
      ba.seg_count = d_segcount; ba.staged_max = staged_max;
      if (m == 1) {
        // a single new point: every target is its own segment, in edge order — no sort needed
        iota_segments_kernel<<<1, 64, 0, st>>>(skeys, np, d_seg.p, d_segcount);
        ix->launches++;
      }
      uint32_t grid = std::min<uint32_t>(np, uint32_t(ix->sm_count) * 8);
      backedge_kernel<<<grid, PRUNE_THREADS, dyn_smem, st>>>(ba);
      ix->launches++;
      INS_CUDA(cudaGetLastError());
      if (debug) {
        INS_CUDA(cudaMemsetAsync(d_misc.p + 4, 0, 3 * sizeof(uint32_t), st));
        check_rows_kernel<<<(ix->rows + 255) / 256, 256, 0, st>>>(ix->d_adj, ix->d_deg, ix->d_exists, ix->rows, R, d_misc.p + 4);
        uint32_t rep[3];
        INS_CUDA(cudaMemcpyAsync(rep, d_misc.p + 4, sizeof(rep), cudaMemcpyDeviceToHost, st));
        INS_CUDA(cudaStreamSynchronize(st));
        if (rep[0]) {
          release_all();
          return fail(SDB_ERR_INTERNAL, "insert debug: " + std::to_string(rep[0]) + " inconsistent rows after a batch of " + std::to_string(m) + " at offset " + std::to_string(c0 + done) + ", first node " + std::to_string(rep[1]) + " deg " + std::to_string(rep[2]));
        }
      }
      done += m;
    }
    INS_CUDA(cudaStreamSynchronize(st));
  }
  uint32_t h_misc[2] = {0, 0};
  INS_CUDA(cudaMemcpy(h_misc, d_misc.p, sizeof(h_misc), cudaMemcpyDeviceToHost));
  release_all();
#undef INS_CHECK
#undef INS_CUDA
  for (uint64_t i = 0; i < n; ++i) {
    if (!ix->h_exists[h32[i]]) {
      ix->h_exists[h32[i]] = 1;
      ix->count++;
    }
    if (h32[i] > ix->max_node_id) ix->max_node_id = h32[i];
  }
  if (h_misc[1]) return fail(SDB_ERR_INTERNAL, "visited list exceeded the robustPrune candidate capacity");
  return SDB_OK;
}

}  // namespace sdb
