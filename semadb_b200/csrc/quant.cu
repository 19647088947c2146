// quant.cu — K7/K9: VectorStore.Fit on the GPU.
//
//  * binaryQuantizer.Fit (shard/vectorstore/binary.go:145-185): per-dimension mean of every
//    stored vector as the threshold (f32 running sum in ascending id order, then /count),
//    then re-encode all rows (binary.go:103-129).
//  * productQuantizer.Fit (product.go:175-236): one Lloyd k-means per sub-vector
//    (utils/kmeans.go:34-150) — farthest-first init from a given first row, <= 100
//    assign/update rounds, stop when no label changes — then codes = labels,
//    flatCentroids, and the K x K centroidDists table with the index metric.
//
// The k-means follows the reference bit for bit, including its aliasing quirk: centroids
// are sub-slices of the input rows (kmeans.go:63,82), so the update step writes the new
// means through into the stored vectors (kmeans.go:144). Here the centroids are *row
// indices* into the device vector array and the update writes in place, which gives the
// same observable behaviour. Sums are accumulated sequentially in row order per
// (cluster, dimension) like kmeans.go:125-137, so every f32 value matches the reference.
// The sub-spaces are independent (product.go:201-233): the farthest-first init runs one CTA per
// sub-space (K-1 dependent rounds), the Lloyd rounds run as grid-wide kernels — assign over
// (sub-space x row tiles), ordered sums with a warp per cluster, write-through — with a per
// sub-space stop flag. sub = dim/M is 8 at C4: the assign is 2*n*K*sub FLOP per sub-space and
// round with an inner dimension of 8, nothing a tensor-core tile could chew on (and bf16/tf32
// products would break the exact argmin); it is an exact FFMA kernel.
#include <algorithm>
#include <cfloat>

#include "common.cuh"
#include "index.cuh"

namespace sdb {

namespace {

constexpr int KM_THREADS = 1024;

__global__ void bq_mean_kernel(const float* vec, uint32_t pitch, const uint32_t* row_ids, uint32_t n, uint32_t dim,
                               float* thr) {
  uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= dim) return;
  float sum = 0.0f;
  for (uint32_t r = 0; r < n; ++r) sum = __fadd_rn(sum, vec[size_t(row_ids[r]) * pitch + d]);  // binary.go:161-163
  thr[d] = __fdiv_rn(sum, float(n));                                                            // binary.go:171-173
}

struct KmArgs {
  float* vec; uint32_t pitch;
  const uint32_t* row_ids; uint32_t n;
  uint32_t sub, K, M, max_iter;
  uint32_t first;
  uint8_t* labels;      // [M][n]
  float* min_dist;      // [M][n]
  float* sums;          // [M][K*sub]
  uint32_t* counts;     // [M][K]
  uint32_t* cent_row;   // [M][K] index into row_ids
  uint32_t* iters;      // [M]
  uint32_t* changes;    // [M] labels changed by the current assign round
  uint32_t* active;     // [M] sub-space still iterating
  uint32_t* any_active; // [1]
  int stage_centroids;  // K*sub floats fit in shared memory
  __device__ __forceinline__ float* X(uint32_t m, uint32_t r) const { return vec + size_t(row_ids[r]) * pitch + m * sub; }
};

struct BestPair {
  float v;
  uint32_t i;
};
// reference scan: strict '>' from (0, id 0), ascending j => highest value, lowest index
__device__ __forceinline__ BestPair better(BestPair a, BestPair b) {
  if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
  return a;
}

// ---- init (kmeans.go:54-83): farthest-first from a given first row. K-1 dependent rounds, each a
// distance pass over the n rows and a block-wide argmax with the reference's first-max rule: one
// CTA per sub-space (the sub-spaces are independent, product.go:201-233).
__global__ void __launch_bounds__(KM_THREADS) km_init_kernel(KmArgs a) {
  extern __shared__ __align__(16) float cs[];  // the last chosen centre [sub]
  __shared__ BestPair red[32];
  const uint32_t m = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  uint8_t* labels = a.labels + size_t(m) * a.n;
  float* md = a.min_dist + size_t(m) * a.n;
  uint32_t* crow = a.cent_row + size_t(m) * a.K;
  for (uint32_t j = tid; j < a.n; j += KM_THREADS) { md[j] = FLT_MAX; labels[j] = 0; }
  if (tid == 0) {
    crow[0] = a.first;
    a.active[m] = 1;
    a.changes[m] = 0;
    a.iters[m] = 0;
  }
  __syncthreads();
  for (uint32_t i = 1; i < a.K; ++i) {
    const float* c = a.X(m, crow[i - 1]);
    for (uint32_t t = tid; t < a.sub; t += KM_THREADS) cs[t] = c[t];
    __syncthreads();
    BestPair best{0.0f, 0u};
    for (uint32_t j = tid; j < a.n; j += KM_THREADS) {
      if (j == a.first) continue;  // only randId is in alreadyCentroid (kmeans.go:62,69)
      float d = float_dist_thread<METRIC_EUCLIDEAN>(a.X(m, j), cs, int(a.sub));
      float cur = md[j];
      if (d < cur) { cur = d; md[j] = d; }
      best = better(best, BestPair{cur, j});
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      BestPair other{__shfl_down_sync(SDB_FULL, best.v, o), __shfl_down_sync(SDB_FULL, best.i, o)};
      best = better(best, other);
    }
    if (lane == 0) red[wid] = best;
    __syncthreads();
    if (wid == 0) {
      BestPair b2 = lane < KM_THREADS / 32 ? red[lane] : BestPair{0.0f, 0u};
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) {
        BestPair other{__shfl_down_sync(SDB_FULL, b2.v, o), __shfl_down_sync(SDB_FULL, b2.i, o)};
        b2 = better(b2, other);
      }
      if (lane == 0) crow[i] = b2.v > 0.0f ? b2.i : 0u;
    }
    __syncthreads();
  }
}

// ---- assign (kmeans.go:100-115): argmin over the K centres, strict '<' (lowest index wins).
// One CTA per (row tile, sub-space): the sub-space's centres sit in shared memory, a thread owns a
// row. SUBT > 0: the row's sub-vector lives in registers. Grid M x ceil(n / tile): every SM works.
constexpr int KM_TILE = 256;
template <int SUBT>
__global__ void __launch_bounds__(KM_TILE) km_assign_kernel(KmArgs a) {
  extern __shared__ __align__(16) float cs[];  // [K][sub] if staged
  const uint32_t m = blockIdx.y;
  if (!a.active[m]) return;
  const uint32_t sub = SUBT > 0 ? uint32_t(SUBT) : a.sub;
  uint32_t* crow = a.cent_row + size_t(m) * a.K;
  if (a.stage_centroids) {
    for (uint32_t t = threadIdx.x; t < a.K * sub; t += KM_TILE) cs[t] = a.X(m, crow[t / sub])[t % sub];
    __syncthreads();
  }
  uint8_t* labels = a.labels + size_t(m) * a.n;
  uint32_t changes = 0;
  for (uint32_t j = blockIdx.x * KM_TILE + threadIdx.x; j < a.n; j += gridDim.x * KM_TILE) {
    const float* xg = a.X(m, j);
    float xr[SUBT > 0 ? SUBT : 1];
    if (SUBT > 0) {
#pragma unroll
      for (int t = 0; t < SUBT; ++t) xr[t] = xg[t];
    }
    const float* x = SUBT > 0 ? xr : xg;
    float best = float_dist_thread<METRIC_EUCLIDEAN>(x, a.stage_centroids ? cs : a.X(m, crow[0]), int(sub));
    uint32_t bid = 0;
    for (uint32_t i = 1; i < a.K; ++i) {
      const float* c = a.stage_centroids ? cs + size_t(i) * sub : a.X(m, crow[i]);
      float d = float_dist_thread<METRIC_EUCLIDEAN>(x, c, int(sub));
      if (d < best) { best = d; bid = i; }
    }
    if (labels[j] != uint8_t(bid)) { ++changes; labels[j] = uint8_t(bid); }
  }
  changes = __reduce_add_sync(SDB_FULL, changes);
  if ((threadIdx.x & 31) == 0 && changes) atomicAdd(a.changes + m, changes);
}

// ---- "no label changed => stop" (kmeans.go:116-118), per sub-space; the round still counts
__global__ void km_flags_kernel(KmArgs a, uint32_t iter) {
  const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= a.M) return;
  if (a.active[m]) {
    a.iters[m] = iter + 1;
    if (a.changes[m] == 0) a.active[m] = 0;
    else atomicOr(a.any_active, 1u);
    a.changes[m] = 0;
  }
}

// ---- update (kmeans.go:121-138): per-(cluster, dimension) sums in ROW ORDER, sequential f32.
// One warp per cluster walks the label array once, 32 rows per step: a ballot marks the rows of
// its cluster and they are added in ascending order, lane jd owning dimension jd — the same
// chain of additions as the reference, without every (cluster, dimension) thread scanning all
// n labels on its own.
__global__ void __launch_bounds__(256) km_update_kernel(KmArgs a) {
  const uint32_t m = blockIdx.y;
  if (!a.active[m]) return;
  const int lane = threadIdx.x & 31;
  const uint32_t k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (k >= a.K) return;
  const uint8_t* labels = a.labels + size_t(m) * a.n;
  for (uint32_t jd0 = 0; jd0 < a.sub; jd0 += 32) {
    const uint32_t jd = jd0 + lane;
    float s = 0.0f;
    uint32_t cnt = 0;
    for (uint32_t r0 = 0; r0 < a.n; r0 += 32) {
      const uint32_t r = r0 + lane;
      uint32_t mask = __ballot_sync(SDB_FULL, r < a.n && labels[r] == uint8_t(k));
      cnt += __popc(mask);
      while (mask) {
        const uint32_t rr = r0 + __ffs(mask) - 1;
        mask &= mask - 1;
        if (jd < a.sub) s = __fadd_rn(s, a.X(m, rr)[jd]);
      }
    }
    if (jd < a.sub) a.sums[(size_t(m) * a.K + k) * a.sub + jd] = s;
    if (jd0 == 0 && lane == 0) a.counts[size_t(m) * a.K + k] = cnt;
  }
}

// ---- means written through the aliased centroid rows, in centroid order (kmeans.go:140-146):
// the centres ARE rows of the stored vectors (kmeans.go:63,82), so this overwrites caller data
// exactly like the reference; an empty cluster keeps its centre.
__global__ void km_write_kernel(KmArgs a) {
  const uint32_t m = blockIdx.x;
  if (!a.active[m]) return;
  const uint32_t* crow = a.cent_row + size_t(m) * a.K;
  const uint32_t* counts = a.counts + size_t(m) * a.K;
  for (uint32_t jd = threadIdx.x; jd < a.sub; jd += blockDim.x) {
    for (uint32_t i = 0; i < a.K; ++i) {
      const uint32_t c = counts[i];
      if (c == 0) continue;
      a.X(m, crow[i])[jd] = __fdiv_rn(a.sums[(size_t(m) * a.K + i) * a.sub + jd], float(c));
    }
  }
}

// codes of the training rows = k-means labels (product.go:216-218); flatCentroids copy
// (product.go:220-223); centroidDists with the index metric (product.go:225-230).
template <int METRIC>
__global__ void pq_finalize_kernel(const float* vec, uint32_t pitch, const uint32_t* row_ids, uint32_t n, uint32_t M,
                                   uint32_t K, uint32_t sub, const uint8_t* labels, const uint32_t* cent_row,
                                   uint8_t* codes, uint32_t codes_pitch, float* flat, float* cdist) {
  const uint32_t m = blockIdx.x;
  const uint32_t* crow = cent_row + size_t(m) * K;
  for (uint32_t r = threadIdx.x; r < n; r += blockDim.x) codes[size_t(row_ids[r]) * codes_pitch + m] = labels[size_t(m) * n + r];
  for (uint32_t t = threadIdx.x; t < K * sub; t += blockDim.x)
    flat[size_t(m) * K * sub + t] = vec[size_t(row_ids[crow[t / sub]]) * pitch + m * sub + (t % sub)];
  __syncthreads();
  for (uint32_t t = threadIdx.x; t < K * K; t += blockDim.x) {
    const uint32_t j = t / K, k = t % K;
    cdist[size_t(m) * K * K + t] = float_dist_thread<METRIC>(flat + (size_t(m) * K + j) * sub, flat + (size_t(m) * K + k) * sub, int(sub));
  }
}

}  // namespace

int fit_locked(sdb_index* ix, uint64_t pq_first_row, int32_t* fitted) {
  ix->vec_epoch++;  // k-means writes centroid means through aliased rows (kmeans.go:144)
  *fitted = 0;
  if (ix->p.quantizer == SDB_QUANT_NONE) return SDB_OK;  // plainStore.Fit (plain.go:72-74)
  cudaStream_t st = ix->stream;
  if (ix->p.quantizer == SDB_QUANT_BINARY) {
    if (ix->bq_fitted || ix->count < ix->p.bq_trigger) return SDB_OK;  // binary.go:148-150
  } else {
    if (ix->pq_fitted || ix->count < ix->p.pq_trigger) return SDB_OK;  // product.go:177-183
  }
  std::vector<uint32_t> rows;
  rows.reserve(ix->count);
  for (uint32_t id = 0; id < ix->rows; ++id)
    if (ix->h_exists[id]) rows.push_back(id);
  const uint32_t n = uint32_t(rows.size());
  if (n == 0) return SDB_OK;
  int rc;
  if ((rc = ix->d_tmp32.ensure(n))) return rc;
  SDB_CUDA(cudaMemcpyAsync(ix->d_tmp32.p, rows.data(), size_t(n) * 4, cudaMemcpyHostToDevice, st));

  if (ix->p.quantizer == SDB_QUANT_BINARY) {
    bq_mean_kernel<<<(ix->p.dim + 127) / 128, 128, 0, st>>>(ix->d_vec, ix->vec_pitch, ix->d_tmp32.p, n, ix->p.dim, ix->d_bq_thr);
    ix->launches++;
    SDB_CUDA(cudaGetLastError());
    ix->bq_fitted = true;
    if ((rc = launch_encode_rows(ix, n, ix->d_tmp32.p, st))) return rc;
    SDB_CUDA(cudaStreamSynchronize(st));
    *fitted = 1;
    return SDB_OK;
  }

  // product quantizer
  if (pq_first_row >= n) return fail(SDB_ERR_INVALID, "pq_first_row out of range");
  const uint32_t M = ix->pqM, K = ix->pqK, sub = ix->pqSub;
  sdb::DevBuf<uint8_t> d_labels;
  sdb::DevBuf<float> d_md, d_sums;
  sdb::DevBuf<uint32_t> d_counts;
  auto cleanup = [&]() { d_labels.release(); d_md.release(); d_sums.release(); d_counts.release(); };
  if ((rc = d_labels.ensure(size_t(M) * n)) || (rc = d_md.ensure(size_t(M) * n)) || (rc = d_sums.ensure(size_t(M) * K * sub)) ||
      (rc = d_counts.ensure(size_t(M) * K * 2 + 4 * size_t(M) + 8))) {
    cleanup();
    return rc;
  }
  size_t nc = size_t(M) * K * sub, nd = size_t(M) * K * K;
  if (!ix->d_pq_centroids && cudaMalloc(reinterpret_cast<void**>(&ix->d_pq_centroids), nc * 4) != cudaSuccess) {
    cleanup();
    return cuda_fail(cudaGetLastError(), "cudaMalloc(flatCentroids)");
  }
  if (!ix->d_pq_cdist && cudaMalloc(reinterpret_cast<void**>(&ix->d_pq_cdist), nd * 4) != cudaSuccess) {
    cleanup();
    return cuda_fail(cudaGetLastError(), "cudaMalloc(centroidDists)");
  }
  KmArgs ka{};
  ka.vec = ix->d_vec; ka.pitch = ix->vec_pitch; ka.row_ids = ix->d_tmp32.p; ka.n = n;
  ka.sub = sub; ka.K = K; ka.M = M; ka.max_iter = 100;  // product.go:209
  ka.first = uint32_t(pq_first_row);
  ka.labels = d_labels.p; ka.min_dist = d_md.p; ka.sums = d_sums.p;
  ka.counts = d_counts.p; ka.cent_row = d_counts.p + size_t(M) * K; ka.iters = d_counts.p + size_t(M) * K * 2;
  ka.changes = ka.iters + M; ka.active = ka.changes + M; ka.any_active = ka.active + M;
  size_t full = size_t(K) * sub * sizeof(float);
  ka.stage_centroids = full <= 96 * 1024;
  cudaError_t e;
#define KM_CUDA(expr)                                                  \
  do {                                                                 \
    if ((e = (expr)) != cudaSuccess) { cleanup(); return cuda_fail(e, #expr); } \
  } while (0)
  KM_CUDA(cudaFuncSetAttribute(km_init_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(size_t(sub) * 4)));
  km_init_kernel<<<M, KM_THREADS, size_t(sub) * 4, st>>>(ka);
  ix->launches++;
  KM_CUDA(cudaGetLastError());
  // Lloyd iterations (kmeans.go:95-147): assign over (sub-space x row tiles), the stop test per
  // sub-space, then the ordered sums and the write-through — sub-spaces that have converged drop out
  const size_t asmem = ka.stage_centroids ? full : 0;
  auto assign = sub == 4 ? km_assign_kernel<4> : sub == 8 ? km_assign_kernel<8> : sub == 16 ? km_assign_kernel<16> : km_assign_kernel<0>;
  KM_CUDA(cudaFuncSetAttribute(assign, cudaFuncAttributeMaxDynamicSharedMemorySize, int(asmem)));
  const dim3 agrid(std::max<uint32_t>(1, std::min<uint32_t>((n + KM_TILE - 1) / KM_TILE, 64)), M);
  const dim3 ugrid((K + 7) / 8, M);
  uint32_t* h_any = nullptr;
  KM_CUDA(cudaMallocHost(reinterpret_cast<void**>(&h_any), sizeof(uint32_t)));
  for (uint32_t it = 0; it < ka.max_iter; ++it) {
    assign<<<agrid, KM_TILE, asmem, st>>>(ka);
    cudaMemsetAsync(ka.any_active, 0, sizeof(uint32_t), st);
    km_flags_kernel<<<(M + 127) / 128, 128, 0, st>>>(ka, it);
    cudaMemcpyAsync(h_any, ka.any_active, sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
    ix->launches += 2;
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) { cudaFreeHost(h_any); cleanup(); return cuda_fail(e, "k-means iteration"); }
    if (*h_any == 0) break;
    km_update_kernel<<<ugrid, 256, 0, st>>>(ka);
    km_write_kernel<<<M, 64, 0, st>>>(ka);
    ix->launches += 2;
  }
  cudaFreeHost(h_any);
  KM_CUDA(cudaGetLastError());
#undef KM_CUDA
  if (ix->store_metric == SDB_METRIC_EUCLIDEAN)
    pq_finalize_kernel<METRIC_EUCLIDEAN><<<M, 256, 0, st>>>(ix->d_vec, ix->vec_pitch, ix->d_tmp32.p, n, M, K, sub, d_labels.p, ka.cent_row, ix->d_codes, ix->codes_pitch, ix->d_pq_centroids, ix->d_pq_cdist);
  else
    pq_finalize_kernel<METRIC_DOT><<<M, 256, 0, st>>>(ix->d_vec, ix->vec_pitch, ix->d_tmp32.p, n, M, K, sub, d_labels.p, ka.cent_row, ix->d_codes, ix->codes_pitch, ix->d_pq_centroids, ix->d_pq_cdist);
  ix->launches++;
  if ((e = cudaGetLastError()) != cudaSuccess) { cleanup(); return cuda_fail(e, "pq_finalize_kernel"); }
  e = cudaStreamSynchronize(st);
  cleanup();
  if (e != cudaSuccess) return cuda_fail(e, "cudaStreamSynchronize(fit)");
  ix->pq_fitted = true;
  *fitted = 1;
  return SDB_OK;
}

}  // namespace sdb
