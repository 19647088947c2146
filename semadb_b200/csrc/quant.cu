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
// One CTA per sub-vector; the sub-spaces are independent (product.go:201-233).
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
  uint32_t sub, K, max_iter;
  uint32_t first;
  uint8_t* labels;      // [M][n]
  float* min_dist;      // [M][n]
  float* sums;          // [M][K*sub]
  uint32_t* counts;     // [M][K]
  uint32_t* cent_row;   // [M][K] index into row_ids
  uint32_t* iters;      // [M]
  int stage_centroids;  // K*sub floats fit in shared memory
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

__global__ void __launch_bounds__(KM_THREADS) kmeans_kernel(KmArgs a) {
  extern __shared__ __align__(16) float cs[];  // staged centroids [K][sub] or one centroid [sub]
  __shared__ BestPair red[32];
  __shared__ uint32_t s_changes;
  __shared__ uint32_t s_pick;
  const uint32_t m = blockIdx.x;
  const uint32_t off = m * a.sub;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  uint8_t* labels = a.labels + size_t(m) * a.n;
  float* md = a.min_dist + size_t(m) * a.n;
  float* sums = a.sums + size_t(m) * a.K * a.sub;
  uint32_t* counts = a.counts + size_t(m) * a.K;
  uint32_t* crow = a.cent_row + size_t(m) * a.K;
  auto X = [&](uint32_t r) -> float* { return a.vec + size_t(a.row_ids[r]) * a.pitch + off; };

  // ---- init (kmeans.go:54-83)
  for (uint32_t j = tid; j < a.n; j += KM_THREADS) { md[j] = FLT_MAX; labels[j] = 0; }
  if (tid == 0) crow[0] = a.first;
  __syncthreads();
  for (uint32_t i = 1; i < a.K; ++i) {
    const float* c = X(crow[i - 1]);
    for (uint32_t t = tid; t < a.sub; t += KM_THREADS) cs[t] = c[t];
    __syncthreads();
    BestPair best{0.0f, 0u};
    for (uint32_t j = tid; j < a.n; j += KM_THREADS) {
      if (j == a.first) continue;  // only randId is in alreadyCentroid (kmeans.go:62,69)
      float d = float_dist_thread<METRIC_EUCLIDEAN>(X(j), cs, int(a.sub));
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

  // ---- Lloyd iterations (kmeans.go:95-147)
  uint32_t iter = 0;
  for (; iter < a.max_iter; ++iter) {
    if (a.stage_centroids) {
      for (uint32_t t = tid; t < a.K * a.sub; t += KM_THREADS) cs[t] = X(crow[t / a.sub])[t % a.sub];
    }
    if (tid == 0) s_changes = 0;
    __syncthreads();
    uint32_t changes = 0;
    for (uint32_t j = tid; j < a.n; j += KM_THREADS) {
      const float* x = X(j);
      float best = float_dist_thread<METRIC_EUCLIDEAN>(x, a.stage_centroids ? cs : X(crow[0]), int(a.sub));
      uint32_t bid = 0;
      for (uint32_t i = 1; i < a.K; ++i) {
        const float* c = a.stage_centroids ? cs + size_t(i) * a.sub : X(crow[i]);
        float d = float_dist_thread<METRIC_EUCLIDEAN>(x, c, int(a.sub));
        if (d < best) { best = d; bid = i; }
      }
      if (labels[j] != uint8_t(bid)) { ++changes; labels[j] = uint8_t(bid); }
    }
    if (changes) atomicAdd(&s_changes, changes);
    __syncthreads();
    if (s_changes == 0) { ++iter; break; }  // kmeans.go:116-118 (this round still counts)
    // update: per-label sums in row order, sequential f32 (kmeans.go:121-138)
    for (uint32_t k = tid; k < a.K; k += KM_THREADS) {
      uint32_t c = 0;
      for (uint32_t r = 0; r < a.n; ++r) c += (labels[r] == k);
      counts[k] = c;
    }
    for (uint32_t t = tid; t < a.K * a.sub; t += KM_THREADS) {
      const uint32_t k = t / a.sub, jd = t % a.sub;
      float s = 0.0f;
      for (uint32_t r = 0; r < a.n; ++r)
        if (labels[r] == k) s = __fadd_rn(s, X(r)[jd]);
      sums[t] = s;
    }
    __syncthreads();
    // means written through the aliased centroid rows, in centroid order (kmeans.go:140-146)
    for (uint32_t jd = tid; jd < a.sub; jd += KM_THREADS) {
      for (uint32_t i = 0; i < a.K; ++i) {
        uint32_t c = counts[i];
        if (c == 0) continue;
        X(crow[i])[jd] = __fdiv_rn(sums[size_t(i) * a.sub + jd], float(c));
      }
    }
    __syncthreads();
  }
  if (tid == 0) a.iters[m] = iter;
  (void)s_pick;
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
      (rc = d_counts.ensure(size_t(M) * K * 2 + M))) {
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
  ka.sub = sub; ka.K = K; ka.max_iter = 100;  // product.go:209
  ka.first = uint32_t(pq_first_row);
  ka.labels = d_labels.p; ka.min_dist = d_md.p; ka.sums = d_sums.p;
  ka.counts = d_counts.p; ka.cent_row = d_counts.p + size_t(M) * K; ka.iters = d_counts.p + size_t(M) * K * 2;
  size_t full = size_t(K) * sub * sizeof(float);
  ka.stage_centroids = full <= 96 * 1024;
  size_t smem = ka.stage_centroids ? full : size_t(sub) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(kmeans_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  if (e != cudaSuccess) { cleanup(); return cuda_fail(e, "cudaFuncSetAttribute(kmeans)"); }
  kmeans_kernel<<<M, KM_THREADS, smem, st>>>(ka);
  ix->launches++;
  if ((e = cudaGetLastError()) != cudaSuccess) { cleanup(); return cuda_fail(e, "kmeans_kernel"); }
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
