// search.cu — dispatch of the beam-search kernel variants (search.cuh).
#include "search.cuh"

#include "index.cuh"

#include <cstdio>
#include <cstdlib>

namespace sdb {

namespace {

template <int KIND, int METRIC, int TRIPS, int SETS, bool LEGACY, int MERGE_MIN, class VT, bool FILTER, bool RETRY, int MINB,
          bool XTRA>
int launch_variant_x(sdb_index* ix, const SearchArgs& a, cudaStream_t stream) {
  auto kern = beam_search_kernel<KIND, METRIC, TRIPS, SETS, LEGACY, MERGE_MIN, VT, FILTER, RETRY, MINB, XTRA>;
  // FloatEvalGrouped keeps short queries (<= 4 float4 per lane) in registers: no shared copy
  constexpr bool QREG = (KIND == EVAL_FLOAT_FIXED) && !LEGACY && TRIPS <= 4;
  const uint32_t qfloats = (KIND == EVAL_ADC || KIND == EVAL_ADC_SMEM || QREG) ? 0 : (a.dim + 3) / 4 * 4;
  const uint32_t qwords = (KIND == EVAL_BITS) ? a.bits_pitch : 0;
  const uint32_t table_floats = (KIND == EVAL_ADC_SMEM) ? a.pqM * a.pqK : 0;
  const size_t smem = warp_smem_bytes<VT, FILTER>(qfloats, qwords, a.vt_slots, table_floats);
  static thread_local int cached_dev = -1;
  static thread_local size_t cached_smem = 0;
  static thread_local int ctas_per_sm = 0;
  if (cached_dev != ix->device || cached_smem != smem) {
    if (smem > ix->smem_optin) return fail(SDB_ERR_INTERNAL, "beam search kernel does not fit in shared memory");
    SDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    int nb = 0;
    SDB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, 32, smem));
    if (nb <= 0) return fail(SDB_ERR_INTERNAL, "beam search kernel cannot be resident");
    if (MINB > 1 && nb > MINB) nb = MINB;
    ctas_per_sm = nb;
    // The unified L1/shared array is also where in-flight global loads land: ask for no more
    // shared memory than the resident CTAs need so the rest stays L1 (more rows in flight).
    const size_t need_smem = size_t(nb) * (smem + 1024);
    int pct = int((need_smem * 100 + ix->smem_per_sm - 1) / ix->smem_per_sm);
    if (const char* e = getenv("SDB_K1_CARVEOUT")) pct = atoi(e);
    if (pct > 100) pct = 100;
    SDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    if (getenv("SDB_DEBUG_OCC"))
      fprintf(stderr, "[sdb] beam_search variant: smem %zu B/CTA, %d CTAs/SM, carveout %d%% (%zu KB needed)\n", smem, nb, pct,
              need_smem >> 10);
    cached_smem = smem;
    cached_dev = ix->device;
  }
  uint32_t resident = uint32_t(ix->sm_count) * ctas_per_sm;
  uint32_t need = RETRY ? uint32_t(ix->sm_count) : a.B;
  uint32_t grid = need < resident ? need : resident;
  if (grid == 0) grid = 1;
  if (!RETRY && (a.flags & 2u) && a.B > grid) {
    // equal number of queries per resident warp: no half-empty last wave
    const uint32_t waves = (a.B + grid - 1) / grid;
    grid = (a.B + waves - 1) / waves;
  }
  kern<<<grid, 32, smem, stream>>>(a, qfloats, qwords);
  ix->launches++;
  SDB_CUDA(cudaGetLastError());
  return SDB_OK;
}

// the start node's overflow edges (after deletes; normally none) get their own instantiation so
// the common case pays nothing for them (the extra loop cost 5 % on C2)
template <int KIND, int METRIC, int TRIPS, int SETS, bool LEGACY, int MERGE_MIN, class VT, bool FILTER, bool RETRY, int MINB>
int launch_variant(sdb_index* ix, const SearchArgs& a, cudaStream_t stream) {
  if (a.n_start_extra != 0)
    return launch_variant_x<KIND, METRIC, TRIPS, SETS, LEGACY, MERGE_MIN, VT, FILTER, RETRY, MINB, true>(ix, a, stream);
  return launch_variant_x<KIND, METRIC, TRIPS, SETS, LEGACY, MERGE_MIN, VT, FILTER, RETRY, MINB, false>(ix, a, stream);
}

template <int KIND, int METRIC, int TRIPS, int SETS, bool LEGACY, int MERGE_MIN, bool FILTER, int MINB, class VT = VisitedCompactN>
int launch_with_retry(sdb_index* ix, SearchArgs a, cudaStream_t stream) {
  int rc = launch_variant<KIND, METRIC, TRIPS, SETS, LEGACY, MERGE_MIN, VT, FILTER, false, MINB>(ix, a, stream);
  if (rc) return rc;
  // second pass over queries whose visited set overflowed (normally none): u32 table, 32768 slots
  a.work_counter = a.work_counter + 2;
  // same evaluator as the first pass (a re-run query walks ~80 hops alone on its SM: the
  // pipelined row gather is what keeps that under a millisecond); the ADC table goes back to
  // global memory because the u32 visited table takes the shared memory
  constexpr int RK = (KIND == EVAL_ADC_SMEM) ? EVAL_ADC : KIND;
  constexpr int RT = (KIND == EVAL_BITS || KIND == EVAL_FLOAT_FIXED) ? TRIPS : 1;  // rows must still be covered
  constexpr int RS = (KIND == EVAL_BITS || KIND == EVAL_FLOAT_FIXED) ? SETS : 1;
  if (getenv("SDB_DEBUG_RETRY")) {
    uint32_t h[2] = {0, 0};
    cudaStreamSynchronize(stream);
    cudaMemcpy(h, a.work_counter - 2, sizeof(h), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[sdb] beam search: %u of %u queries overflowed the compact visited table -> retry launch\n", h[1], a.B);
  }
  return launch_variant<RK, METRIC, RT, RS, false, 0, VisitedTable<15>, FILTER, true, 1>(ix, a, stream);
}

// Tuning knob for A/B runs on the GPU (not part of the ABI): SDB_K1_VARIANT picks the
// evaluator layout / list-update form of the dim-128 kernel.
int k1_variant() {
  const char* e = getenv("SDB_K1_VARIANT");
  return e ? atoi(e) : -1;
}

template <int METRIC>
int launch_float(sdb_index* ix, const SearchArgs& a, bool filtered, cudaStream_t stream) {
  if (filtered) return launch_with_retry<EVAL_FLOAT_GENERIC, METRIC, 1, 1, false, 0, true, 1>(ix, a, stream);
  if (a.dim % 32 == 0) {
    switch (a.dim / 32) {
      case 4:
        switch (k1_variant()) {
          // round-1 baseline (unpipelined evaluator, sequential list update, 8192-slot table)
          case 0: return launch_with_retry<EVAL_FLOAT_FIXED, METRIC, 4, 8, true, 0, false, 12, VisitedCompact>(ix, a, stream);
          case 1: return launch_with_retry<EVAL_FLOAT_FIXED, METRIC, 4, 6, false, 0, false, 12>(ix, a, stream);
          case 2: return launch_with_retry<EVAL_FLOAT_FIXED, METRIC, 4, 5, false, 2, false, 12, VisitedCompactN>(ix, a, stream);
          default: return launch_with_retry<EVAL_FLOAT_FIXED, METRIC, 4, 6, false, 2, false, 12>(ix, a, stream);
        }
      case 8: return launch_with_retry<EVAL_FLOAT_FIXED, METRIC, 8, 3, false, 2, false, 12>(ix, a, stream);
      case 12: return launch_with_retry<EVAL_FLOAT_FIXED, METRIC, 12, 2, false, 2, false, 12>(ix, a, stream);
      case 24: return launch_with_retry<EVAL_FLOAT_FIXED, METRIC, 24, 1, false, 2, false, 12>(ix, a, stream);
      default: break;
    }
  }
  return launch_with_retry<EVAL_FLOAT_GENERIC, METRIC, 1, 1, false, 2, false, 12>(ix, a, stream);
}

}  // namespace

int launch_search(sdb_index* ix, uint32_t B, const float* d_queries, uint32_t k, uint32_t L, uint64_t* d_out_ids,
                  float* d_out_dists, uint32_t* d_out_counts, uint32_t* d_vis_ids, float* d_vis_dists,
                  uint32_t* d_vis_len, uint32_t vis_cap, const uint32_t* d_filter_seed, uint32_t n_filter_seed,
                  const uint32_t* d_filter_bits, cudaStream_t stream) {
  // Visited-table size for this launch: start at 5888 slots; if the previous search on this
  // handle sent more than 0.1 % of its queries to the RETRY launch (they visited more nodes than
  // 87.5 % of the table), step up. Read only when the stream is idle: never adds a sync.
  if (ix->retry_check_pending && cudaStreamQuery(ix->last_search_stream) == cudaSuccess) {
    uint32_t h_retry = 0;
    SDB_CUDA(cudaMemcpy(&h_retry, ix->d_work.p + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    // A re-run query costs about as much as a whole first-pass wave (it walks its ~80 hops
    // alone on an SM), so a handful per batch already outweighs the ~5-10 % the bigger table
    // costs everybody: step up above 0.1 % of the batch (or any at all in a small batch).
    if ((uint64_t(h_retry) * 1000 > ix->last_B || (h_retry > 0 && ix->last_B <= 2000)) && ix->vt_level < 2) ix->vt_level++;
    ix->retry_check_pending = false;
  }
  int rc;
  if ((rc = ix->d_hops.ensure(B))) return rc;
  if ((rc = ix->d_ndist.ensure(B))) return rc;
  if ((rc = ix->d_work.ensure(4 + size_t(B)))) return rc;
  static const uint32_t kSlots[3] = {5888, 8192, 12288};
  SDB_CUDA(cudaMemsetAsync(ix->d_work.p, 0, 4 * sizeof(uint32_t), stream));
  SearchArgs a{};
  a.vt_slots = kSlots[ix->vt_level];
  if (const char* e = getenv("SDB_VT_SLOTS")) a.vt_slots = uint32_t(atoi(e)) / 8 * 8;
  a.vec = ix->d_vec;
  a.vec_pitch = ix->vec_pitch;
  a.bits = ix->d_bits;
  a.bits_pitch = ix->bits_pitch;
  a.words = ix->words;
  a.codes = ix->d_codes;
  a.codes_pitch = ix->codes_pitch;
  a.pqM = ix->pqM;
  a.pqK = ix->pqK;
  a.bq_thr = ix->d_bq_thr;
  a.bit_metric = ix->bq_metric;
  a.adj = ix->d_adj;
  a.R = ix->p.degree_bound;
  a.start_extra = ix->d_start_extra.p;
  a.n_start_extra = uint32_t(ix->h_start_extra.size());
  a.rows = ix->rows;
  a.queries = d_queries;
  a.dim = ix->p.dim;
  a.B = B;
  a.L = L;
  a.k = k;
  a.out_ids = d_out_ids;
  a.out_dists = d_out_dists;
  a.out_counts = d_out_counts;
  a.out_hops = ix->d_hops.p;
  a.out_ndist = ix->d_ndist.p;
  a.vis_ids = d_vis_ids;
  a.vis_dists = d_vis_dists;
  a.vis_len = d_vis_len;
  a.vis_cap = vis_cap;
  a.filter_seed = d_filter_seed;
  a.n_filter_seed = n_filter_seed;
  a.filter_bits = d_filter_bits;
  {
    const char* e = getenv("SDB_K1_FLAGS");
    a.flags = e ? uint32_t(atoi(e)) : 0u;
  }
  if (ix->gather) a.pg = *ix->gather;
  a.work_counter = ix->d_work.p;
  a.retry_count = ix->d_work.p + 1;
  a.retry_list = ix->d_work.p + 4;
  ix->last_B = B;
  ix->retry_check_pending = true;
  ix->last_search_stream = stream;
  const bool filtered = d_filter_bits != nullptr;

  if (ix->p.quantizer == SDB_QUANT_BINARY && ix->bq_fitted) {
    // KIND = bits: METRIC = hamming / jaccard, TRIPS = 128-byte chunks per row, SETS = pipeline depth
    const bool jac = ix->bq_metric == SDB_METRIC_JACCARD;
    const uint32_t nch = (ix->bits_pitch + 15) / 16;
    if (filtered) {
      return jac ? launch_with_retry<EVAL_BITS, METRIC_JACCARD, 4, 2, false, 0, true, 1>(ix, a, stream)
                 : launch_with_retry<EVAL_BITS, METRIC_HAMMING, 4, 2, false, 0, true, 1>(ix, a, stream);
    }
    if (jac) {
      if (nch <= 1) return launch_with_retry<EVAL_BITS, METRIC_JACCARD, 1, 8, false, 2, false, 12>(ix, a, stream);
      return launch_with_retry<EVAL_BITS, METRIC_JACCARD, 4, 2, false, 2, false, 12>(ix, a, stream);
    }
    if (nch <= 1) return launch_with_retry<EVAL_BITS, METRIC_HAMMING, 1, 8, false, 2, false, 12>(ix, a, stream);
    if (nch <= 2) return launch_with_retry<EVAL_BITS, METRIC_HAMMING, 2, 4, false, 2, false, 12>(ix, a, stream);
    return launch_with_retry<EVAL_BITS, METRIC_HAMMING, 4, 2, false, 2, false, 12>(ix, a, stream);
  }
  if (ix->p.quantizer == SDB_QUANT_PRODUCT && ix->pq_fitted) {
    if ((rc = ix->d_adc.ensure(size_t(B) * ix->pqM * ix->pqK))) return rc;
    if ((rc = launch_adc_tables(ix, B, d_queries, ix->d_adc.p, stream))) return rc;
    a.adc = ix->d_adc.p;
    // ADC table in shared memory when at least two query-warps per SM can hold theirs (C4:
    // 96 x 256 x 4 B = 96 KB => exactly two); otherwise the table is read through L1/L2
    const size_t fixed = warp_smem_bytes<VisitedCompactN, false>(0, 0, 0, ix->pqM * ix->pqK) + 1024;
    const size_t room = ix->smem_per_sm / 2 > fixed ? (ix->smem_per_sm / 2 - fixed) / 2 : 0;  // 16-bit visited slots
    if (!filtered && room >= 4096 && !getenv("SDB_ADC_GLOBAL")) {
      if (a.vt_slots > room) a.vt_slots = uint32_t(room) / 8 * 8;
      const uint32_t nch = (ix->pqM + 15) / 16;
      if (nch <= 2) return launch_with_retry<EVAL_ADC_SMEM, 0, 2, 1, false, 2, false, 12>(ix, a, stream);
      if (nch <= 4) return launch_with_retry<EVAL_ADC_SMEM, 0, 4, 1, false, 2, false, 12>(ix, a, stream);
      if (nch <= 6) return launch_with_retry<EVAL_ADC_SMEM, 0, 6, 1, false, 2, false, 12>(ix, a, stream);
      return launch_with_retry<EVAL_ADC_SMEM, 0, 8, 1, false, 2, false, 12>(ix, a, stream);
    }
    return filtered ? launch_with_retry<EVAL_ADC, 0, 1, 1, false, 0, true, 1>(ix, a, stream)
                    : launch_with_retry<EVAL_ADC, 0, 1, 1, false, 2, false, 12>(ix, a, stream);
  }
  switch (ix->store_metric) {
    case SDB_METRIC_EUCLIDEAN: return launch_float<METRIC_EUCLIDEAN>(ix, a, filtered, stream);
    case SDB_METRIC_DOT: return launch_float<METRIC_DOT>(ix, a, filtered, stream);
    case SDB_METRIC_COSINE: return launch_float<METRIC_COSINE>(ix, a, filtered, stream);
    default: return fail(SDB_ERR_INVALID, "metric not supported by the Vamana search kernel");
  }
}

}  // namespace sdb

using namespace sdb;

extern "C" int sdb_search_batch_gather_device(sdb_index* ix, uint32_t B, const float* d_queries, uint32_t k,
                                              uint32_t search_size, uint64_t* d_out_ids, float* d_out_dists,
                                              uint32_t* d_out_counts, const sdb_peer_gather* pg, void* stream) {
  if (!ix || !pg) return fail(SDB_ERR_INVALID, "null argument");
  if (B == 0) return SDB_OK;
  if (!d_queries || !d_out_ids || !d_out_dists || !d_out_counts) return fail(SDB_ERR_INVALID, "null argument");
  if (pg->n_peers == 0 || pg->n_peers > SDB_MAX_PEERS || pg->shard >= (1u << 24))
    return fail(SDB_ERR_INVALID, "peer gather: n_peers must be 1..16");
  if (k == 0 || search_size < k) return fail(SDB_ERR_INVALID, "searchSize must be greater than or equal to k");
  PeerGather g{};
  g.n = pg->n_peers;
  g.shard = pg->shard;
  g.limit = pg->per_shard_limit == 0 || pg->per_shard_limit > k ? k : pg->per_shard_limit;
  g.tag = uint64_t(pg->shard) << 40;
  for (uint32_t p = 0; p < pg->n_peers; ++p) {
    if (!pg->ids[p] || !pg->dists[p] || !pg->counts[p]) return fail(SDB_ERR_INVALID, "peer gather: null peer buffer");
    g.ids[p] = pg->ids[p];
    g.dists[p] = pg->dists[p];
    g.counts[p] = pg->counts[p];
  }
  std::lock_guard<std::mutex> lk(ix->mu);
  SDB_CUDA(cudaSetDevice(ix->device));
  ix->gather = &g;
  int rc = launch_search(ix, B, d_queries, k, search_size, d_out_ids, d_out_dists, d_out_counts, nullptr, nullptr, nullptr, 0,
                         nullptr, 0, nullptr, static_cast<cudaStream_t>(stream));
  ix->gather = nullptr;
  return rc;
}

