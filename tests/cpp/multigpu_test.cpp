// multigpu_test.cpp — the cross-shard exchange driven the way INTEGRATION.md's Go sketch does it:
// ONE process, one shard per GPU, cudaDeviceEnablePeerAccess between the devices, and nothing but
// the C ABI (include/semadb_b200.h): sdb_search_batch_gather_device (each GPU's search stores its
// top-k straight into every peer's gather buffer), sdb_peer_barrier_device, sdb_merge_topk_device.
// Mirrors the fan-out / fan-in of ClusterNode.SearchPoints (cluster/actions.go:316-376): the merged
// lists must be identical on every GPU and equal to merging the per-shard results on the host.
// Exits 0 with "SKIP" when fewer than two GPUs are visible.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../../include/semadb_b200.h"

#define CK(x)                                                                     \
  do {                                                                            \
    cudaError_t e_ = (x);                                                         \
    if (e_ != cudaSuccess) { std::fprintf(stderr, "CUDA %s: %s\n", #x, cudaGetErrorString(e_)); return 2; } \
  } while (0)
#define SDB(x)                                                                  \
  do {                                                                          \
    int rc_ = (x);                                                              \
    if (rc_ != SDB_OK) { std::fprintf(stderr, "%s -> %d: %s\n", #x, rc_, sdb_last_error()); return 3; } \
  } while (0)

int main() {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 2) {
    std::printf("multigpu_test: SKIP (needs two GPUs, found %d)\n", ndev);
    return 0;
  }
  const int S = 2;
  const uint32_t dim = 32, n = 20000, B = 512, k = 10, L = 75;
  for (int a = 0; a < S; ++a)
    for (int b = 0; b < S; ++b)
      if (a != b) {
        int can = 0;
        CK(cudaDeviceCanAccessPeer(&can, a, b));
        if (!can) { std::printf("multigpu_test: SKIP (no peer access %d -> %d)\n", a, b); return 0; }
        CK(cudaSetDevice(a));
        cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e);
        cudaGetLastError();
      }
  std::mt19937 rng(7);
  std::normal_distribution<float> nd(0.f, 1.f);
  std::vector<float> q(size_t(B) * dim);
  for (auto& v : q) v = nd(rng);
  sdb_index* ix[S];
  float* d_q[S];
  uint64_t *g_ids[S], *l_ids[S], *m_ids[S];
  float *g_d[S], *l_d[S], *m_d[S];
  uint32_t *g_c[S], *l_c[S], *m_c[S], *flags[S];
  cudaStream_t st[S];
  std::vector<std::vector<uint64_t>> h_ids(S, std::vector<uint64_t>(size_t(B) * k));
  std::vector<std::vector<float>> h_d(S, std::vector<float>(size_t(B) * k));
  std::vector<std::vector<uint32_t>> h_c(S, std::vector<uint32_t>(B));
  for (int s = 0; s < S; ++s) {
    sdb_params p{};
    p.dim = dim; p.metric = SDB_METRIC_EUCLIDEAN; p.search_size = L; p.degree_bound = 64; p.alpha = 1.2f;
    p.quantizer = SDB_QUANT_NONE; p.bq_threshold = NAN; p.device = s;
    SDB(sdb_index_create(&p, &ix[s]));
    std::vector<float> start(dim, 0.f);
    start[0] = 1.f;
    SDB(sdb_index_set_start(ix[s], start.data()));
    std::vector<float> x(size_t(n) * dim);
    for (auto& v : x) v = nd(rng);
    std::vector<uint64_t> ids(n);
    for (uint32_t i = 0; i < n; ++i) ids[i] = i + 2;
    SDB(sdb_insert_batch(ix[s], n, ids.data(), x.data()));
    // the per-shard answer through the plain host entry point (what a shard returns on its own)
    SDB(sdb_search_batch(ix[s], B, q.data(), k, L, nullptr, 0, h_ids[s].data(), h_d[s].data(), h_c[s].data()));
    CK(cudaSetDevice(s));
    CK(cudaStreamCreate(&st[s]));
    CK(cudaMalloc(&d_q[s], q.size() * 4));
    CK(cudaMemcpy(d_q[s], q.data(), q.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&g_ids[s], size_t(S) * B * k * 8)); CK(cudaMalloc(&g_d[s], size_t(S) * B * k * 4)); CK(cudaMalloc(&g_c[s], size_t(S) * B * 4));
    CK(cudaMalloc(&l_ids[s], size_t(B) * k * 8)); CK(cudaMalloc(&l_d[s], size_t(B) * k * 4)); CK(cudaMalloc(&l_c[s], size_t(B) * 4));
    CK(cudaMalloc(&m_ids[s], size_t(B) * k * 8)); CK(cudaMalloc(&m_d[s], size_t(B) * k * 4)); CK(cudaMalloc(&m_c[s], size_t(B) * 4));
    CK(cudaMalloc(&flags[s], 2 * SDB_MAX_PEERS * 4));
    CK(cudaMemset(flags[s], 0, 2 * SDB_MAX_PEERS * 4));
    CK(cudaMemset(g_c[s], 0xFF, size_t(S) * B * 4));  // poison: every slot must be written by a peer
    CK(cudaDeviceSynchronize());
  }
  // two steps (epochs 1, 2): the second re-uses the buffers after the first merge has completed
  for (uint32_t epoch = 1; epoch <= 2; ++epoch) {
    for (int s = 0; s < S; ++s) {
      sdb_peer_gather pg{};
      pg.n_peers = S; pg.shard = uint32_t(s); pg.per_shard_limit = sdb_shard_limit(k, S, 75);
      for (int p = 0; p < S; ++p) { pg.ids[p] = g_ids[p]; pg.dists[p] = g_d[p]; pg.counts[p] = g_c[p]; }
      SDB(sdb_search_batch_gather_device(ix[s], B, d_q[s], k, L, l_ids[s], l_d[s], l_c[s], &pg, st[s]));
      SDB(sdb_peer_barrier_device(s, S, uint32_t(s), flags, epoch, st[s]));
      SDB(sdb_merge_topk_device(s, S, B, k, g_ids[s], g_d[s], g_c[s], m_ids[s], m_d[s], m_c[s], st[s]));
    }
    for (int s = 0; s < S; ++s) {
      CK(cudaSetDevice(s));
      CK(cudaStreamSynchronize(st[s]));
      SDB(sdb_peer_barrier_check(s, 0));
    }
    // host-side merge of the per-shard results with the shard tag of the fused path
    std::vector<uint64_t> in_ids(size_t(S) * B * k), want_ids(size_t(B) * k), got_ids(size_t(B) * k);
    std::vector<float> in_d(size_t(S) * B * k), want_d(size_t(B) * k), got_d(size_t(B) * k);
    std::vector<uint32_t> in_c(size_t(S) * B), want_c(B), got_c(B);
    for (int s = 0; s < S; ++s) {
      for (size_t i = 0; i < size_t(B) * k; ++i) {
        const uint64_t id = h_ids[s][i];
        in_ids[size_t(s) * B * k + i] = id ? (id | (uint64_t(s) << 40)) : 0;
        in_d[size_t(s) * B * k + i] = h_d[s][i];
      }
      for (uint32_t b = 0; b < B; ++b) in_c[size_t(s) * B + b] = h_c[s][b];
    }
    SDB(sdb_merge_topk(0, S, B, k, in_ids.data(), in_d.data(), in_c.data(), want_ids.data(), want_d.data(), want_c.data()));
    for (int s = 0; s < S; ++s) {
      CK(cudaSetDevice(s));
      CK(cudaMemcpy(got_ids.data(), m_ids[s], got_ids.size() * 8, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(got_d.data(), m_d[s], got_d.size() * 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(got_c.data(), m_c[s], got_c.size() * 4, cudaMemcpyDeviceToHost));
      for (uint32_t b = 0; b < B; ++b) {
        if (got_c[b] != want_c[b]) { std::fprintf(stderr, "epoch %u gpu %d query %u: count %u != %u\n", epoch, s, b, got_c[b], want_c[b]); return 1; }
        for (uint32_t j = 0; j < got_c[b]; ++j) {
          const size_t o = size_t(b) * k + j;
          if (got_ids[o] != want_ids[o] || std::memcmp(&got_d[o], &want_d[o], 4) != 0) {
            std::fprintf(stderr, "epoch %u gpu %d query %u rank %u: (%llx, %g) != (%llx, %g)\n", epoch, s, b, j,
                         (unsigned long long)got_ids[o], got_d[o], (unsigned long long)want_ids[o], want_d[o]);
            return 1;
          }
        }
      }
    }
  }
  for (int s = 0; s < S; ++s) sdb_index_destroy(ix[s]);
  std::printf("multigpu_test: ok (2 GPUs, one process, fused exchange == host merge of per-shard results, 2 epochs)\n");
  return 0;
}
