// host_test.cpp — tests of the C++ host mirror (semadb_b200/host), written to read like the
// reference's own Go tests of the same surfaces:
//   codec  (no GPU)  conversion/conversion_test.go:11-47, conversion/keys_test.go:10-19,
//                    models validation (models/index.go:284-313, models/search.go:277-306),
//                    diskstore/memstore semantics
//   gpu              shard/index/vamana/vamana_test.go: Test_Insert (29-46), Test_InsertInvalidIds
//                    (77-90), Test_ConcurrentCUD (92-140), Test_Flush (177-211), Test_EmptySearch
//                    (213-228), Test_Search (230-252), Test_SearchFilter (254-276), plus
//                    persistence (shard_vector_test.go:422-459: reopen from the bucket) and the
//                    query coalescer (many threads, one batch).
// usage: host_test codec | host_test gpu
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <queue>
#include <random>
#include <set>
#include <thread>

#include "../../semadb_b200/host/coalescer.hpp"
#include "../../semadb_b200/host/gpuvamana.hpp"
#include "../../semadb_b200/host/search.hpp"

using namespace semadb;
using vamana::IndexVamana;
using vamana::IndexVectorChange;

static int g_fail = 0;
#define CHECK(cond)                                                         \
  do {                                                                      \
    if (!(cond)) {                                                          \
      std::fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); \
      ++g_fail;                                                             \
    }                                                                       \
  } while (0)
#define REQUIRE_OK(expr)                                                                    \
  do {                                                                                      \
    Error _e = (expr);                                                                      \
    if (_e) {                                                                               \
      std::fprintf(stderr, "FAIL %s:%d: %s -> %s\n", __FILE__, __LINE__, #expr, _e.msg.c_str()); \
      ++g_fail;                                                                             \
      return;                                                                               \
    }                                                                                       \
  } while (0)

static models::IndexVectorVamanaParameters vamanaParams(unsigned dim = 2) {  // vamana_test.go:21-27
  models::IndexVectorVamanaParameters p;
  p.VectorSize = dim;
  p.DistanceMetric = models::DistanceEuclidean;
  p.SearchSize = 75;
  p.DegreeBound = 64;
  p.Alpha = 1.2f;
  return p;
}

static std::vector<IndexVectorChange> randPoints(int size, int offset, unsigned dim, std::mt19937& rng) {  // vamana_test.go:48-61
  std::uniform_real_distribution<float> u(0.0f, 1.0f);
  std::vector<IndexVectorChange> out(size);
  for (int i = 0; i < size; ++i) {
    out[i].Id = uint64_t(i + offset + 2);  // 0 is invalid, 1 is the start node
    out[i].Vector.resize(dim);
    for (unsigned d = 0; d < dim; ++d) out[i].Vector[d] = u(rng);
  }
  return out;
}

// checkConnectivity (vamana_test.go:63-75): BFS over the persisted n<id>e lists from node 1
static size_t reachable(const diskstore::MemBucket& b) {
  std::set<uint64_t> seen{1};
  std::queue<uint64_t> q;
  q.push(1);
  while (!q.empty()) {
    uint64_t v = q.front();
    q.pop();
    std::string val;
    if (!b.Get(conversion::NodeKey(v, 'e'), &val)) continue;
    for (uint64_t u : conversion::BytesToEdgeList(val))
      if (seen.insert(u).second) q.push(u);
  }
  return seen.size();
}

static void test_codec() {
  // conversion_test.go:11-17, keys_test.go:10-19
  CHECK(conversion::BytesToUint64(conversion::Uint64ToBytes(0x0123456789ABCDEFull)) == 0x0123456789ABCDEFull);
  CHECK(conversion::Uint64ToBytes(1) == std::string("\x01\0\0\0\0\0\0\0", 8));
  std::string key = conversion::NodeKey(42, 'a');
  uint64_t id = 0;
  CHECK(key.size() == 10 && key[0] == 'n' && uint8_t(key[1]) == 42 && key[9] == 'a');
  CHECK(conversion::NodeIdFromKey(key, 'a', &id) && id == 42);
  CHECK(!conversion::NodeIdFromKey(key, 'b', &id));
  CHECK(!conversion::NodeIdFromKey("n123", 'a', &id));
  // conversion_test.go:27-47: float32 slices round trip bit for bit; 1.0f = 00 00 80 3f
  std::mt19937 rng(7);
  std::uniform_real_distribution<float> u(0.0f, 1.0f);
  for (int n : {128, 768, 960, 1536}) {
    std::vector<float> f(n);
    for (float& x : f) x = u(rng);
    std::string b = conversion::Float32ToBytes(f);
    CHECK(b.size() == size_t(n) * 4);
    CHECK(conversion::BytesToFloat32(b) == f);
  }
  CHECK(conversion::Float32ToBytes(std::vector<float>{1.0f}) == std::string("\x00\x00\x80\x3f", 4));
  std::vector<uint64_t> edges{2, 3, 0xFFFFFFFFFFull};
  CHECK(conversion::BytesToEdgeList(conversion::EdgeListToBytes(edges)) == edges);
  CHECK(conversion::EdgeListToBytes(std::vector<uint64_t>{258}) == std::string("\x02\x01\0\0\0\0\0\0", 8));
  // memstore.go: read-only buckets refuse writes; prefix scan
  diskstore::MemBucket mb;
  CHECK(!mb.Put("nab", "1") && !mb.Put("nac", "2") && !mb.Put("x", "3"));
  int seen = 0;
  mb.PrefixScan("na", [&](const std::string&, const std::string&) { ++seen; return Ok(); });
  CHECK(seen == 2);
  mb.SetReadOnly(true);
  CHECK(bool(mb.Put("k", "v")) && bool(mb.Delete("x")));
  // models/index.go:284-313, models/search.go:277-306, models/quantizer.go:41-76
  auto p = vamanaParams();
  CHECK(!p.Validate());
  p.SearchSize = 76;
  CHECK(p.Validate().msg == "search size must be between 25 and 75, got 76");
  p = vamanaParams();
  p.DegreeBound = 31;
  CHECK(p.Validate().msg == "degree bound must be between 32 and 64, got 31");
  p = vamanaParams(3);
  p.DistanceMetric = models::DistanceHaversine;
  CHECK(p.Validate().msg == "haversine distance metric requires vector size 2 got 3");
  p = vamanaParams();
  p.DistanceMetric = "manhattan";
  CHECK(p.Validate().msg == "unknown distance metric manhattan");
  p = vamanaParams();
  p.Quantizer_ = models::Quantizer{models::QuantizerProduct, std::nullopt, models::ProductQuantizerParameters{300, 8, 5000}};
  CHECK(p.Validate().msg == "numCentroids must be between 2 and 256, got 300");
  models::SearchVectorVamanaOptions o;
  o.Vector = {1, 2};
  o.SearchSize = 25;
  o.Limit = 30;
  CHECK(o.Validate().msg == "searchSize must be greater than or equal to limit");
  o.Limit = 76;
  CHECK(o.Validate().msg == "invalid limit 76 for vector query, expected 1-75");
}

static void test_insert_search_flush() {
  std::mt19937 rng(1);
  for (int size : {1, 100, 4242}) {  // Test_Insert (vamana_test.go:29-46)
    diskstore::MemBucket bucket;
    std::unique_ptr<IndexVamana> inv;
    REQUIRE_OK(IndexVamana::New("test", vamanaParams(), &bucket, &inv, 0, 11));
    REQUIRE_OK(inv->InsertUpdateDelete(randPoints(size, 0, 2, rng)));
    CHECK(reachable(bucket) == size_t(size) + 1);
  }
  {  // Test_InsertInvalidIds (vamana_test.go:77-90)
    diskstore::MemBucket bucket;
    std::unique_ptr<IndexVamana> inv;
    REQUIRE_OK(IndexVamana::New("test", vamanaParams(), &bucket, &inv, 0, 11));
    for (uint64_t bad : {uint64_t(0), uint64_t(1)}) {
      IndexVectorChange c;
      c.Id = bad;
      c.Vector = {0.5f, 0.5f};
      CHECK(bool(inv->InsertUpdateDelete({c})));
    }
  }
  {  // Test_Flush (vamana_test.go:177-211): 42 points -> 43 'v' keys, 43 'e' keys, maxId 43
    diskstore::MemBucket bucket;
    std::unique_ptr<IndexVamana> inv;
    REQUIRE_OK(IndexVamana::New("test", vamanaParams(), &bucket, &inv, 0, 11));
    REQUIRE_OK(inv->InsertUpdateDelete(randPoints(42, 0, 2, rng)));
    int vec = 0, edge = 0, other = 0;
    uint64_t max_id = 0;
    bucket.ForEach([&](const std::string& k, const std::string& v) {
      if (k == vamana::MAXNODEIDKEY) max_id = conversion::BytesToUint64(v);
      else if (k.back() == 'v') ++vec;
      else if (k.back() == 'e') ++edge;
      else ++other;
      return Ok();
    });
    CHECK(vec == 43 && edge == 43 && other == 0 && max_id == 43);
  }
  {  // Test_EmptySearch (vamana_test.go:213-228)
    diskstore::MemBucket bucket;
    std::unique_ptr<IndexVamana> inv;
    REQUIRE_OK(IndexVamana::New("test", vamanaParams(), &bucket, &inv, 0, 11));
    models::SearchVectorVamanaOptions o;
    o.Vector = {0.5f, 0.5f};
    std::vector<uint64_t> rs;
    std::vector<models::SearchResult> res;
    REQUIRE_OK(inv->Search(o, nullptr, &rs, &res));
    CHECK(rs.empty() && res.empty());
  }
  {  // Test_Search (230-252), Test_SearchFilter (254-276), reopen from the bucket
    diskstore::MemBucket bucket;
    std::unique_ptr<IndexVamana> inv;
    REQUIRE_OK(IndexVamana::New("test", vamanaParams(), &bucket, &inv, 0, 11));
    auto rps = randPoints(200, 0, 2, rng);
    REQUIRE_OK(inv->InsertUpdateDelete(rps));
    auto self_search = [&](IndexVamana* ix) {
      for (const auto& rp : rps) {
        models::SearchVectorVamanaOptions o;
        o.Vector = rp.Vector;
        std::vector<uint64_t> rs;
        std::vector<models::SearchResult> res;
        REQUIRE_OK(ix->Search(o, nullptr, &rs, &res));
        CHECK(rs.size() == 10 && res.size() == 10);
        CHECK(res[0].NodeId == rp.Id && res[0].Distance == 0.0f && res[0].HybridScore == -0.0f);
      }
    };
    self_search(inv.get());
    std::vector<uint64_t> filter{rps[0].Id, rps[1].Id, rps[2].Id};
    models::SearchVectorVamanaOptions o;
    o.Vector = rps[0].Vector;
    std::vector<uint64_t> rs;
    std::vector<models::SearchResult> res;
    REQUIRE_OK(inv->Search(o, &filter, &rs, &res));
    CHECK(rs.size() == 3 && res.size() == 3 && res[0].NodeId == rps[0].Id);
    // a non-nil EMPTY bitmap is still a filter: nothing is seeded, nothing passes (search.go:33-51,93-95)
    std::vector<uint64_t> empty_filter;
    REQUIRE_OK(inv->Search(o, &empty_filter, &rs, &res));
    CHECK(rs.empty() && res.empty());
    // a filter naming a point that does not exist fails like GetMany (search.go:45-48)
    std::vector<uint64_t> missing{rps[0].Id, 999999};
    CHECK(bool(inv->Search(o, &missing, &rs, &res)));
    o.SearchSize = 5;  // search.go:23-25
    o.Limit = 10;
    CHECK(bool(inv->Search(o, nullptr, &rs, &res)));
    // a second index over the same bucket sees the same graph (shard_vector_test.go:444-459)
    const size_t keys_before = bucket.Size();
    std::unique_ptr<IndexVamana> again;
    REQUIRE_OK(IndexVamana::New("test", vamanaParams(), &bucket, &again, 0, 99));
    CHECK(again->MaxNodeId() == 201);
    self_search(again.get());
    REQUIRE_OK(again->Flush());
    CHECK(bucket.Size() == keys_before);  // nothing was dirty
  }
}

static void test_cud_and_persistence() {  // Test_ConcurrentCUD (vamana_test.go:92-140)
  std::mt19937 rng(2);
  diskstore::MemBucket bucket;
  std::unique_ptr<IndexVamana> inv;
  REQUIRE_OK(IndexVamana::New("test", vamanaParams(), &bucket, &inv, 0, 5));
  REQUIRE_OK(inv->InsertUpdateDelete(randPoints(50, 0, 2, rng)));
  std::vector<IndexVectorChange> ch = randPoints(50, 50, 2, rng);  // insert more
  auto upd = randPoints(25, 25, 2, rng);                            // update some
  ch.insert(ch.end(), upd.begin(), upd.end());
  for (int i = 0; i < 25; ++i) {                                    // delete some
    IndexVectorChange c;
    c.Id = uint64_t(i + 2);
    ch.push_back(c);
  }
  REQUIRE_OK(inv->InsertUpdateDelete(ch));
  CHECK(reachable(bucket) == 76);  // 75 points + start node
  for (int i = 0; i < 25; ++i) {
    CHECK(!bucket.Get(conversion::NodeKey(uint64_t(i + 2), 'v'), nullptr));
    CHECK(!bucket.Get(conversion::NodeKey(uint64_t(i + 2), 'e'), nullptr));
  }
  // no edge points at a deleted node, no self edges (shard_vector_test.go:198-225)
  bucket.ForEach([&](const std::string& k, const std::string& v) {
    uint64_t id = 0;
    if (!conversion::NodeIdFromKey(k, 'e', &id)) return Ok();
    for (uint64_t e : conversion::BytesToEdgeList(v)) CHECK(e != id && !(e >= 2 && e < 27));
    return Ok();
  });
  // updated points are found at their new vectors after a reopen
  std::unique_ptr<IndexVamana> again;
  REQUIRE_OK(IndexVamana::New("test", vamanaParams(), &bucket, &again, 0, 6));
  for (const auto& rp : upd) {
    models::SearchVectorVamanaOptions o;
    o.Vector = rp.Vector;
    std::vector<uint64_t> rs;
    std::vector<models::SearchResult> res;
    REQUIRE_OK(again->Search(o, nullptr, &rs, &res));
    CHECK(!res.empty() && res[0].NodeId == rp.Id && res[0].Distance == 0.0f);
  }
}

static void test_quantized_persistence() {
  // hamming metric: binary store with threshold 0.5 (vectorstore.go:56-66); only n<id>q is
  // written (binary.go:298-309) and a reopened index serves searches from the codes alone
  std::mt19937 rng(3);
  auto p = vamanaParams(128);
  p.DistanceMetric = models::DistanceHamming;
  diskstore::MemBucket bucket;
  std::unique_ptr<IndexVamana> inv;
  REQUIRE_OK(IndexVamana::New("bits", p, &bucket, &inv, 0, 5));
  std::vector<IndexVectorChange> pts(600);
  std::bernoulli_distribution bit(0.5);
  for (size_t i = 0; i < pts.size(); ++i) {
    pts[i].Id = i + 2;
    pts[i].Vector.resize(128);
    for (float& x : pts[i].Vector) x = bit(rng) ? 1.0f : 0.0f;
  }
  REQUIRE_OK(inv->InsertUpdateDelete(pts));
  CHECK(bucket.Get(conversion::NodeKey(2, 'q'), nullptr) && !bucket.Get(conversion::NodeKey(2, 'v'), nullptr));
  CHECK(bucket.Get(vamana::binaryQuantizerThresholdKey, nullptr));
  std::unique_ptr<IndexVamana> again;
  REQUIRE_OK(IndexVamana::New("bits", p, &bucket, &again, 0, 6));
  for (int i = 0; i < 50; ++i) {
    models::SearchVectorVamanaOptions o;
    o.Vector = pts[i].Vector;
    std::vector<uint64_t> rs, rs2;
    std::vector<models::SearchResult> a, b;
    REQUIRE_OK(inv->Search(o, nullptr, &rs, &a));
    REQUIRE_OK(again->Search(o, nullptr, &rs2, &b));
    CHECK(a.size() == b.size() && !a.empty() && a[0].Distance == 0.0f);
    for (size_t j = 0; j < a.size() && j < b.size(); ++j) CHECK(a[j].NodeId == b[j].NodeId && a[j].Distance == b[j].Distance);
  }
}

static void test_coalescer() {
  std::mt19937 rng(4);
  diskstore::MemBucket bucket;
  std::unique_ptr<IndexVamana> inv;
  REQUIRE_OK(IndexVamana::New("test", vamanaParams(16), &bucket, &inv, 0, 5));
  auto pts = randPoints(5000, 0, 16, rng);
  REQUIRE_OK(inv->InsertUpdateDelete(pts));
  // ground truth: one direct batch
  const uint32_t B = 2000, K = 10;
  std::vector<float> q(size_t(B) * 16);
  for (uint32_t b = 0; b < B; ++b) std::copy(pts[b].Vector.begin(), pts[b].Vector.end(), q.begin() + size_t(b) * 16);
  std::vector<uint64_t> ids(size_t(B) * K);
  std::vector<float> d(size_t(B) * K);
  std::vector<uint32_t> cnt(B);
  REQUIRE_OK(inv->SearchBatch(q.data(), B, K, 75, ids.data(), d.data(), cnt.data()));
  vamana::SearchCoalescer co(inv.get(), 512, std::chrono::microseconds(500));
  std::atomic<int> bad{0};
  std::vector<std::thread> th;
  for (int t = 0; t < 32; ++t)
    th.emplace_back([&, t] {
      for (uint32_t b = t; b < B; b += 32) {
        models::SearchVectorVamanaOptions o;
        o.Vector = pts[b].Vector;
        o.Limit = K;
        std::vector<uint64_t> rs;
        std::vector<models::SearchResult> res;
        Error e = co.Search(o, &rs, &res);
        if (e || res.size() != cnt[b]) { ++bad; continue; }
        for (uint32_t j = 0; j < cnt[b]; ++j)
          if (res[j].NodeId != ids[size_t(b) * K + j] || res[j].Distance != d[size_t(b) * K + j]) ++bad;
      }
    });
  for (auto& x : th) x.join();
  CHECK(bad.load() == 0);
  CHECK(co.queries() == B);
  CHECK(co.batches() < B / 4);  // requests really were coalesced
  // per-request filters through the coalescer (shard/index/search.go:59-85): every third request
  // carries its own bitmap, every 30th an empty one, one names a missing point; each result must
  // equal the same request served alone
  {
    const uint32_t B2 = 600;
    std::vector<std::vector<uint64_t>> filt(B2);
    for (uint32_t b = 0; b < B2; ++b) {
      if (b % 3) continue;
      if (b % 30 == 0) continue;  // stays empty (but is passed as a filter)
      for (uint64_t j = 0; j < 40 + b % 50; ++j) filt[b].push_back(pts[(b * 7 + j * 13) % pts.size()].Id);
    }
    filt[33].push_back(1u << 30);  // missing point: this request alone must fail
    std::vector<std::vector<models::SearchResult>> want(B2);
    std::vector<int> want_err(B2, 0);
    for (uint32_t b = 0; b < B2; ++b) {
      models::SearchVectorVamanaOptions o;
      o.Vector = pts[b].Vector;
      o.Limit = K;
      std::vector<uint64_t> rs;
      want_err[b] = bool(inv->Search(o, b % 3 == 0 ? &filt[b] : nullptr, &rs, &want[b]));
    }
    CHECK(want_err[33] == 1 && want_err[30] == 0 && want[30].empty() && want[0].empty() && !want[3].empty());
    std::atomic<int> bad2{0};
    std::vector<std::thread> th2;
    for (int t = 0; t < 32; ++t)
      th2.emplace_back([&, t] {
        for (uint32_t b = t; b < B2; b += 32) {
          models::SearchVectorVamanaOptions o;
          o.Vector = pts[b].Vector;
          o.Limit = K;
          std::vector<uint64_t> rs;
          std::vector<models::SearchResult> res;
          Error e = co.Search(o, b % 3 == 0 ? &filt[b] : nullptr, &rs, &res);
          if (bool(e) != bool(want_err[b])) { ++bad2; continue; }
          if (e) continue;
          if (res.size() != want[b].size()) { ++bad2; continue; }
          for (size_t j = 0; j < res.size(); ++j)
            if (res[j].NodeId != want[b][j].NodeId || res[j].Distance != want[b][j].Distance) ++bad2;
        }
      });
    for (auto& x : th2) x.join();
    CHECK(bad2.load() == 0);
  }
  std::fprintf(stderr, "coalescer: %llu queries in %llu batches\n", (unsigned long long)co.queries(),
               (unsigned long long)co.batches());
}

// TestSearch_OrVector / TestSearch_And (shard/index/search_test.go:289-457) over the GPU path: two
// vector searches of the same neighbourhood with weight 0.5 each merge into results whose
// HybridScore is -distance; `_and` keeps the common ids only.
static void test_search_parallel_merge() {
  diskstore::MemBucket bucket;
  std::unique_ptr<IndexVamana> inv;
  REQUIRE_OK(IndexVamana::New("test", vamanaParams(2), &bucket, &inv, 0, 7));
  std::vector<IndexVectorChange> pts;
  for (uint64_t i = 0; i < 100; ++i) pts.push_back(IndexVectorChange{i + 2, {float(i), float(i + 1)}});
  REQUIRE_OK(inv->InsertUpdateDelete(pts));
  const float w = 0.5f;
  models::SearchVectorVamanaOptions o;
  o.Vector = {42.0f, 43.0f};
  o.Limit = 5;
  o.Weight = w;
  std::vector<uint64_t> s1, s2, fs;
  std::vector<models::SearchResult> r1, r2, fr;
  REQUIRE_OK(inv->Search(o, nullptr, &s1, &r1));
  REQUIRE_OK(inv->Search(o, nullptr, &s2, &r2));
  CHECK(r1.size() == 5);
  REQUIRE_OK(index::SearchParallelMerge({r1, r2}, true, 0, &fs, &fr));
  CHECK(fr.size() == 5 && fs.size() == 5);
  CHECK(!fr.empty() && fr[0].NodeId == 44);  // point 42 has node id 44
  for (size_t i = 0; i < fr.size(); ++i) {
    CHECK(fr[i].HybridScore == -fr[i].Distance);  // two weights of 0.5 add up
    if (i + 1 < fr.size()) CHECK(fr[i].HybridScore >= fr[i + 1].HybridScore);
  }
  o.Limit = 3;
  std::vector<uint64_t> s3;
  std::vector<models::SearchResult> r3;
  REQUIRE_OK(inv->Search(o, nullptr, &s3, &r3));
  REQUIRE_OK(index::SearchParallelMerge({r1, r3}, false, 0, &fs, &fr));
  CHECK(fr.size() == 3 && fs == std::vector<uint64_t>({43, 44, 45}));
  REQUIRE_OK(index::SearchParallelMerge({r3}, false, 0, &fs, &fr));  // single member: passed through
  CHECK(fr.size() == r3.size());
}

int main(int argc, char** argv) {
  const std::string mode = argc > 1 ? argv[1] : "codec";
  test_codec();
  if (mode == "gpu") {
    if (sdb_device_count() == 0) {
      std::fprintf(stderr, "no CUDA device: the GPU index has no CPU fallback\n");
      return 2;
    }
    test_insert_search_flush();
    test_cud_and_persistence();
    test_quantized_persistence();
    test_coalescer();
    test_search_parallel_merge();
  }
  if (g_fail) {
    std::fprintf(stderr, "%d check(s) failed\n", g_fail);
    return 1;
  }
  std::printf("host_test %s: ok\n", mode.c_str());
  return 0;
}
