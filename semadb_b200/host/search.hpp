// search.hpp — host mirror of indexManager.searchParallel's merge step
// (shard/index/search.go:246-298) over the C ABI (sdb_hybrid_merge): the members of an `_and` /
// `_or` query have been searched; their ranked lists are combined on the GPU — union or
// intersection of the result-id sets, HybridScore of duplicates added, sorted by HybridScore
// descending. Members that are not ranked searches (inverted-index sets) stay in Go.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <string>
#include <vector>

#include "../../include/semadb_b200.h"
#include "gpuvamana.hpp"
#include "models.hpp"

namespace semadb {
namespace index {

// member_results[i]: the SearchResult list of member query i (its result set = the ids of the list).
inline Error SearchParallelMerge(const std::vector<std::vector<models::SearchResult>>& member_results,
                                         bool is_disjunction, int device, std::vector<uint64_t>* final_set,
                                         std::vector<models::SearchResult>* final_results) {
  final_set->clear();
  final_results->clear();
  if (member_results.empty()) return Error();
  if (member_results.size() == 1) {  // shortcut, no merging required (search.go:246-249)
    *final_results = member_results[0];
    for (const auto& r : member_results[0]) final_set->push_back(r.NodeId);
    return Error();
  }
  const uint32_t S = uint32_t(member_results.size());
  size_t kmax = 1;
  for (const auto& m : member_results) kmax = std::max(kmax, m.size());
  const uint32_t k = uint32_t(kmax);
  std::vector<uint64_t> ids(size_t(S) * k, 0), out_ids(size_t(S) * k);
  std::vector<float> h(size_t(S) * k, 0.0f), d(size_t(S) * k, std::numeric_limits<float>::quiet_NaN());
  std::vector<float> out_h(size_t(S) * k), out_d(size_t(S) * k);
  std::vector<uint32_t> cnt(S);
  uint32_t out_cnt = 0;
  for (uint32_t s = 0; s < S; ++s) {
    cnt[s] = uint32_t(member_results[s].size());
    for (size_t r = 0; r < member_results[s].size(); ++r) {
      ids[size_t(s) * k + r] = member_results[s][r].NodeId;
      h[size_t(s) * k + r] = member_results[s][r].HybridScore;
      d[size_t(s) * k + r] = member_results[s][r].Distance;
    }
  }
  const int rc = sdb_hybrid_merge(device, S, 1, k, is_disjunction ? 1 : 0, ids.data(), h.data(), d.data(), cnt.data(),
                                  out_ids.data(), out_h.data(), out_d.data(), &out_cnt);
  if (rc != SDB_OK) return Error(std::string("parallel search failed: ") + sdb_last_error());
  for (uint32_t j = 0; j < out_cnt; ++j) {
    models::SearchResult r;
    r.NodeId = out_ids[j];
    r.Distance = out_d[j];
    r.HybridScore = out_h[j];
    final_results->push_back(r);
    final_set->push_back(out_ids[j]);
  }
  std::sort(final_set->begin(), final_set->end());
  return Error();
}

}  // namespace index
}  // namespace semadb
