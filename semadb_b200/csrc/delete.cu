// delete.cu — the update/delete half of insertUpdateDelete (shard/index/vamana/vamana.go:
// 136-263): EdgeScan (node.go:142-199), pruneDeleteNeighbour (prune.go:12-84),
// removeInboundEdges (prune.go:88-154), row removal, re-insert of updated points.
//
//   1. edge_scan_kernel: one thread per adjacency slot, streaming the whole adjacency array
//      (HBM-bound, rows*R*4 bytes): marks hasInbound[target] for every edge of a valid
//      (existing, not in the delete set) node and flags the node toPrune if the target is in
//      the delete set. Delete set and hasInbound are bitmaps over rows (L2-resident).
//   2. cub select -> toPrune / toSave id lists in ascending id order (the reference iterates
//      a Go map; the oracle and this kernel fix ascending ids).
//   3. prune_delete_kernel: one CTA per toPrune node A. Candidates = A's surviving edges,
//      then the surviving edges of every deleted neighbour, in that order, deduplicated
//      keeping first occurrences (DistSet.Add with a VisitedMap), distances from A
//      (DistanceFromPoint), stable sort, then robustPrune if more than R candidates, else all
//      of them except A itself. Every CTA writes only its own row and reads only rows of
//      deleted nodes, so the CTAs are independent — same result as the reference's loop.
//   4. toSave ids are appended to the start node's edge list if absent (host side: a handful
//      of ids); edges beyond R go to the overflow list searched by the beam-search kernel.
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>

#include <algorithm>
#include <string>

#include "prune.cuh"

namespace sdb {

namespace {

__global__ void edge_scan_kernel(const uint32_t* __restrict__ adj, uint64_t n_slots, uint32_t R, uint32_t fixed_src,
                                 const uint8_t* __restrict__ exists, const uint32_t* __restrict__ delbits,
                                 uint32_t* __restrict__ inbound, uint8_t* __restrict__ prune_mark) {
  const uint64_t t = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= n_slots) return;
  const uint32_t e = adj[t];
  if (e == INVALID_ID) return;
  const uint32_t src = fixed_src != INVALID_ID ? fixed_src : uint32_t(t / R);
  if (!exists[src] || ((delbits[src >> 5] >> (src & 31)) & 1u)) return;
  atomicOr(&inbound[e >> 5], 1u << (e & 31));
  if ((delbits[e >> 5] >> (e & 31)) & 1u) prune_mark[src] = 1;
}

__global__ void save_mark_kernel(const uint8_t* __restrict__ exists, const uint32_t* __restrict__ delbits,
                                 const uint32_t* __restrict__ inbound, uint32_t rows, uint8_t* __restrict__ save_mark) {
  const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= rows) return;
  const bool del = (delbits[id >> 5] >> (id & 31)) & 1u;
  const bool in = (inbound[id >> 5] >> (id & 31)) & 1u;
  save_mark[id] = (exists[id] && !del && !in && id != START_ID) ? 1 : 0;
}

struct PruneDelArgs {
  StoreView s;
  uint32_t* adj; uint32_t* deg; uint8_t* dirty; uint32_t R; float alpha;
  const uint32_t* delbits;
  const uint32_t* prune_ids; uint32_t n_prune;
  const uint32_t* start_extra; uint32_t n_start_extra;
  int cap;  // candidate capacity carved per CTA
  uint32_t* error_flag;
};

__device__ __forceinline__ bool in_del(const uint32_t* delbits, uint32_t id) {
  return (delbits[id >> 5] >> (id & 31)) & 1u;
}

__global__ void __launch_bounds__(PRUNE_THREADS) prune_delete_kernel(PruneDelArgs a) {
  extern __shared__ __align__(16) unsigned char dyn_raw[];
  PruneShared sh;
  sh.carve(dyn_raw, a.cap);
  __shared__ int s_n;
  const int lane = threadIdx.x & 31;
  const int g = lane & 7;
  const int grp = threadIdx.x >> 3;
  for (uint32_t b = blockIdx.x; b < a.n_prune; b += gridDim.x) {
    const uint32_t A = a.prune_ids[b];
    __syncthreads();
    // ---- gather candidate ids in the reference's order (prune.go:24-57), single thread: the
    // lists are short (<= (R+extra)*(R+1)) and the order matters
    if (threadIdx.x == 0) {
      int n = 0, nexp = 0;
      const uint32_t dA = a.deg[A];
      const uint32_t nA = dA + (A == START_ID ? a.n_start_extra : 0u);
      // sh.sid doubles as the toExpand list while gathering
      for (uint32_t t = 0; t < nA; ++t) {
        const uint32_t e = t < dA ? a.adj[size_t(A) * a.R + t] : a.start_extra[t - dA];
        if (in_del(a.delbits, e)) sh.sid[nexp++] = e;
        else if (n < a.cap) sh.id[n++] = e;
      }
      for (int x = 0; x < nexp; ++x) {
        const uint32_t B = sh.sid[x];
        const uint32_t dB = a.deg[B];
        for (uint32_t t = 0; t < dB; ++t) {
          const uint32_t c = a.adj[size_t(B) * a.R + t];
          if (!in_del(a.delbits, c)) {
            if (n < a.cap) sh.id[n++] = c;
            else atomicExch(a.error_flag, 1u);
          }
        }
      }
      if (nexp == 0) atomicExch(a.error_flag, 2u);  // prune.go:37-40
      s_n = n;
    }
    __syncthreads();
    int n = s_n;
    // ---- dedupe keeping first occurrences (DistSet.Add, distset.go:203-212)
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const uint32_t v = sh.id[i];
      bool dup = false;
      for (int j = 0; j < i && !dup; ++j) dup = sh.id[j] == v;
      sh.removed[i] = dup ? 1 : 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int m = 0;
      for (int i = 0; i < n; ++i)
        if (!sh.removed[i]) sh.sid[m++] = sh.id[i];
      s_n = m;
    }
    __syncthreads();
    n = s_n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) sh.id[i] = sh.sid[i];
    __syncthreads();
    sh.n = n;
    // ---- distances from A (prune.go:59)
    const unsigned char* xa = global_row(a.s, A);
    for (int j0 = 0; j0 < n; j0 += PRUNE_GROUPS) {
      const int j = j0 + grp;
      const bool act = j < n;
      if (!__any_sync(SDB_FULL, act)) continue;
      const unsigned char* yj = global_row(a.s, sh.id[act ? j : 0]);
      const float d = row_dist(a.s, xa, yj, g);
      if (act && g == 0) sh.dist[j] = d;
    }
    __syncthreads();
    stable_sort_by_dist(sh);  // candidateSet.Sort() (prune.go:66)
    int cnt;
    if (n > int(a.R)) {
      robust_prune_cta(a.s, sh, nullptr, 0, A, int(a.R), a.alpha);  // prune.go:68-70
      cnt = *sh.cnt;
    } else {
      // room for every candidate except A itself (prune.go:72-81)
      if (threadIdx.x == 0) {
        int c = 0;
        for (int i = 0; i < n; ++i)
          if (sh.sid[i] != A) sh.edges[c++] = sh.sid[i];
        *sh.cnt = c;
      }
      __syncthreads();
      cnt = *sh.cnt;
    }
    for (uint32_t t = threadIdx.x; t < a.R; t += blockDim.x)
      a.adj[size_t(A) * a.R + t] = t < uint32_t(cnt) ? sh.edges[t] : INVALID_ID;
    if (threadIdx.x == 0) {
      a.deg[A] = uint32_t(cnt);
      a.dirty[A] = 1;
    }
  }
}

__global__ void drop_rows_kernel(const uint32_t* ids, uint32_t n, uint32_t R, uint32_t* adj, uint32_t* deg,
                                 uint8_t* exists) {
  const uint32_t b = blockIdx.x;
  if (b >= n) return;
  const uint32_t id = ids[b];
  for (uint32_t t = threadIdx.x; t < R; t += blockDim.x) adj[size_t(id) * R + t] = INVALID_ID;
  if (threadIdx.x == 0) {
    deg[id] = 0;
    exists[id] = 0;
  }
}

}  // namespace

int upload_start_extra(sdb_index* ix) {
  const size_t n = ix->h_start_extra.size();
  if (n == 0) return SDB_OK;
  int rc = ix->d_start_extra.ensure(n);
  if (rc) return rc;
  SDB_CUDA(cudaMemcpyAsync(ix->d_start_extra.p, ix->h_start_extra.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice,
                           ix->stream));
  SDB_CUDA(cudaStreamSynchronize(ix->stream));
  return SDB_OK;
}

namespace {

struct ScanBufs {  // device scratch of one EdgeScan, freed on scope exit
  DevBuf<uint32_t> d_del, d_inb, d_prune, d_save, d_misc;
  DevBuf<uint8_t> d_pmark, d_smark;
  DevBuf<unsigned char> d_tmp;
  uint32_t n_prune = 0, n_save = 0;
  ~ScanBufs() {
    d_del.release(); d_inb.release(); d_prune.release(); d_save.release(); d_misc.release(); d_pmark.release();
    d_smark.release(); d_tmp.release();
  }
};

// EdgeScan (node.go:142-199): fills sb.d_prune / sb.d_save (ascending ids) and their counts.
int edge_scan_device(sdb_index* ix, const std::vector<uint8_t>& h_del, ScanBufs& sb) {
  cudaStream_t st = ix->stream;
  const uint32_t rows = ix->rows, R = ix->p.degree_bound;
  const size_t words = (size_t(rows) + 31) / 32;
  std::vector<uint32_t> h_bits(words, 0);
  for (uint32_t id = 0; id < rows; ++id)
    if (h_del[id]) h_bits[id >> 5] |= 1u << (id & 31);
  int rc;
  if ((rc = sb.d_del.ensure(words)) || (rc = sb.d_inb.ensure(words)) || (rc = sb.d_pmark.ensure(rows)) ||
      (rc = sb.d_smark.ensure(rows)) || (rc = sb.d_prune.ensure(rows)) || (rc = sb.d_save.ensure(rows)) ||
      (rc = sb.d_misc.ensure(8)))
    return rc;
  SDB_CUDA(cudaMemcpyAsync(sb.d_del.p, h_bits.data(), words * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
  SDB_CUDA(cudaMemsetAsync(sb.d_inb.p, 0, words * sizeof(uint32_t), st));
  SDB_CUDA(cudaMemsetAsync(sb.d_pmark.p, 0, rows, st));
  SDB_CUDA(cudaMemsetAsync(sb.d_misc.p, 0, 8 * sizeof(uint32_t), st));
  const uint64_t n_slots = uint64_t(rows) * R;
  edge_scan_kernel<<<unsigned((n_slots + 255) / 256), 256, 0, st>>>(ix->d_adj, n_slots, R, INVALID_ID, ix->d_exists,
                                                                     sb.d_del.p, sb.d_inb.p, sb.d_pmark.p);
  ix->launches++;
  SDB_CUDA(cudaGetLastError());
  const uint32_t n_extra = uint32_t(ix->h_start_extra.size());
  if (n_extra) {
    edge_scan_kernel<<<(n_extra + 255) / 256, 256, 0, st>>>(ix->d_start_extra.p, n_extra, R, START_ID, ix->d_exists,
                                                             sb.d_del.p, sb.d_inb.p, sb.d_pmark.p);
    ix->launches++;
    SDB_CUDA(cudaGetLastError());
  }
  save_mark_kernel<<<(rows + 255) / 256, 256, 0, st>>>(ix->d_exists, sb.d_del.p, sb.d_inb.p, rows, sb.d_smark.p);
  ix->launches++;
  SDB_CUDA(cudaGetLastError());
  thrust::counting_iterator<uint32_t> iota(0);
  size_t tb1 = 0, tb2 = 0;
  cub::DeviceSelect::Flagged(nullptr, tb1, iota, sb.d_pmark.p, sb.d_prune.p, sb.d_misc.p, int(rows), st);
  cub::DeviceSelect::Flagged(nullptr, tb2, iota, sb.d_smark.p, sb.d_save.p, sb.d_misc.p + 1, int(rows), st);
  if ((rc = sb.d_tmp.ensure(std::max(tb1, tb2) + 16))) return rc;
  size_t tb = std::max(tb1, tb2);
  SDB_CUDA(cub::DeviceSelect::Flagged(sb.d_tmp.p, tb, iota, sb.d_pmark.p, sb.d_prune.p, sb.d_misc.p, int(rows), st));
  tb = std::max(tb1, tb2);
  SDB_CUDA(cub::DeviceSelect::Flagged(sb.d_tmp.p, tb, iota, sb.d_smark.p, sb.d_save.p, sb.d_misc.p + 1, int(rows), st));
  ix->launches += 2;
  uint32_t h_cnt[2] = {0, 0};
  SDB_CUDA(cudaMemcpyAsync(h_cnt, sb.d_misc.p, sizeof(h_cnt), cudaMemcpyDeviceToHost, st));
  SDB_CUDA(cudaStreamSynchronize(st));
  sb.n_prune = h_cnt[0];
  sb.n_save = h_cnt[1];
  return SDB_OK;
}

// removeInboundEdges (prune.go:88-154) for the rows flagged in h_del (size ix->rows).
int remove_inbound_edges(sdb_index* ix, const std::vector<uint8_t>& h_del) {
  cudaStream_t st = ix->stream;
  const uint32_t R = ix->p.degree_bound;
  ScanBufs sb;
  int rc = edge_scan_device(ix, h_del, sb);
  if (rc) return rc;
  const uint32_t n_extra = uint32_t(ix->h_start_extra.size());
  // pruneDeleteNeighbour for every toPrune node
  if (sb.n_prune) {
    PruneDelArgs pa{};
    pa.s = make_view(ix);
    pa.adj = ix->d_adj; pa.deg = ix->d_deg; pa.dirty = ix->d_dirty; pa.R = R; pa.alpha = ix->p.alpha;
    pa.delbits = sb.d_del.p; pa.prune_ids = sb.d_prune.p; pa.n_prune = sb.n_prune;
    pa.start_extra = ix->d_start_extra.p; pa.n_start_extra = n_extra;
    const size_t cap = (size_t(R) + n_extra) * (R + 1) + 1;
    const size_t smem = PruneShared::bytes(int(cap));
    if (smem + 1024 > ix->smem_optin)
      return fail(SDB_ERR_INTERNAL, "start node has too many overflow edges for pruneDeleteNeighbour");
    pa.cap = int(cap);
    pa.error_flag = sb.d_misc.p + 2;
    SDB_CUDA(cudaFuncSetAttribute(prune_delete_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    const uint32_t grid = std::min<uint32_t>(sb.n_prune, uint32_t(ix->sm_count) * 3);
    prune_delete_kernel<<<grid, PRUNE_THREADS, smem, st>>>(pa);
    ix->launches++;
    SDB_CUDA(cudaGetLastError());
    uint32_t h_err = 0;
    uint8_t start_mark = 0;
    SDB_CUDA(cudaMemcpyAsync(&h_err, sb.d_misc.p + 2, sizeof(h_err), cudaMemcpyDeviceToHost, st));
    SDB_CUDA(cudaMemcpyAsync(&start_mark, sb.d_pmark.p + START_ID, 1, cudaMemcpyDeviceToHost, st));
    SDB_CUDA(cudaStreamSynchronize(st));
    if (h_err == 1) return fail(SDB_ERR_INTERNAL, "pruneDeleteNeighbour candidate list overflow");
    if (h_err == 2) return fail(SDB_ERR_INTERNAL, "no neighbours to be deleted for a point flagged by EdgeScan");
    if (start_mark) ix->h_start_extra.clear();  // a pruned start node has at most R edges again
  }
  // orphans go back to the start node (prune.go:137-151)
  if (sb.n_save) {
    std::vector<uint32_t> save(sb.n_save), row(R);
    uint32_t dS = 0;
    SDB_CUDA(cudaMemcpyAsync(save.data(), sb.d_save.p, sb.n_save * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    SDB_CUDA(cudaMemcpyAsync(row.data(), ix->d_adj + size_t(START_ID) * R, R * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    SDB_CUDA(cudaMemcpyAsync(&dS, ix->d_deg + START_ID, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    SDB_CUDA(cudaStreamSynchronize(st));
    for (uint32_t p : save) {
      bool present = false;
      for (uint32_t t = 0; t < dS && !present; ++t) present = row[t] == p;
      for (uint32_t x : ix->h_start_extra) present = present || x == p;
      if (present) continue;
      if (dS < R) row[dS++] = p;
      else ix->h_start_extra.push_back(p);
    }
    SDB_CUDA(cudaMemcpyAsync(ix->d_adj + size_t(START_ID) * R, row.data(), R * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    SDB_CUDA(cudaMemcpyAsync(ix->d_deg + START_ID, &dS, sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    SDB_CUDA(cudaMemsetAsync(ix->d_dirty + START_ID, 1, 1, st));
    SDB_CUDA(cudaStreamSynchronize(st));
    if ((rc = upload_start_extra(ix))) return rc;
  }
  return SDB_OK;
}

}  // namespace

// Ascending ids of the nodes whose edge list changed since the last clearing call
// (graphNode.isDirty / CheckAndClearDirty, node.go:17,104-110): what a flush must rewrite.
int dirty_edges_locked(sdb_index* ix, uint64_t cap, uint64_t* ids_out, uint64_t* n_out, int clear) {
  cudaStream_t st = ix->stream;
  const uint32_t rows = ix->rows;
  DevBuf<uint32_t> d_ids, d_cnt;
  DevBuf<unsigned char> d_tmp;
  struct Rel {
    DevBuf<uint32_t>&a, &b; DevBuf<unsigned char>& c;
    ~Rel() { a.release(); b.release(); c.release(); }
  } rel{d_ids, d_cnt, d_tmp};
  int rc;
  if ((rc = d_ids.ensure(rows)) || (rc = d_cnt.ensure(1))) return rc;
  thrust::counting_iterator<uint32_t> iota(0);
  size_t tb = 0;
  cub::DeviceSelect::Flagged(nullptr, tb, iota, ix->d_dirty, d_ids.p, d_cnt.p, int(rows), st);
  if ((rc = d_tmp.ensure(tb + 16))) return rc;
  SDB_CUDA(cub::DeviceSelect::Flagged(d_tmp.p, tb, iota, ix->d_dirty, d_ids.p, d_cnt.p, int(rows), st));
  ix->launches++;
  uint32_t n = 0;
  SDB_CUDA(cudaMemcpyAsync(&n, d_cnt.p, sizeof(n), cudaMemcpyDeviceToHost, st));
  SDB_CUDA(cudaStreamSynchronize(st));
  *n_out = n;
  const uint32_t take = uint32_t(std::min<uint64_t>(n, cap));
  std::vector<uint32_t> h(take);
  if (take) SDB_CUDA(cudaMemcpy(h.data(), d_ids.p, take * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  for (uint32_t i = 0; i < take; ++i) ids_out[i] = h[i];
  if (clear && take == n) SDB_CUDA(cudaMemsetAsync(ix->d_dirty, 0, rows, st));
  SDB_CUDA(cudaStreamSynchronize(st));
  return SDB_OK;
}

int edge_scan_locked(sdb_index* ix, uint64_t n_delete, const uint64_t* delete_ids, uint64_t* to_prune,
                     uint64_t* n_prune, uint64_t* to_save, uint64_t* n_save) {
  std::vector<uint8_t> h_del(ix->rows, 0);
  for (uint64_t i = 0; i < n_delete; ++i)
    if (delete_ids[i] < ix->rows) h_del[delete_ids[i]] = 1;
  ScanBufs sb;
  int rc = edge_scan_device(ix, h_del, sb);
  if (rc) return rc;
  std::vector<uint32_t> tp(sb.n_prune), ts(sb.n_save);
  if (sb.n_prune) SDB_CUDA(cudaMemcpy(tp.data(), sb.d_prune.p, sb.n_prune * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  if (sb.n_save) SDB_CUDA(cudaMemcpy(ts.data(), sb.d_save.p, sb.n_save * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  for (uint32_t i = 0; i < sb.n_prune; ++i) to_prune[i] = tp[i];
  for (uint32_t i = 0; i < sb.n_save; ++i) to_save[i] = ts[i];
  *n_prune = sb.n_prune;
  *n_save = sb.n_save;
  return SDB_OK;
}

int insert_update_delete_locked(sdb_index* ix, uint64_t n, const uint64_t* ids, const float* vectors,
                                const uint8_t* has_vector) {
  const uint32_t dim = ix->p.dim;
  std::vector<uint64_t> ins_ids, upd_ids;
  std::vector<uint32_t> del_ids;
  std::vector<float> ins_vecs, upd_vecs;
  for (uint64_t i = 0; i < n; ++i) {
    if (ids[i] == START_ID) return fail(SDB_ERR_RESERVED_ID, "cannot modify point with start id: 1");  // vamana.go:150
    if (ids[i] == 0) return fail(SDB_ERR_RESERVED_ID, "invalid point id: 0");                           // vamana.go:154
    if (ids[i] >= (uint64_t(1) << 31) - 1) return fail(SDB_ERR_INVALID, "node id too large for a device index");
    const bool exists = ids[i] < ix->rows && ix->h_exists[ids[i]];
    const bool hv = has_vector ? has_vector[i] != 0 : true;
    if (!exists && !hv) continue;  // nothing to do (vamana.go:161-163)
    if (!exists) {
      ins_ids.push_back(ids[i]);
      ins_vecs.insert(ins_vecs.end(), vectors + i * dim, vectors + (i + 1) * dim);
    } else if (hv) {
      upd_ids.push_back(ids[i]);
      upd_vecs.insert(upd_vecs.end(), vectors + i * dim, vectors + (i + 1) * dim);
    } else {
      del_ids.push_back(uint32_t(ids[i]));
    }
  }
  {
    // the same id twice in one call is a race in the reference; refuse it here
    std::vector<uint64_t> all(ins_ids);
    all.insert(all.end(), upd_ids.begin(), upd_ids.end());
    for (uint32_t d : del_ids) all.push_back(d);
    std::sort(all.begin(), all.end());
    if (std::adjacent_find(all.begin(), all.end()) != all.end())
      return fail(SDB_ERR_INVALID, "the same point id appears more than once in one InsertUpdateDelete call");
  }
  int rc;
  if (!ins_ids.empty() && (rc = insert_batch_locked(ix, ins_ids.size(), ins_ids.data(), ins_vecs.data()))) return rc;
  if (!upd_ids.empty() || !del_ids.empty()) {
    if (ix->rows <= START_ID || !ix->h_exists[START_ID]) return fail(SDB_ERR_STATE, "failed to get start point");
    std::vector<uint8_t> h_del(ix->rows, 0);
    for (uint64_t id : upd_ids) h_del[id] = 1;
    for (uint32_t id : del_ids) h_del[id] = 1;
    if ((rc = remove_inbound_edges(ix, h_del))) return rc;
  }
  if (!del_ids.empty()) {
    // vecStore.Delete + nodeStore.Delete (vamana.go:231-236)
    if ((rc = ix->d_tmp32.ensure(del_ids.size()))) return rc;
    SDB_CUDA(cudaMemcpyAsync(ix->d_tmp32.p, del_ids.data(), del_ids.size() * sizeof(uint32_t), cudaMemcpyHostToDevice,
                             ix->stream));
    drop_rows_kernel<<<unsigned(del_ids.size()), 64, 0, ix->stream>>>(ix->d_tmp32.p, uint32_t(del_ids.size()),
                                                                      ix->p.degree_bound, ix->d_adj, ix->d_deg,
                                                                      ix->d_exists);
    ix->launches++;
    SDB_CUDA(cudaGetLastError());
    SDB_CUDA(cudaStreamSynchronize(ix->stream));
    for (uint32_t id : del_ids) {
      ix->h_exists[id] = 0;
      ix->count--;
    }
    ix->vec_epoch++;
  }
  if (!upd_ids.empty() && (rc = insert_batch_locked(ix, upd_ids.size(), upd_ids.data(), upd_vecs.data(), true))) return rc;
  return SDB_OK;
}

}  // namespace sdb
