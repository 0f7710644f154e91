// Multi-GPU plumbing: one process per GPU, NCCL over NVLink 5 / NVSwitch.
//
// The scan shards by contiguous record range: no data-path collective; the
// only exchange is ONE all-gatherv of the compacted, ordered hit lists at the end (rank-order
// concatenation == global listing order), and an 8-byte min-reduce for a range-sharded chain.
// NCCL has no native gatherv: counts are all-gathered first, then every (rank, query) segment
// is a grouped ncclBroadcast straight into its final position of the gathered list.
//
// libnccl is dlopen()ed at first use so the library also loads on machines without NCCL and
// shares the copy a host process (e.g. torch) may already have mapped.
#include "corpus.h"
#include <dlfcn.h>
#include <vector>
#include <string.h>

namespace fei {
namespace {

typedef struct { char internal[128]; } ncclUniqueId;
typedef void* ncclComm_t;
enum { ncclSuccess = 0 };
enum { ncclInt64 = 4, ncclUint64 = 5 };
enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 };

struct Nccl {
  void* h = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  ncclComm_t comm = nullptr;
  int nranks = 0, rank = -1;
  DevBuf counts_dev, gathered, scratch, lists;
  CompactScratch compact;
  std::mutex mu;                 // one collective at a time per communicator
};
Nccl g;

int load_nccl() {
  if (g.h) return FEI_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) { g.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (g.h) break; }
  if (!g.h) { set_error("cannot dlopen libnccl.so.2: %s", dlerror()); return FEI_E_NCCL; }
#define SYM(field, name) *(void**)(&g.field) = dlsym(g.h, name); if (!g.field) { set_error("libnccl lacks %s", name); return FEI_E_NCCL; }
  SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
  SYM(AllGather, "ncclAllGather") SYM(Broadcast, "ncclBroadcast") SYM(AllReduce, "ncclAllReduce")
  SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  return FEI_OK;
}

int nccl_fail(int rc, const char* what) {
  set_error("NCCL error %d (%s) in %s", rc, g.GetErrorString ? g.GetErrorString(rc) : "?", what);
  return FEI_E_NCCL;
}
#define FEI_NCCL(call) do { int rc__ = (call); if (rc__ != ncclSuccess) return nccl_fail(rc__, #call); } while (0)

}  // namespace
}  // namespace fei

using namespace fei;

extern "C" int fei_comm_unique_id(uint8_t* id) {
  if (!id) { set_error("null id"); return FEI_E_BADARG; }
  FEI_TRY(load_nccl());
  ncclUniqueId u;
  FEI_NCCL(g.GetUniqueId(&u));
  memcpy(id, u.internal, FEI_NCCL_ID_BYTES);
  return FEI_OK;
}

extern "C" int fei_comm_init(const uint8_t* id, int nranks, int rank) {
  FEI_TRY(require_ready());
  if (!id || nranks < 1 || rank < 0 || rank >= nranks) { set_error("bad communicator arguments"); return FEI_E_BADARG; }
  FEI_TRY(load_nccl());
  if (g.comm) { g.CommDestroy(g.comm); g.comm = nullptr; }
  ncclUniqueId u; memcpy(u.internal, id, FEI_NCCL_ID_BYTES);
  FEI_NCCL(g.CommInitRank(&g.comm, nranks, u, rank));
  g.nranks = nranks; g.rank = rank;
  return FEI_OK;
}

extern "C" int fei_comm_destroy(void) {
  if (g.comm) { g.CommDestroy(g.comm); g.comm = nullptr; }
  g.counts_dev.release(); g.gathered.release(); g.scratch.release(); g.lists.release();
  g.compact.blk_counts.release(); g.compact.blk_offsets.release(); g.compact.totals.release();
  g.nranks = 0; g.rank = -1;
  return FEI_OK;
}

// All-gatherv of the scan result of every rank.  Two wire formats, chosen per call from the gathered counts:
//   sparse : the compacted per-query index lists (8 B per hit), one grouped ncclBroadcast per (rank, query)
//            segment straight into its final position;
//   dense  : when the lists would be larger than the per-record hit masks (4 B per record) — typical for
//            many-pattern batches where most records hit — the masks are all-gathered instead (one
//            ncclAllGather) and the global ordered lists are compacted from them on demand.
// Either way rank-order concatenation is the global listing order.
extern "C" int fei_comm_allgather_hits(fei_corpus* c, uint32_t nq, uint64_t* const* hits, const uint64_t* cap,
                                       uint64_t* nhits_total, uint64_t* counts_out) {
  FEI_TRY(require_ready());
  if (!c || nq == 0 || nq > 32 || nq != c->last_nq) { set_error("no matching scan result on this corpus (run fei_scan_count / fei_scan_hits first)"); return FEI_E_STATE; }
  if (!g.comm) { set_error("fei_comm_init() has not been called"); return FEI_E_STATE; }
  std::lock_guard<std::mutex> lock(g.mu);
  std::lock_guard<std::mutex> clock(c->mu);
  cudaStream_t s = ctx().stream;
  const int R = g.nranks;
  const uint32_t W = nq + 2;                                   // per-rank record: counts[nq], n, global_base
  FEI_TRY(g.counts_dev.ensure((size_t)(R + 1) * 34 * sizeof(uint64_t)));
  uint64_t mine_h[34];
  for (uint32_t q = 0; q < nq; ++q) mine_h[q] = c->last_counts[q];
  mine_h[nq] = c->n; mine_h[nq + 1] = c->global_base;
  uint64_t* mine = g.counts_dev.as<uint64_t>() + (size_t)R * 34;
  FEI_CUDA(cudaMemcpyAsync(mine, mine_h, W * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
  FEI_NCCL(g.AllGather(mine, g.counts_dev.p, W, ncclUint64, g.comm, s));
  std::vector<uint64_t> info((size_t)R * W);
  FEI_CUDA(cudaMemcpyAsync(info.data(), g.counts_dev.p, info.size() * 8, cudaMemcpyDeviceToHost, s));
  FEI_CUDA(cudaStreamSynchronize(s));
  std::vector<uint64_t> tot(nq, 0);
  uint64_t list_entries = 0, n_total = 0, n_max = 0;
  for (int r = 0; r < R; ++r) {
    for (uint32_t q = 0; q < nq; ++q) { uint64_t v = info[(size_t)r * W + q]; tot[q] += v; list_entries += v; if (counts_out) counts_out[(size_t)r * nq + q] = v; }
    uint64_t nr = info[(size_t)r * W + nq];
    n_total += nr; if (nr > n_max) n_max = nr;
  }
  for (uint32_t q = 0; q < nq; ++q) if (nhits_total) nhits_total[q] = tot[q];
  const bool want_host = hits && cap;
  const bool dense = list_entries * 8 > (uint64_t)R * n_max * 4;
  bool truncated = false;
  if (!dense) {
    std::vector<uint64_t> qbase(nq + 1, 0);
    for (uint32_t q = 0; q < nq; ++q) qbase[q + 1] = qbase[q] + tot[q];
    FEI_TRY(g.gathered.ensure((qbase[nq] + 1) * sizeof(uint64_t)));
    FEI_NCCL(g.GroupStart());
    for (uint32_t q = 0; q < nq; ++q) {
      uint64_t pos = qbase[q];
      for (int r = 0; r < R; ++r) {
        uint64_t cnt = info[(size_t)r * W + q];
        if (cnt) {
          const void* src = c->hit_lists.as<uint64_t>() + (size_t)q * c->hit_list_stride;   // only read on the root
          FEI_NCCL(g.Broadcast(src, g.gathered.as<uint64_t>() + pos, cnt, ncclUint64, r, g.comm, s));
        }
        pos += cnt;
      }
    }
    FEI_NCCL(g.GroupEnd());
    for (uint32_t q = 0; want_host && q < nq; ++q) {
      if (!hits[q]) continue;
      uint64_t take = tot[q] < cap[q] ? tot[q] : cap[q];
      if (take < tot[q]) truncated = true;
      if (take) FEI_CUDA(cudaMemcpyAsync(hits[q], g.gathered.as<uint64_t>() + qbase[q], take * 8, cudaMemcpyDeviceToHost, s));
    }
  } else {
    // dense: gather the masks (every rank contributes n_max entries; the tail of short shards is ignored)
    if (c->hits.bytes < n_max * sizeof(uint32_t)) {            // shorter shard than the longest one: grow, keep contents
      DevBuf nb;
      FEI_TRY(nb.alloc(n_max * sizeof(uint32_t)));
      FEI_CUDA(cudaMemsetAsync(nb.p, 0, n_max * sizeof(uint32_t), s));
      if (c->n) FEI_CUDA(cudaMemcpyAsync(nb.p, c->hits.p, c->n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
      FEI_CUDA(cudaStreamSynchronize(s));
      void* tp = nb.p; nb.p = c->hits.p; c->hits.p = tp;
      size_t tb = nb.bytes; nb.bytes = c->hits.bytes; c->hits.bytes = tb;
    }
    FEI_TRY(g.gathered.ensure((size_t)R * n_max * sizeof(uint32_t)));
    FEI_NCCL(g.AllGather(c->hits.p, g.gathered.p, n_max, 3 /* ncclUint32 */, g.comm, s));
    if (want_host) {
      CompactScratch& sc = g.compact; DevBuf& lists = g.lists;
      std::vector<uint64_t> written(nq, 0);
      for (int r = 0; r < R; ++r) {
        uint64_t nr = info[(size_t)r * W + nq], gb = info[(size_t)r * W + nq + 1];
        uint64_t cnt[32], stride = 1;
        FEI_TRY(compact_masks(g.gathered.as<uint32_t>() + (size_t)r * n_max, nr, nq, gb, sc, cnt, &lists, &stride, nullptr, s));
        for (uint32_t q = 0; q < nq; ++q) {
          if (!hits[q] || !cnt[q]) continue;
          uint64_t room = cap[q] > written[q] ? cap[q] - written[q] : 0;
          uint64_t take = cnt[q] < room ? cnt[q] : room;
          if (take < cnt[q]) truncated = true;
          if (take) FEI_CUDA(cudaMemcpyAsync(hits[q] + written[q], lists.as<uint64_t>() + (size_t)q * stride, take * 8, cudaMemcpyDeviceToHost, s));
          written[q] += take;
        }
        FEI_CUDA(cudaStreamSynchronize(s));
      }
    }
  }
  FEI_CUDA(cudaStreamSynchronize(s));
  if (truncated) { set_error("gathered hit buffer too small for at least one query (see nhits_total)"); return FEI_E_CAPACITY; }
  return FEI_OK;
}

extern "C" int fei_comm_allreduce_first_bad(int64_t* first_bad, int32_t* bad_kind) {
  FEI_TRY(require_ready());
  if (!first_bad || !bad_kind) { set_error("null argument"); return FEI_E_BADARG; }
  if (!g.comm) { set_error("fei_comm_init() has not been called"); return FEI_E_STATE; }
  cudaStream_t s = ctx().stream;
  // key = index*4 + kind, "no failure" = max uint64; min over ranks keeps the reference's first failure
  uint64_t key = *first_bad < 0 ? ~0ull : ((uint64_t)*first_bad << 2 | (uint64_t)(*bad_kind & 3));
  FEI_TRY(g.scratch.ensure(16));
  FEI_CUDA(cudaMemcpyAsync(g.scratch.p, &key, 8, cudaMemcpyHostToDevice, s));
  FEI_NCCL(g.AllReduce(g.scratch.p, g.scratch.p, 1, ncclUint64, ncclMin, g.comm, s));
  FEI_CUDA(cudaMemcpyAsync(&key, g.scratch.p, 8, cudaMemcpyDeviceToHost, s));
  FEI_CUDA(cudaStreamSynchronize(s));
  if (key == ~0ull) { *first_bad = -1; *bad_kind = 0; }
  else { *first_bad = (int64_t)(key >> 2); *bad_kind = (int32_t)(key & 3); }
  return FEI_OK;
}
