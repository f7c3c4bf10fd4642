// Exchange operators of a distributed plan (SURVEY 8e "Q3 joins"): the reference is single-process, so nothing here mirrors
// a reference file -- these are the two nodes a distributed planner inserts around the reference's operators.
//
//   Broadcast(child)                 every rank executes `child` over its shard; the result rows of ALL ranks, in rank
//                                    order, become this node's output on every rank (the build side of a broadcast join).
//                                    One small all-gather carries the ranks' row counts, NULL counts, integer column
//                                    statistics and table identities; ONE grouped NCCL exchange then moves every column
//                                    slice straight into its place in the gathered columns (no padding, no compaction).
//   FinalAggregate(child, keys, ops) `child` produced PARTIAL groups per rank (groups may straddle the shards).  The rows
//                                    are hash-partitioned on the first key, exchanged all-to-all (group -> owner rank), and
//                                    re-aggregated locally with the merge operator of every value column (SUM for partial
//                                    sums and counts, MIN, MAX).  The result stays sharded: every final group lives on
//                                    exactly one rank.
// Both run on the context's stream through the library-owned communicator (comm.cu: NCCL, or the in-process rendezvous
// group of the single-GPU tests).  The child of a Broadcast is purely local, so it runs under its own speculation scope
// (learned device-side counts replayed without host round trips, fused.cu); the collectives themselves are issued exactly
// once per execution: the node memoises its output for the duration of one C-ABI call, so that a re-run of an enclosing
// speculative pipeline never repeats them on one rank only.
#include <cstring>

#include "comm.h"
#include "launch.h"
#include "plan.h"

namespace qgpu {

namespace {

struct BroadcastMemo {
  unsigned long long epoch = 0;
  View view;
  uint64_t signature = 0;
};

struct ColMeta {
  long long null_count, has_stats, vmin, vmax, phys;
};

void mark_order_free(PlanNode& n) {
  if (n.kind == PK_HASH_JOIN) {
    n.order_free = true;
    return;
  }
  if (n.kind == PK_PROJECTION || n.kind == PK_FILTER) mark_order_free(*n.children[0]);
}

int world_of(Ctx* ctx) { return ctx->comm ? ctx->comm->world : 1; }

// host words -> device words as KERNEL ARGUMENTS (no pageable-copy synchronisation on the per-step path)
struct Words32 {
  long long v[32];
};
__global__ void k_store_words(long long* __restrict__ dst, Words32 w, int n) {
  if ((int)threadIdx.x < n) dst[threadIdx.x] = w.v[threadIdx.x];
}
void store_words(Ctx* ctx, long long* dst, const long long* vals, size_t n) {
  for (size_t at = 0; at < n; at += 32) {
    Words32 w;
    const int m = (int)std::min<size_t>(32, n - at);
    for (int i = 0; i < m; ++i) w.v[i] = vals[at + i];
    LAUNCH(ctx, k_store_words, 1, 32, 0, dst + at, w, m);
  }
}

}  // namespace

uint64_t local_subtree_signature(PlanNode& n);  // fused.cu

struct PruneParams {
  int n_dest, n_cols, key_width, pad;
  long long lo[COMM_MAX_WORLD], hi[COMM_MAX_WORLD];
  const unsigned char* src[16];
  unsigned char* dst[16];
  int width[16];
};
// pass 1: per row the set of destination ranks (key inside their probe range) + per-destination row counts
__global__ void k_prune_count(const void* __restrict__ keys, int64_t n, PruneParams P, unsigned char* __restrict__ mask,
                              unsigned long long* __restrict__ counts) {
  __shared__ unsigned int sc[COMM_MAX_WORLD];
  if (threadIdx.x < COMM_MAX_WORLD) sc[threadIdx.x] = 0;
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const long long k = P.key_width == 8 ? ((const long long*)keys)[i] : (long long)((const int*)keys)[i];
    unsigned m = 0;
    for (int r = 0; r < P.n_dest; ++r)
      if (k >= P.lo[r] && k <= P.hi[r]) {
        m |= 1u << r;
        atomicAdd(&sc[r], 1u);
      }
    mask[i] = (unsigned char)m;
  }
  __syncthreads();
  if (threadIdx.x < P.n_dest && sc[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)sc[threadIdx.x]);
}
// pass 2: every selected row of every column into its destination's slice of the send buffers (order inside a slice is
// arbitrary: the consumer is order-free)
__global__ void k_prune_scatter(int64_t n, PruneParams P, const unsigned char* __restrict__ mask, unsigned long long* __restrict__ cursor) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t n_words = (n + 31) >> 5;
  for (int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < n_words; w += warps) {
    const int64_t i = (w << 5) + lane;
    const unsigned m = i < n ? mask[i] : 0u;
    for (int r = 0; r < P.n_dest; ++r) {  // warp-aggregated reservation: one atomic per (warp, destination)
      const unsigned b = __ballot_sync(0xffffffffu, (m >> r) & 1u);
      if (!b) continue;
      unsigned long long base = 0;
      if (lane == __ffs(b) - 1) base = atomicAdd(&cursor[r], (unsigned long long)__popc(b));
      base = __shfl_sync(0xffffffffu, base, __ffs(b) - 1);
      if ((m >> r) & 1u) {
        const unsigned long long pos = base + __popc(b & ((1u << lane) - 1u));
        for (int c = 0; c < P.n_cols; ++c) {
          const int wd = P.width[c];
          if (wd == 8) ((unsigned long long*)P.dst[c])[pos] = ((const unsigned long long*)P.src[c])[i];
          else if (wd == 4) ((unsigned int*)P.dst[c])[pos] = ((const unsigned int*)P.src[c])[i];
          else if (wd == 16) ((ulonglong2*)P.dst[c])[pos] = ((const ulonglong2*)P.src[c])[i];
          else if (wd == 2) ((unsigned short*)P.dst[c])[pos] = ((const unsigned short*)P.src[c])[i];
          else P.dst[c][pos] = P.src[c][i];
        }
      }
    }
  }
}

// Key-range pruned broadcast (dynamic partition pruning by zone maps): a build row travels only to the ranks whose
// probe-side key range contains its key -- with range-sharded inputs (orders and lineitem are both sorted by order key) every
// rank receives ~1 / world of the rows instead of all of them, and builds a correspondingly smaller join table.
static View broadcast_pruned(PlanNode& node, const View& in, const std::vector<DColP>& mine, const std::vector<long long>& am,
                             size_t meta_words, size_t HDR, int world, int rank) {
  Ctx* ctx = node.ctx;
  const int nc = (int)mine.size();
  const int64_t n = in.num_rows;
  PruneParams P;
  memset(&P, 0, sizeof(P));
  P.n_dest = world;
  P.n_cols = nc;
  P.key_width = phys_width(mine[node.prune_key_col]->phys);
  for (int r = 0; r < world; ++r) {
    P.lo[r] = am[(size_t)r * meta_words + 3];
    P.hi[r] = am[(size_t)r * meta_words + 4];
  }
  DBufP mask = ctx->alloc(std::max<size_t>((size_t)n, 16));
  DBufP counts = ctx->alloc_zero(8 * (size_t)world), matrix = ctx->alloc(8 * (size_t)world * world);
  if (n > 0) LAUNCH(ctx, k_prune_count, grid_for(ctx, n, 256), 256, 0, mine[node.prune_key_col]->data->ptr, n, P, (unsigned char*)mask->ptr, (unsigned long long*)counts->ptr);
  comm_all_gather_any(ctx, counts->ptr, matrix->ptr, 8 * (size_t)world);
  std::vector<long long> mx((size_t)world * world);
  ctx->d2h_sync(mx.data(), matrix->ptr, mx.size() * 8);  // second (and last) host round trip: who sends how much to whom
  std::vector<int64_t> send0((size_t)world + 1, 0), recv0((size_t)world + 1, 0);
  for (int r = 0; r < world; ++r) {
    send0[r + 1] = send0[r] + mx[(size_t)rank * world + r];
    recv0[r + 1] = recv0[r] + mx[(size_t)r * world + rank];
  }
  const int64_t n_send = send0[world], n_recv = recv0[world];
  std::vector<DBufP> sendbuf((size_t)nc);
  std::vector<DColP> out((size_t)nc);
  std::vector<Xfer> sends, recvs;
  for (int c = 0; c < nc; ++c) {
    const int w = phys_width(mine[c]->phys);
    sendbuf[c] = ctx->alloc(std::max<size_t>((size_t)n_send * w, 16));
    P.src[c] = (const unsigned char*)mine[c]->data->ptr;
    P.dst[c] = (unsigned char*)sendbuf[c]->ptr;
    P.width[c] = w;
    auto d = std::make_shared<DCol>();
    d->type = mine[c]->type;
    d->phys = mine[c]->phys;
    d->length = n_recv;
    d->data = ctx->alloc(std::max<size_t>((size_t)n_recv * w, 16));
    // conservative statistics: the bounds over everything the ranks hold contain what arrived here
    bool stats = true;
    long long mn = INT64_MAX, mxv = INT64_MIN;
    for (int r = 0; r < world; ++r) {
      const long long* m = &am[(size_t)r * meta_words + HDR + 5 * c];
      if (am[(size_t)r * meta_words] == 0) continue;
      if (!m[1]) stats = false;
      mn = std::min(mn, m[2]);
      mxv = std::max(mxv, m[3]);
    }
    if (c == node.prune_key_col) {  // the key column is additionally bounded by this rank's probe range
      mn = std::max(mn, (long long)P.lo[rank]);
      mxv = std::min(mxv, (long long)P.hi[rank]);
    }
    if (stats && mn <= mxv) {
      d->has_stats = true;
      d->vmin = mn;
      d->vmax = mxv;
    }
    for (int r = 0; r < world; ++r) {
      sends.push_back({r, (char*)sendbuf[c]->ptr + (size_t)send0[r] * w, (size_t)(send0[r + 1] - send0[r]) * w});
      recvs.push_back({r, (char*)d->data->ptr + (size_t)recv0[r] * w, (size_t)(recv0[r + 1] - recv0[r]) * w});
    }
    out[c] = d;
  }
  if (n > 0 && n_send > 0) {
    DBufP cursor = ctx->alloc(8 * (size_t)world);
    std::vector<long long> c0(send0.begin(), send0.end() - 1);
    store_words(ctx, (long long*)cursor->ptr, c0.data(), (size_t)world);
    LAUNCH(ctx, k_prune_scatter, grid_for(ctx, n, 256), 256, 0, n, P, (const unsigned char*)mask->ptr, (unsigned long long*)cursor->ptr);
  }
  comm_exchange(ctx, sends, recvs);
  ctx->trace("broadcast: pruned exchange");
  View v;
  v.schema = in.schema;
  v.num_rows = n_recv;
  v.num_batches = 1;
  for (int c = 0; c < nc; ++c) v.cols.push_back({out[c], nullptr});
  int64_t total = 0;
  for (int r = 0; r < world; ++r) total += am[(size_t)r * meta_words];
  node.strategy = "broadcast[key-range pruned: " + std::to_string(n_recv) + " of " + std::to_string(total) + " rows of " + std::to_string(world) +
                  " ranks needed here, " + std::to_string(n_send) + " sent]";
  return v;
}

// executes (once per C-ABI call) and returns the memo of a Broadcast node
static BroadcastMemo& broadcast_run(PlanNode& node) {
  Ctx* ctx = node.ctx;
  auto memo = std::static_pointer_cast<BroadcastMemo>(node.exchange_cache);
  if (!memo) {
    memo = std::make_shared<BroadcastMemo>();
    node.exchange_cache = memo;
  }
  if (memo->epoch == ctx->exec_epoch && memo->epoch != 0) return *memo;
  PlanNode& child = *node.children[0];
  if (node.order_free) mark_order_free(child);
  const uint64_t child_sig = local_subtree_signature(child);
  ctx->trace(nullptr);
  View in = run_speculated(node, child_sig, [&] { return child.execute(); });
  in.resolve();
  ctx->trace("broadcast: child plan");
  const int world = world_of(ctx);
  const int rank = ctx->comm ? ctx->comm->rank : 0;
  if (world == 1) {
    memo->view = in;
    memo->signature = child_sig;
    memo->epoch = ctx->exec_epoch;
    return *memo;
  }
  const int nc = (int)in.cols.size();
  const int64_t n_local = in.num_rows;
  std::vector<DColP> mine((size_t)nc);
  for (int c = 0; c < nc; ++c) {
    if (!in.cols[c].base) throw_internal("Broadcast: column '" + in.schema.fields[c].name + "' was not uploaded to the GPU table");
    mine[c] = materialize(ctx, in.cols[c], n_local);
    if (mine[c]->phys == PH_STR || mine[c]->phys == PH_BIT)
      throw_internal("Broadcast exchanges fixed-width columns only (column '" + in.schema.fields[c].name + "' is " + in.schema.fields[c].type.str() + ")");
  }
  // ---- meta block: [n_rows, child signature, probe key range (has, lo, hi); per column: NULLs, has_stats, min, max, phys] --------------------------
  constexpr size_t HDR = 5;  // n_rows, child signature, probe range: has / lo / hi
  const size_t meta_words = HDR + 5 * (size_t)nc;
  DBufP meta = ctx->alloc_zero(meta_words * 8), all_meta = ctx->alloc(meta_words * 8 * (size_t)world);
  std::vector<long long> h(meta_words, 0);
  h[0] = n_local;
  h[1] = (long long)child_sig;
  // pruned broadcast: the value range of this rank's PROBE-side key column (table statistics, cached on the column)
  if (node.prune_table) {
    TableImpl& pt = *node.prune_table;
    pt.consolidate();
    if (node.prune_probe_col >= 0 && node.prune_probe_col < (int)pt.cols.size() && pt.cols[node.prune_probe_col]) {
      DCol& pc = *pt.cols[node.prune_probe_col];
      {
        SpecSuspend cached_work(ctx);
        ensure_stats(ctx, pc);
      }
      if (pc.has_stats && pc.null_count == 0) {
        h[2] = 1;
        h[3] = (long long)pc.vmin;
        h[4] = (long long)pc.vmax;
      } else if (pt.num_rows == 0) {
        h[2] = 1;  // an empty probe side needs no build rows at all
        h[3] = 1;
        h[4] = 0;
      }
    }
  }
  for (int c = 0; c < nc; ++c) {
    h[HDR + 5 * c] = mine[c]->null_count;
    h[HDR + 5 * c + 4] = (long long)mine[c]->phys;
  }
  std::vector<char> want_stats((size_t)nc, 0);
  for (int c = 0; c < nc; ++c) {
    const Phys ph = mine[c]->phys;
    want_stats[c] = ph == PH_I8 || ph == PH_I16 || ph == PH_I32 || ph == PH_I64 || ph == PH_D64 || ph == PH_U8 || ph == PH_U16 || ph == PH_U32;
    h[HDR + 5 * c + 1] = want_stats[c] ? 1 : 0;
  }
  store_words(ctx, (long long*)meta->ptr, h.data(), meta_words);
  for (int c = 0; c < nc; ++c)
    if (want_stats[c]) stats_to_device(ctx, *mine[c], (long long*)meta->ptr + HDR + 5 * c + 2);
  comm_all_gather_any(ctx, meta->ptr, all_meta->ptr, meta_words * 8);
  std::vector<long long> am(meta_words * (size_t)world);
  ctx->d2h_sync(am.data(), all_meta->ptr, am.size() * 8);  // the one host round trip of the exchange
  ctx->trace("broadcast: materialise + meta all-gather");
  std::vector<int64_t> rows((size_t)world), row0((size_t)world + 1, 0);
  uint64_t sig = 0x42524f4144ULL;
  for (int r = 0; r < world; ++r) {
    rows[r] = am[(size_t)r * meta_words];
    row0[r + 1] = row0[r] + rows[r];
    sig = (sig ^ (uint64_t)am[(size_t)r * meta_words + 1]) * 0xff51afd7ed558ccdULL;
    sig = (sig ^ (uint64_t)rows[r]) * 0xff51afd7ed558ccdULL;
    sig ^= sig >> 29;
  }
  // ---- pruned variant: every rank receives only the rows whose key lies in ITS probe-side key range ---------------------
  if (node.prune_table && node.order_free) {
    bool ok = node.prune_key_col >= 0 && node.prune_key_col < nc && nc <= 16;
    for (int r = 0; r < world && ok; ++r) ok = am[(size_t)r * meta_words + 2] != 0;
    for (int c = 0; c < nc && ok; ++c)
      for (int r = 0; r < world && ok; ++r) {
        const long long* m = &am[(size_t)r * meta_words + HDR + 5 * c];
        ok = m[0] == 0 && (Phys)m[4] == mine[c]->phys && mine[c]->phys != PH_NULL;
      }
    const int kw = ok ? phys_width(mine[node.prune_key_col]->phys) : 0;
    ok = ok && (kw == 4 || kw == 8) && mine[node.prune_key_col]->phys != PH_F32 && mine[node.prune_key_col]->phys != PH_F64 &&
         mine[node.prune_key_col]->phys != PH_U64;
    if (ok) {
      View v = broadcast_pruned(node, in, mine, am, meta_words, HDR, world, rank);
      for (int r = 0; r < world; ++r) sig = (sig ^ (uint64_t)am[(size_t)r * meta_words + 3] ^ ((uint64_t)am[(size_t)r * meta_words + 4] << 1)) * 0xff51afd7ed558ccdULL;
      memo->view = v;
      memo->signature = sig ^ (uint64_t)v.num_rows;
      memo->epoch = ctx->exec_epoch;
      return *memo;
    }
  }
  const int64_t total = row0[world];
  // ---- gathered columns + the grouped exchange -----------------------------------------------------------------------
  std::vector<Xfer> sends, recvs;
  std::vector<DColP> out((size_t)nc);
  std::vector<DBufP> vtmp;  // validity bitmaps travel as whole words per rank and are re-packed at the destination rows
  struct Repack {
    int col;
    DBufP words;
    std::vector<int64_t> word0;
  };
  std::vector<Repack> repacks;
  for (int c = 0; c < nc; ++c) {
    auto d = std::make_shared<DCol>();
    d->type = mine[c]->type;
    d->length = total;
    // every rank materialises the same physical layout for a column unless one holds only NULL placeholders
    Phys phys = mine[c]->phys;
    long long nulls = 0;
    bool stats = true;
    long long mn = INT64_MAX, mx = INT64_MIN;
    for (int r = 0; r < world; ++r) {
      const long long* m = &am[(size_t)r * meta_words + HDR + 5 * c];
      nulls += m[0];
      if (rows[r] > 0 && (Phys)m[4] != PH_NULL) {
        if (phys == PH_NULL) phys = (Phys)m[4];
        else if ((Phys)m[4] != phys)
          throw_internal("Broadcast: ranks disagree on the physical layout of column '" + in.schema.fields[c].name + "'");
      }
      if (rows[r] > m[0]) {  // the rank has non-NULL values
        if (!m[1]) stats = false;
        mn = std::min(mn, m[2]);
        mx = std::max(mx, m[3]);
      }
    }
    d->null_count = nulls;
    if (phys == PH_NULL || nulls == total) {
      d->phys = PH_NULL;
      d->null_count = total;
      out[c] = d;
      continue;
    }
    d->phys = phys;
    const int w = phys_width(phys);
    d->data = ctx->alloc(std::max<size_t>((size_t)total * w, 16));
    if (stats && mn <= mx) {
      d->has_stats = true;
      d->vmin = mn;
      d->vmax = mx;
    }
    const bool have_data = mine[c]->phys == phys && mine[c]->data;  // else: NULL placeholders only, peers get zeros
    DBufP zeros;
    if (!have_data && n_local > 0) {
      zeros = ctx->alloc_zero((size_t)n_local * w);
      vtmp.push_back(zeros);
    }
    for (int r = 0; r < world; ++r) {
      sends.push_back({r, have_data ? mine[c]->data->ptr : (zeros ? zeros->ptr : nullptr), (size_t)n_local * w});
      recvs.push_back({r, (char*)d->data->ptr + (size_t)row0[r] * w, (size_t)rows[r] * w});
    }
    if (nulls > 0) {
      d->validity = ctx->alloc_zero(std::max<size_t>((size_t)((total + 31) >> 5) * 4, 4));
      Repack rp;
      rp.col = c;
      rp.word0.assign((size_t)world + 1, 0);
      for (int r = 0; r < world; ++r) rp.word0[r + 1] = rp.word0[r] + ((rows[r] + 31) >> 5);
      rp.words = ctx->alloc(std::max<size_t>((size_t)rp.word0[world] * 4, 4));
      DBufP my_bits = mine[c]->validity;
      if (mine[c]->phys == PH_NULL) {  // all NULL on this rank
        my_bits = ctx->alloc_zero(std::max<size_t>((size_t)((n_local + 31) >> 5) * 4, 4));
        vtmp.push_back(my_bits);
      } else if (!my_bits) {  // no NULLs on this rank: all ones
        my_bits = ctx->alloc(std::max<size_t>((size_t)((n_local + 31) >> 5) * 4, 4));
        fill_bits(ctx, (uint32_t*)my_bits->ptr, 0, ((n_local + 31) >> 5) * 32, true);
        vtmp.push_back(my_bits);
      }
      for (int r = 0; r < world; ++r) {
        sends.push_back({r, my_bits->ptr, (size_t)((n_local + 31) >> 5) * 4});
        recvs.push_back({r, (char*)rp.words->ptr + (size_t)rp.word0[r] * 4, (size_t)((rows[r] + 31) >> 5) * 4});
      }
      repacks.push_back(rp);
    }
    out[c] = d;
  }
  comm_exchange(ctx, sends, recvs);
  for (Repack& rp : repacks)
    for (int r = 0; r < world; ++r)
      copy_bits(ctx, (uint32_t*)out[rp.col]->validity->ptr, row0[r], (const uint32_t*)rp.words->ptr + rp.word0[r], 0, rows[r]);
  ctx->trace("broadcast: grouped exchange");
  // the send buffers (`mine`, this function's locals) are freed in stream order behind the exchange
  View v;
  v.schema = in.schema;
  v.num_rows = total;
  v.num_batches = total > 0 || in.num_batches > 0 ? 1 : 0;
  for (int c = 0; c < nc; ++c) v.cols.push_back({out[c], nullptr});
  (void)rank;
  memo->view = v;
  memo->signature = sig;
  memo->epoch = ctx->exec_epoch;
  node.strategy = "broadcast[all-gather of " + std::to_string(total) + " rows over " + std::to_string(world) + " ranks]";
  return *memo;
}

View run_broadcast(PlanNode& node) { return broadcast_run(node).view; }
uint64_t broadcast_signature(PlanNode& node) { return broadcast_run(node).signature; }

// ------------------------------------------------------------------------------------------------
// FinalAggregate
// ------------------------------------------------------------------------------------------------
static std::shared_ptr<ExprNode> column_expr(int index) {
  auto e = std::make_shared<ExprNode>();
  e->kind = QGPU_IR_COLUMN;
  e->col_index = index;
  return e;
}

struct KeyRanges {
  int n;
  long long lo[COMM_MAX_WORLD], hi[COMM_MAX_WORLD];
};
// bit i = key[i] lies in one of the ranges
__global__ void k_range_bits(const void* __restrict__ keys, int width, int64_t n, KeyRanges R, uint32_t* __restrict__ bits) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t n_words = (n + 31) >> 5;
  for (int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < n_words; w += warps) {
    const int64_t i = (w << 5) + lane;
    bool in = false;
    if (i < n) {
      const long long k = width == 8 ? ((const long long*)keys)[i] : (long long)((const int*)keys)[i];
      for (int r = 0; r < R.n; ++r) in = in || (k >= R.lo[r] && k <= R.hi[r]);
    }
    const uint32_t m = __ballot_sync(0xffffffffu, in);
    if (lane == 0) bits[w] = m;
  }
}

// [a | b] of two fixed-width NULL-free columns of the same layout
static DColP concat_fixed(Ctx* ctx, const DColP& a, const DColP& b) {
  if (b->length == 0) return a;
  if (a->length == 0) return b;
  if (a->phys != b->phys || a->null_count || b->null_count) throw_internal("FinalAggregate: cannot concatenate these columns");
  const int w = phys_width(a->phys);
  auto d = std::make_shared<DCol>();
  d->type = a->type;
  d->phys = a->phys;
  d->length = a->length + b->length;
  d->data = ctx->alloc((size_t)d->length * w);
  CUDA_CHECK(cudaMemcpyAsync(d->data->ptr, a->data->ptr, (size_t)a->length * w, cudaMemcpyDeviceToDevice, ctx->stream));
  CUDA_CHECK(cudaMemcpyAsync((char*)d->data->ptr + (size_t)a->length * w, b->data->ptr, (size_t)b->length * w, cudaMemcpyDeviceToDevice, ctx->stream));
  return d;
}

// hash exchange + local re-aggregation of `in` (every rank calls this together); result columns in `in`'s column order
static View exchange_and_merge(PlanNode& node, const View& in) {
  Ctx* ctx = node.ctx;
  const int world = world_of(ctx), rank = ctx->comm->rank;
  const int nc = (int)in.cols.size();
  // ---- partial groups -> a table, rows grouped by owner rank -------------------------------------------------------------
  TableImpl part;
  part.ctx = ctx;
  part.schema = in.schema;
  part.num_rows = in.num_rows;
  part.num_batches = 1;
  part.cols.resize((size_t)nc);
  for (int c = 0; c < nc; ++c) part.cols[c] = materialize(ctx, in.cols[c], in.num_rows);
  std::vector<int64_t> offs;
  std::shared_ptr<TableImpl> grouped = hash_partition_table(part, node.exchange_keys[0], world, offs);
  ctx->trace("final-aggregate: hash partition");
  // ---- counts: my row of the (source x destination) matrix to everybody ---------------------------------------------------
  std::vector<long long> my_counts((size_t)world);
  for (int r = 0; r < world; ++r) my_counts[r] = offs[r + 1] - offs[r];
  DBufP cnt = ctx->alloc((size_t)world * 8), all_cnt = ctx->alloc((size_t)world * world * 8);
  store_words(ctx, (long long*)cnt->ptr, my_counts.data(), (size_t)world);
  comm_all_gather_any(ctx, cnt->ptr, all_cnt->ptr, (size_t)world * 8);
  std::vector<long long> matrix((size_t)world * world);
  ctx->d2h_sync(matrix.data(), all_cnt->ptr, matrix.size() * 8);
  std::vector<int64_t> recv0((size_t)world + 1, 0);
  for (int r = 0; r < world; ++r) recv0[r + 1] = recv0[r] + matrix[(size_t)r * world + rank];
  const int64_t n_recv = recv0[world];
  // ---- all-to-all of every column in one grouped exchange -----------------------------------------------------------------
  std::vector<Xfer> sends, recvs;
  View rv;
  rv.schema = in.schema;
  rv.num_rows = n_recv;
  rv.num_batches = 1;
  for (int c = 0; c < nc; ++c) {
    const DCol& s = *grouped->cols[c];
    const int w = phys_width(s.phys);
    auto d = std::make_shared<DCol>();
    d->type = s.type;
    d->phys = s.phys;
    d->length = n_recv;
    d->data = ctx->alloc(std::max<size_t>((size_t)n_recv * w, 16));
    for (int r = 0; r < world; ++r) {
      sends.push_back({r, (char*)s.data->ptr + (size_t)offs[r] * w, (size_t)(offs[r + 1] - offs[r]) * w});
      recvs.push_back({r, (char*)d->data->ptr + (size_t)recv0[r] * w, (size_t)(recv0[r + 1] - recv0[r]) * w});
    }
    rv.cols.push_back({d, nullptr});
  }
  comm_exchange(ctx, sends, recvs);
  ctx->trace("final-aggregate: counts + all-to-all");
  // ---- local re-aggregation with the merge operators ------------------------------------------------------------------------
  Schema agg_schema;
  std::vector<std::shared_ptr<ExprNode>> exprs;  // keep the nodes alive while compiled
  std::vector<std::shared_ptr<Compiled>> keys;
  std::vector<int> order;  // aggregate output position -> child column
  for (int k : node.exchange_keys) {
    exprs.push_back(column_expr(k));
    keys.push_back(compile_expr(*exprs.back(), rv.schema));
    agg_schema.fields.push_back(in.schema.fields[k]);
    order.push_back(k);
  }
  std::vector<AggSpec> specs;
  for (size_t i = 0; i < node.exchange_cols.size(); ++i) {
    const int c = node.exchange_cols[i];
    exprs.push_back(column_expr(c));
    AggSpec s;
    s.op = node.exchange_ops[i] == 1 ? QGPU_AGG_MIN : (node.exchange_ops[i] == 2 ? QGPU_AGG_MAX : QGPU_AGG_SUM);
    s.arg = compile_expr(*exprs.back(), rv.schema);
    s.return_type = in.schema.fields[c].type;
    s.expr_type = in.schema.fields[c].type;
    specs.push_back(s);
    agg_schema.fields.push_back(in.schema.fields[c]);
    order.push_back(c);
  }
  View merged = run_aggregate(ctx, rv, keys, specs, agg_schema, nullptr);
  merged.resolve();
  ctx->trace("final-aggregate: re-aggregate");
  View out;
  out.schema = in.schema;
  out.num_rows = merged.num_rows;
  out.num_batches = merged.num_batches;
  out.cols.resize((size_t)nc);
  for (size_t p = 0; p < order.size(); ++p) out.cols[(size_t)order[p]] = merged.cols[p];
  return out;
}

View run_final_aggregate(PlanNode& node) {
  Ctx* ctx = node.ctx;
  View in = node.child_view(0);
  const int world = world_of(ctx);
  if (world == 1) {
    node.strategy = "final-aggregate[single rank: identity]";
    return in;
  }
  const int rank = ctx->comm->rank;
  const int nc = (int)in.cols.size();
  if (node.exchange_keys.empty()) throw_internal("FinalAggregate needs at least one key column");
  std::vector<char> seen((size_t)nc, 0);
  for (int k : node.exchange_keys) {
    if (k < 0 || k >= nc) throw_internal("FinalAggregate: key column index out of range");
    seen[k] = 1;
  }
  for (int c : node.exchange_cols) {
    if (c < 0 || c >= nc) throw_internal("FinalAggregate: value column index out of range");
    seen[c] = 1;
  }
  for (int c = 0; c < nc; ++c) {
    if (!seen[c]) throw_internal("FinalAggregate: every child column must be a key or a merged value");
    if (!in.cols[c].base) throw_internal("FinalAggregate: column '" + in.schema.fields[c].name + "' is not resident");
  }
  ctx->trace("final-aggregate: child plan");
  // ---- which groups can exist on another rank at all?  Key-range pruning: the ranks publish the value range of the first
  // key over their partial groups; only rows whose key falls into ANOTHER rank's range are contested and take part in the
  // exchange -- with range-sharded inputs (lineitem is sorted by l_orderkey) that is a handful of boundary groups, with
  // unrelated sharding it is everything.  The decision is taken from the same gathered ranges on every rank.
  const int64_t n = in.num_rows;
  DColP k0 = materialize(ctx, in.cols[node.exchange_keys[0]], n);
  const int kw = phys_width(k0->phys);
  const bool int_key = (kw == 4 || kw == 8) && k0->phys != PH_F32 && k0->phys != PH_F64 && k0->phys != PH_U64 && k0->null_count == 0;
  DBufP meta = ctx->alloc_zero(32), all_meta = ctx->alloc(32 * (size_t)world);
  long long hdr[2] = {n, int_key ? 1 : 0};
  store_words(ctx, (long long*)meta->ptr, hdr, 2);
  if (int_key) stats_to_device(ctx, *k0, (long long*)meta->ptr + 2);
  comm_all_gather_any(ctx, meta->ptr, all_meta->ptr, 32);
  std::vector<long long> am(4 * (size_t)world);
  ctx->d2h_sync(am.data(), all_meta->ptr, am.size() * 8);
  bool prunable = true;
  for (int r = 0; r < world; ++r) prunable = prunable && am[4 * r + 1] != 0;
  bool any_overlap = false;
  KeyRanges mine_contested;
  mine_contested.n = 0;
  if (prunable) {
    for (int a = 0; a < world; ++a)
      for (int b = a + 1; b < world; ++b) {
        if (am[4 * a] == 0 || am[4 * b] == 0) continue;
        const long long lo = std::max(am[4 * a + 2], am[4 * b + 2]), hi = std::min(am[4 * a + 3], am[4 * b + 3]);
        if (lo > hi) continue;
        any_overlap = true;
        if (a == rank || b == rank) {
          mine_contested.lo[mine_contested.n] = lo;
          mine_contested.hi[mine_contested.n] = hi;
          mine_contested.n++;
        }
      }
  }
  ctx->trace("final-aggregate: key ranges");
  if (prunable && !any_overlap) {
    node.strategy = "final-aggregate[key ranges of the " + std::to_string(world) + " ranks are disjoint: " + std::to_string(n) + " local groups are final]";
    return in;
  }
  if (!prunable) {
    View out = exchange_and_merge(node, in);
    out.schema = node.schema;
    node.strategy = "final-aggregate[hash exchange of " + std::to_string(n) + " partial groups over " + std::to_string(world) + " ranks -> " +
                    std::to_string(out.num_rows) + " groups here]";
    return out;
  }
  // split: contested rows go through the exchange, the others are final where they are
  IdxP contested, safe;
  if (n > 0 && mine_contested.n > 0) {
    DBufP bits = ctx->alloc(std::max<size_t>((size_t)((n + 31) >> 5) * 4, 4));
    LAUNCH(ctx, k_range_bits, grid_for(ctx, n, 256), 256, 0, k0->data->ptr, kw, n, mine_contested, (uint32_t*)bits->ptr);
    contested = rows_by_bit(ctx, bits, n, 0);
    safe = rows_by_bit(ctx, bits, n, 1);
  } else {
    contested = std::make_shared<IdxVec>();
    contested->buf = ctx->alloc(8);
    safe = nullptr;  // everything
  }
  View cv = apply_selection_view(ctx, in, contested);
  const int64_t n_contested = contested->length;
  View merged = exchange_and_merge(node, cv);
  View sv = safe ? apply_selection_view(ctx, in, safe) : in;
  View out;
  out.schema = node.schema;
  out.num_rows = sv.num_rows + merged.num_rows;
  out.num_batches = 1;
  for (int c = 0; c < nc; ++c) {
    DColP a = materialize(ctx, sv.cols[c], sv.num_rows), b = materialize(ctx, merged.cols[c], merged.num_rows);
    out.cols.push_back({concat_fixed(ctx, a, b), nullptr});
  }
  node.strategy = "final-aggregate[key-range pruning: " + std::to_string(n - n_contested) + " local groups final, hash exchange of " +
                  std::to_string(n_contested) + " contested partial groups over " + std::to_string(world) + " ranks -> " +
                  std::to_string(merged.num_rows) + " merged here]";
  return out;
}

}  // namespace qgpu
