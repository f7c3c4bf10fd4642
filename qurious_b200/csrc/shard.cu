// Sharded (multi-GPU) aggregates: exact merge of per-shard partial aggregate states.
//
// The reference is single-process (no exchange operator exists in qurious); this is the one distribution
// strategy the B200 build adds for low/medium-cardinality aggregates over row-range shards:
//   1. every process runs the SAME plan over its shard up to -- but not including -- the finalisation of
//      HashAggregate / NoGroupingAggregate (aggregate/hash.rs:89-107, no_grouping.rs:48-61), i.e. it stops at
//      the per-group accumulator state (SUM words, counts, MIN/MAX, first row);
//   2. the state is packed into a fixed-size, self-describing block of 64-bit words in HBM
//      (k_pack_records): header + one record per group = packed KEY VALUES (so records from different shards
//      can be matched by value; no dictionary or statistics need to agree across shards), first row +
//      shard row offset, and (lo, hi, count) per aggregate;
//   3. the caller all-gathers the blocks (ncclAllGather -- qurious_b200/distributed.py uses torch.distributed);
//   4. every process merges the gathered records (k_merge_records, one CTA, deterministic order: 128-bit
//      integer adds, MIN/MAX, count sums) and runs the ordinary finalisation (finish_aggregate) on the merged
//      state.  Integer/decimal results are bit-identical to a single-GPU run over the whole table; Float64
//      sums differ only in reduction order (tolerance 1e-12, north_star).
#include <algorithm>
#include <cstring>

#include "comm.h"
#include "epilogue.h"
#include "launch.h"
#include "plan.h"

namespace qgpu {

#define SH_MAGIC 0x5147505553544154ULL  // "QGPUSTAT"
#define SH_HDR 8                        // header words
#define SH_MAX_MERGE 4096               // records one merge CTA handles
#define SH_MAX_KEYS 8
#define SH_MAX_AGGS 24

// header: [0] magic [1] n_groups [2] n_keys [3] n_aggs [4] overflow (too many groups / key too long) [5] max_groups
// record: n_keys x {tag, w1, w2} | first_row | n_aggs x {lo, hi, cnt}
//   tag: 0 = NULL key, 1 = fixed-width value (w1 = lo, w2 = hi, sign-extended), 2 + len = Utf8 of len <= 16 bytes
static inline int rec_words(int n_keys, int n_aggs) { return 3 * n_keys + 1 + 3 * n_aggs; }

struct PackArgs {
  int n_keys, n_aggs, max_groups, pad;
  long long row_offset;
  long long n_host;
  const long long* n_dev;
  const long long* first_row;
  ColRef key[SH_MAX_KEYS];
  const unsigned long long* lo[SH_MAX_AGGS];
  const unsigned long long* hi[SH_MAX_AGGS];
  const unsigned long long* cnt[SH_MAX_AGGS];
};

__global__ void __launch_bounds__(256) k_pack_records(const __grid_constant__ PackArgs a, unsigned long long* __restrict__ out) {
  const long long n = a.n_dev ? *a.n_dev : a.n_host;
  const int rw = 3 * a.n_keys + 1 + 3 * a.n_aggs;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    out[0] = SH_MAGIC;
    out[1] = (unsigned long long)min(n, (long long)a.max_groups);
    out[2] = (unsigned long long)a.n_keys;
    out[3] = (unsigned long long)a.n_aggs;
    if (n > a.max_groups) out[4] = 1;
    out[5] = (unsigned long long)a.max_groups;
  }
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < n && g < a.max_groups; g += stride) {
    unsigned long long* r = out + SH_HDR + g * rw;
    const long long row = a.first_row[g];
    for (int k = 0; k < a.n_keys; ++k) {
      const Val v = load_col(a.key[k], row);
      unsigned long long tag = 0, w1 = 0, w2 = 0;
      if (v.valid) {
        if (a.key[k].phys == PH_STR) {
          const int len = (int)v.hi;
          if (len > 16) {
            out[4] = 1;
          } else {
            const unsigned char* s = (const unsigned char*)v.lo;
            for (int i = 0; i < len; ++i) {
              if (i < 8) w1 |= (unsigned long long)s[i] << (8 * i);
              else w2 |= (unsigned long long)s[i] << (8 * (i - 8));
            }
            tag = 2 + (unsigned long long)len;
          }
        } else {
          tag = 1;
          w1 = v.lo;
          w2 = v.hi;
          if (a.key[k].phys != PH_I128 && a.key[k].phys != PH_D64) w2 = 0;
        }
      }
      r[3 * k] = tag;
      r[3 * k + 1] = w1;
      r[3 * k + 2] = w2;
    }
    r[3 * a.n_keys] = (unsigned long long)(row + a.row_offset);
    for (int i = 0; i < a.n_aggs; ++i) {
      unsigned long long* q = r + 3 * a.n_keys + 1 + 3 * i;
      q[0] = a.lo[i][g];
      q[1] = a.hi[i][g];
      q[2] = a.cnt[i][g];
    }
  }
}

struct KeyOut {
  void* data;
  int32_t* offsets;   // Utf8
  uint32_t* validity;
  int phys;
  int pad;
};
struct MergeArgs {
  int n_states, n_keys, n_aggs, max_groups;
  long long block_words;
  int kind[SH_MAX_AGGS];  // AccKind
  unsigned long long* lo[SH_MAX_AGGS];
  unsigned long long* hi[SH_MAX_AGGS];
  unsigned long long* cnt[SH_MAX_AGGS];
  KeyOut key[SH_MAX_KEYS];
  long long* first_row;
  long long* n_groups_out;
  unsigned long long* key_nulls;  // [n_keys]
  int* err;                       // 1: bad block, 2: overflow flagged by a shard, 3: too many records
};

__device__ __forceinline__ bool i128_less(unsigned long long alo, unsigned long long ahi, unsigned long long blo, unsigned long long bhi) {
  return ((long long)ahi < (long long)bhi) || (ahi == bhi && alo < blo);
}

// One CTA merges every record of every shard.  Deterministic: a group's records are folded in (shard, slot) order.
__global__ void __launch_bounds__(1024) k_merge_records(const __grid_constant__ MergeArgs a, const unsigned long long* __restrict__ in) {
  __shared__ unsigned int pos[SH_MAX_MERGE];     // word offset of record i inside `in`
  __shared__ unsigned short leader[SH_MAX_MERGE];
  __shared__ unsigned short gid[SH_MAX_MERGE];
  __shared__ int base[64];
  __shared__ int total_s, bad_s, ng_s;
  const int tid = threadIdx.x;
  const int rw = 3 * a.n_keys + 1 + 3 * a.n_aggs;
  if (tid == 0) {
    int t = 0, bad = 0;
    for (int s = 0; s < a.n_states; ++s) {
      const unsigned long long* h = in + (long long)s * a.block_words;
      base[s] = t;
      if (h[0] != SH_MAGIC || (int)h[2] != a.n_keys || (int)h[3] != a.n_aggs) bad = 1;
      else if (h[4]) bad = 2;
      else t += (int)h[1];
    }
    if (!bad && t > SH_MAX_MERGE) bad = 3;
    total_s = bad ? 0 : t;
    bad_s = bad;
    if (bad) *a.err = bad;
  }
  __syncthreads();
  const int M = total_s;
  for (int s = 0; s < a.n_states; ++s) {
    const unsigned long long* h = in + (long long)s * a.block_words;
    const int n = bad_s ? 0 : (int)h[1];
    for (int g = tid; g < n; g += blockDim.x) pos[base[s] + g] = (unsigned int)((long long)s * a.block_words + SH_HDR + (long long)g * rw);
  }
  __syncthreads();
  // leader of record i = first record with the same key words
  for (int i = tid; i < M; i += blockDim.x) {
    const unsigned long long* ri = in + pos[i];
    int l = i;
    for (int j = 0; j < i; ++j) {
      const unsigned long long* rj = in + pos[j];
      bool eq = true;
      for (int w = 0; w < 3 * a.n_keys && eq; ++w) eq = ri[w] == rj[w];
      if (eq) {
        l = j;
        break;
      }
    }
    leader[i] = (unsigned short)l;
  }
  __syncthreads();
  if (tid == 0) {
    int ng = 0;
    for (int i = 0; i < M; ++i)
      if (leader[i] == i) gid[i] = (unsigned short)ng++;
    ng_s = ng;
    *a.n_groups_out = ng;
  }
  __syncthreads();
  const int NG = ng_s;
  // one thread per merged group folds that group's records in order
  for (int i = tid; i < M; i += blockDim.x) {
    if (leader[i] != i) continue;
    const int t = gid[i];
    long long first = INT64_MAX;
    for (int ag = 0; ag < a.n_aggs; ++ag) {
      const int kind = a.kind[ag];
      unsigned long long lo = 0, hi = 0, cnt = 0;
      bool have = false;
      for (int j = i; j < M; ++j) {
        if (leader[j] != i) continue;
        const unsigned long long* q = in + pos[j] + 3 * a.n_keys + 1 + 3 * ag;
        const unsigned long long l2 = q[0], h2 = q[1];
        cnt += q[2];
        if (!have) {
          lo = l2;
          hi = h2;
          have = true;
          continue;
        }
        switch (kind) {
          case AK_COUNT: break;
          case AK_SUM_I64: lo += l2; break;
          case AK_SUM_DEC: {
            const u128 s = (((u128)hi << 64) | lo) + (((u128)h2 << 64) | l2);
            lo = (unsigned long long)s;
            hi = (unsigned long long)(s >> 64);
            break;
          }
          case AK_SUM_F64: lo = (unsigned long long)__double_as_longlong(__longlong_as_double((long long)lo) + __longlong_as_double((long long)l2)); break;
          case AK_MIN_I64: case AK_MIN_F64: if ((long long)l2 < (long long)lo) lo = l2; break;   // f64 travels as its total-order key
          case AK_MAX_I64: case AK_MAX_F64: if ((long long)l2 > (long long)lo) lo = l2; break;
          case AK_MIN_U64: if (l2 < lo) lo = l2; break;
          case AK_MAX_U64: if (l2 > lo) lo = l2; break;
          case AK_MIN_DEC: if (i128_less(l2, h2, lo, hi)) { lo = l2; hi = h2; } break;
          case AK_MAX_DEC: if (i128_less(lo, hi, l2, h2)) { lo = l2; hi = h2; } break;
          default: break;
        }
      }
      a.lo[ag][t] = lo;
      a.hi[ag][t] = hi;
      a.cnt[ag][t] = cnt;
    }
    for (int j = i; j < M; ++j)
      if (leader[j] == i) first = min(first, (long long)in[pos[j] + 3 * a.n_keys]);
    a.first_row[t] = first;
    // fixed-width key values
    const unsigned long long* r = in + pos[i];
    for (int k = 0; k < a.n_keys; ++k) {
      const unsigned long long tag = r[3 * k], w1 = r[3 * k + 1], w2 = r[3 * k + 2];
      const KeyOut& ko = a.key[k];
      switch (ko.phys) {
        case PH_I8: case PH_U8: ((uint8_t*)ko.data)[t] = (uint8_t)w1; break;
        case PH_I16: case PH_U16: ((uint16_t*)ko.data)[t] = (uint16_t)w1; break;
        case PH_I32: case PH_U32: ((uint32_t*)ko.data)[t] = (uint32_t)w1; break;
        case PH_I64: case PH_U64: ((unsigned long long*)ko.data)[t] = w1; break;
        case PH_I128: ((ulonglong2*)ko.data)[t] = make_ulonglong2(w1, w2); break;
        default: break;  // PH_STR below
      }
      if (tag == 0) atomicAdd(&a.key_nulls[k], 1ull);
    }
  }
  __syncthreads();
  // validity words + Utf8 offsets/bytes (group count is small: serial per key column)
  for (int k = tid; k < a.n_keys; k += blockDim.x) {
    const KeyOut& ko = a.key[k];
    for (int w = 0; w < (NG + 31) / 32; ++w) ko.validity[w] = 0;
    int off = 0;
    for (int i = 0; i < M; ++i) {
      if (leader[i] != i) continue;
      const int t = gid[i];
      const unsigned long long* r = in + pos[i];
      const unsigned long long tag = r[3 * k];
      if (tag) ko.validity[t >> 5] |= 1u << (t & 31);
      if (ko.phys == PH_STR) {
        ko.offsets[t] = off;
        const int len = tag >= 2 ? (int)(tag - 2) : 0;
        for (int b = 0; b < len; ++b) {
          const unsigned long long w = b < 8 ? r[3 * k + 1] : r[3 * k + 2];
          ((char*)ko.data)[off + b] = (char)((w >> (8 * (b & 7))) & 0xff);
        }
        off += len;
      }
    }
    if (ko.phys == PH_STR) ko.offsets[NG] = off;
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
PlanNode* find_aggregate_node(PlanNode& root) {
  PlanNode* n = &root;
  while (n->kind == PK_PROJECTION || n->kind == PK_FILTER) n = n->children[0].get();
  if (n->kind != PK_AGGREGATE) throw_internal("sharded execution needs a plan of the form (Projection|Filter)* <- Aggregate <- ...");
  return n;
}

int64_t shard_state_bytes(PlanNode& root, int32_t max_groups) {
  PlanNode* agg = find_aggregate_node(root);
  if (max_groups < 1 || max_groups > SH_MAX_MERGE) throw_internal("max_groups must be in [1, 4096]");
  return 8 * (int64_t)(SH_HDR + (int64_t)max_groups * rec_words((int)agg->group_exprs.size(), (int)agg->aggs.size()));
}

static Phys key_phys(const DType& t) {
  switch (t.id) {
    case QGPU_T_INT8: return PH_I8;
    case QGPU_T_INT16: return PH_I16;
    case QGPU_T_INT32: case QGPU_T_DATE32: case QGPU_T_TIME32: return PH_I32;
    case QGPU_T_INT64: case QGPU_T_DATE64: case QGPU_T_TIME64: return PH_I64;
    case QGPU_T_UINT8: return PH_U8;
    case QGPU_T_UINT16: return PH_U16;
    case QGPU_T_UINT32: return PH_U32;
    case QGPU_T_UINT64: return PH_U64;
    case QGPU_T_DECIMAL128: return PH_I128;
    case QGPU_T_UTF8: return PH_STR;
    default: throw_internal("unsupported group key type for sharded execution: " + t.str());
  }
}

void shard_partial_state(PlanNode& root, int64_t row_offset, int32_t max_groups, void* out_buf, int64_t cap_bytes) {
  PlanNode* agg = find_aggregate_node(root);
  Ctx* ctx = agg->ctx;
  const int nk = (int)agg->group_exprs.size(), na = (int)agg->aggs.size();
  if (nk > SH_MAX_KEYS || na > SH_MAX_AGGS) throw_internal("too many keys / aggregates for sharded execution");
  const int64_t need = shard_state_bytes(root, max_groups);
  if (cap_bytes < need) throw_internal("state buffer too small: need " + std::to_string(need) + " bytes");
  auto pend = std::make_shared<AggPending>();
  agg->merged_override.reset();
  agg->defer = pend.get();
  try {
    agg->execute();
  } catch (...) {
    agg->defer = nullptr;
    throw;
  }
  agg->defer = nullptr;
  agg->shard_pending = pend;
  CUDA_CHECK(cudaMemsetAsync(out_buf, 0, (size_t)need, ctx->stream));
  PackArgs a;
  memset(&a, 0, sizeof(a));
  a.n_keys = nk;
  a.n_aggs = na;
  a.max_groups = max_groups;
  a.row_offset = row_offset;
  if (!pend->set) {  // grouped aggregate over zero input batches (hash.rs:146-148): no groups
    a.n_host = 0;
    DBufP z = ctx->alloc_zero(8);
    a.first_row = (const long long*)z->ptr;
    LAUNCH(ctx, k_pack_records, 1, 256, 0, a, (unsigned long long*)out_buf);
    return;
  }
  GroupAccs& accs = pend->accs;
  a.n_host = accs.n_groups;
  a.n_dev = accs.n_groups_dev ? (const long long*)accs.n_groups_dev->ptr : nullptr;
  a.first_row = (const long long*)accs.first_row->ptr;
  for (int k = 0; k < nk; ++k) {
    Compiled& kc = *pend->keys[k];
    if (!kc.is_column_ref) throw_internal("sharded execution supports plain column group keys only");
    const LazyCol& lc = pend->input.cols[kc.column_ref];
    if (!lc.base) throw_internal("group key column was not uploaded");
    key_phys(kc.result_type);
    ColRef& r = a.key[k];
    const DCol& d = *lc.base;
    r.phys = d.phys;
    r.data = d.data ? d.data->ptr : nullptr;
    r.offsets = d.offsets ? (const int32_t*)d.offsets->ptr : nullptr;
    r.validity = d.validity ? (const uint32_t*)d.validity->ptr : nullptr;
    r.idx = lc.idx ? lc.idx->ptr() : nullptr;
  }
  for (int i = 0; i < na; ++i) {
    a.lo[i] = (const unsigned long long*)accs.lo[i]->ptr;
    a.hi[i] = (const unsigned long long*)accs.hi[i]->ptr;
    a.cnt[i] = (const unsigned long long*)accs.cnt[i]->ptr;
  }
  const int64_t n_bound = std::min<int64_t>(std::max<int64_t>(accs.n_groups, 1), max_groups);
  LAUNCH(ctx, k_pack_records, grid_for(ctx, n_bound, 256), 256, 0, a, (unsigned long long*)out_buf);
}

View shard_execute_merged(PlanNode& root, const void* gathered, int32_t n_states, int32_t max_groups) {
  PlanNode* agg = find_aggregate_node(root);
  Ctx* ctx = agg->ctx;
  std::shared_ptr<AggPending> pend = agg->shard_pending;
  const int nk = (int)agg->group_exprs.size(), na = (int)agg->aggs.size();
  if (n_states < 1 || n_states > 64) throw_internal("n_states must be in [1, 64]");
  // plan-side descriptions (keys / specs) come from the local partial run; a shard whose input had zero batches
  // compiles them against its own (empty) scan schema
  std::vector<std::shared_ptr<Compiled>> keys;
  std::vector<AggSpec> specs;
  std::vector<int> kinds;
  View local_input;
  if (pend && pend->set) {
    keys = pend->keys;
    specs = pend->specs;
    kinds = pend->accs.kind;
    local_input = pend->input;
  } else {
    throw_internal("qgpu_plan_execute_merged: call qgpu_plan_partial_state on this plan first (and the local shard must have at "
                   "least one batch)");
  }
  const int64_t m_max = std::min<int64_t>((int64_t)n_states * max_groups, SH_MAX_MERGE);
  const bool grouped = nk > 0;
  MergeArgs a;
  memset(&a, 0, sizeof(a));
  a.n_states = n_states;
  a.n_keys = nk;
  a.n_aggs = na;
  a.max_groups = max_groups;
  a.block_words = SH_HDR + (long long)max_groups * rec_words(nk, na);
  GroupAccs accs;
  accs.n_groups = grouped ? m_max : 1;
  accs.kind = kinds;
  // every zero-initialised buffer of the merge (3 per aggregate, first rows, flags, 3 per key) out of ONE allocation and
  // ONE memset: ~35 cudaMallocAsync + cudaMemsetAsync pairs per merged step otherwise sit between the all-gather and
  // the merge kernel
  Slab slab(ctx, (size_t)(3 * na + 1) * Slab::need((size_t)m_max * 8) + Slab::need(8 * (3 + SH_MAX_KEYS)) +
                     (size_t)nk * (Slab::need((size_t)((m_max + 31) / 32) * 4 + 4) + Slab::need((size_t)m_max * 16 + 16) +
                                   Slab::need((size_t)(m_max + 1) * 4)),
            true);
  for (int i = 0; i < na; ++i) {
    DBufP lo = slab.take((size_t)m_max * 8), hi = slab.take((size_t)m_max * 8), cnt = slab.take((size_t)m_max * 8);
    accs.lo.push_back(lo);
    accs.hi.push_back(hi);
    accs.cnt.push_back(cnt);
    a.kind[i] = kinds[i];
    a.lo[i] = (unsigned long long*)lo->ptr;
    a.hi[i] = (unsigned long long*)hi->ptr;
    a.cnt[i] = (unsigned long long*)cnt->ptr;
  }
  accs.first_row = slab.take((size_t)m_max * 8);
  a.first_row = (long long*)accs.first_row->ptr;
  // [0] n_groups, [1 .. 1+nk) key NULL counts, [1+nk] error code: the last nk+1 words travel to the host together
  // with finish_aggregate's own flags (one synchronisation for the whole merge)
  DBufP misc = slab.take(8 * (3 + SH_MAX_KEYS));
  a.n_groups_out = (long long*)misc->ptr;
  a.key_nulls = (unsigned long long*)((char*)misc->ptr + 8);
  a.err = (int*)((char*)misc->ptr + 8 + 8 * nk);
  if (grouped) accs.n_groups_dev = misc;  // first word is the merged group count
  std::vector<DColP> key_cols;
  for (int k = 0; k < nk; ++k) {
    auto c = std::make_shared<DCol>();
    c->type = keys[k]->result_type;
    c->phys = key_phys(c->type);
    c->length = m_max;
    c->validity = slab.take((size_t)((m_max + 31) / 32) * 4 + 4);
    if (c->phys == PH_STR) {
      c->data = slab.take((size_t)m_max * 16 + 16);
      c->offsets = slab.take((size_t)(m_max + 1) * 4);
      c->str_bytes = m_max * 16;
      c->str_bytes_is_bound = true;
      c->max_str_len = 16;  // state records carry Utf8 keys of <= 16 bytes: no k_str_maxlen pass + round trip per merge
    } else {
      c->data = slab.take((size_t)m_max * std::max(phys_width(c->phys), 1) + 16);
    }
    c->null_count = 1;  // replaced by the real count inside finish_aggregate
    a.key[k].data = c->data->ptr;
    a.key[k].offsets = c->offsets ? (int32_t*)c->offsets->ptr : nullptr;
    a.key[k].validity = (uint32_t*)c->validity->ptr;
    a.key[k].phys = c->phys;
    key_cols.push_back(c);
  }
  LAUNCH(ctx, k_merge_records, 1, 1024, 0, a, (const unsigned long long*)gathered);
  if (!grouped) {
    // n_groups_out is 1 when any shard contributed its single record; nothing else to do
  }
  View merged = finish_aggregate(ctx, local_input, keys, specs, agg->schema, accs, grouped ? &key_cols : nullptr, a.key_nulls);
  const int err = (int)accs.side_word;
  if (err == 2)
    throw_internal("sharded aggregate: a shard produced more than max_groups groups (or a key longer than 16 bytes); use hash "
                   "repartition for high-cardinality keys");
  if (err) throw_internal("sharded aggregate: malformed state block (" + std::to_string(err) + ")");
  for (auto& c : key_cols) c->length = merged.num_rows;
  agg->merged_override = std::make_shared<View>(merged);
  agg->strategy = "sharded-merge(" + std::to_string(n_states) + " states) <- " + agg->strategy;
  return root.execute();
}

// Whole sharded step below the C ABI: shard-local aggregate -> state block -> peer exchange + exact merge + finalisation
// in ONE kernel (epilogue.cu) -> the operators above the aggregate.  DENSE-eligible plans (Q1 / Q6) run as scan kernel +
// epilogue; everything else packs its generic accumulators (k_pack_records) into the same block format first, so ranks
// whose shards chose different local strategies still merge.  The result's metadata is pending (View::pending).
View shard_execute_fused(PlanNode& root, int64_t row_offset, int32_t max_groups) {
  PlanNode* agg = find_aggregate_node(root);
  Ctx* ctx = agg->ctx;
  if (!ctx->comm) throw QError(QGPU_ERR_NCCL, "NcclError: no communicator: call qgpu_comm_init first");
  if (max_groups < 1 || max_groups > EPI_MAXG) throw_internal("max_groups must be in [1, 4096]");
  agg->merged_override.reset();
  View merged;
  if (!try_fused_scan_aggregate_sharded(*agg, row_offset, max_groups, &merged)) {
    const int nk = (int)agg->group_exprs.size(), na = (int)agg->aggs.size();
    if (nk > EPI_MAXK || na > EPI_MAXAGG) throw_internal("too many keys / aggregates for sharded execution");
    const int64_t need = shard_state_bytes(root, max_groups);
    DBufP rec = ctx->alloc((size_t)need);
    shard_partial_state(root, row_offset, max_groups, rec->ptr, need);
    std::shared_ptr<AggPending> pend = agg->shard_pending;
    if (!pend || !pend->set)
      throw_internal("sharded execution: the local shard must have at least one batch (CREATE TABLE without INSERT cannot take part)");
    std::vector<DType> key_types;
    for (auto& k : pend->keys) key_types.push_back(k->result_type);
    EpiParams E;
    memset(&E, 0, sizeof(E));
    E.src = EPI_SRC_PACKED;
    epilogue_describe(ctx, key_types, pend->specs, pend->accs.kind, agg->schema, E);
    merged = epilogue_execute(ctx, E, key_types, pend->specs, agg->schema, true, row_offset, max_groups, rec);
    agg->strategy = "sharded[state block -> peer exchange + merge over " + std::to_string(ctx->comm->world) + " ranks] <- " + agg->strategy;
  }
  agg->merged_override = std::make_shared<View>(merged);
  return root.execute();
}

}  // namespace qgpu

// ================================================================================================
// Hash partitioning of a resident table ("high-cardinality group-by and join inputs are
// hash-repartitioned with NCCL all-to-all").  partition id = mix64(key) % n_parts; the output table holds the
// same rows grouped by partition (order inside a partition unspecified) -- the contiguous per-partition slices
// are what the caller hands to ncclSend/ncclRecv (torch.distributed.all_to_all_single).
// ================================================================================================
namespace qgpu {

#define PART_MAX 1024
#define PART_THREADS 256
#define PART_ITEMS 8
#define PART_TILE (PART_THREADS * PART_ITEMS)

__device__ __forceinline__ uint64_t part_mix(uint64_t x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return x;
}
__device__ __forceinline__ int64_t part_key(const void* key, int width, int64_t row) {
  return width == 8 ? ((const long long*)key)[row] : (int64_t)((const int*)key)[row];
}

__global__ void __launch_bounds__(PART_THREADS) k_part_hist(const void* __restrict__ key, int width, int64_t n, int n_parts,
                                                            unsigned long long* __restrict__ counts) {
  __shared__ unsigned int h[PART_MAX];
  for (int i = threadIdx.x; i < n_parts; i += blockDim.x) h[i] = 0;
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < n; row += stride)
    atomicAdd(&h[(unsigned)(part_mix((uint64_t)part_key(key, width, row)) % (unsigned)n_parts)], 1u);
  __syncthreads();
  for (int i = threadIdx.x; i < n_parts; i += blockDim.x)
    if (h[i]) atomicAdd(&counts[i], (unsigned long long)h[i]);
}

struct PartCols {
  int n_cols;
  int width[16];
  const unsigned char* src[16];
  unsigned char* dst[16];
};

// one tile of PART_TILE rows per iteration: shared-memory histogram -> one global reservation per (CTA, partition)
// -> scatter of every column
__global__ void __launch_bounds__(PART_THREADS) k_part_scatter(const void* __restrict__ key, int width, int64_t n, int n_parts,
                                                               unsigned long long* __restrict__ cursor, const __grid_constant__ PartCols pc) {
  __shared__ unsigned int h[PART_MAX];
  __shared__ unsigned long long base[PART_MAX];
  const int64_t n_tiles = (n + PART_TILE - 1) / PART_TILE;
  for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    for (int i = threadIdx.x; i < n_parts; i += blockDim.x) h[i] = 0;
    __syncthreads();
    unsigned part[PART_ITEMS], rank[PART_ITEMS];
#pragma unroll
    for (int j = 0; j < PART_ITEMS; ++j) {
      const int64_t row = t * PART_TILE + (int64_t)j * PART_THREADS + threadIdx.x;
      part[j] = 0xffffffffu;
      if (row < n) {
        part[j] = (unsigned)(part_mix((uint64_t)part_key(key, width, row)) % (unsigned)n_parts);
        rank[j] = atomicAdd(&h[part[j]], 1u);
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_parts; i += blockDim.x)
      if (h[i]) base[i] = atomicAdd(&cursor[i], (unsigned long long)h[i]);
    __syncthreads();
#pragma unroll
    for (int j = 0; j < PART_ITEMS; ++j) {
      if (part[j] == 0xffffffffu) continue;
      const int64_t row = t * PART_TILE + (int64_t)j * PART_THREADS + threadIdx.x;
      const unsigned long long d = base[part[j]] + rank[j];
      for (int c = 0; c < pc.n_cols; ++c) {
        if (pc.width[c] == 16) ((ulonglong2*)pc.dst[c])[d] = ((const ulonglong2*)pc.src[c])[row];
        else if (pc.width[c] == 8) ((unsigned long long*)pc.dst[c])[d] = ((const unsigned long long*)pc.src[c])[row];
        else if (pc.width[c] == 4) ((unsigned int*)pc.dst[c])[d] = ((const unsigned int*)pc.src[c])[row];
        else if (pc.width[c] == 2) ((unsigned short*)pc.dst[c])[d] = ((const unsigned short*)pc.src[c])[row];
        else pc.dst[c][d] = pc.src[c][row];
      }
    }
    __syncthreads();
  }
}

std::shared_ptr<TableImpl> hash_partition_table(TableImpl& t, int key_col, int n_parts, std::vector<int64_t>& offsets) {
  Ctx* ctx = t.ctx;
  t.consolidate();
  if (n_parts < 1 || n_parts > PART_MAX) throw_internal("n_parts must be in [1, 1024]");
  if (key_col < 0 || key_col >= (int)t.cols.size() || !t.cols[key_col]) throw_internal("partition key column is not resident");
  const int64_t n = t.num_rows;
  const DCol& kc = *t.cols[key_col];
  if (!(kc.phys == PH_I64 || kc.phys == PH_I32 || kc.phys == PH_D64 || kc.phys == PH_U64 || kc.phys == PH_U32) || kc.null_count != 0)
    throw_internal("hash partitioning needs a NULL-free 32/64-bit integer, date or narrowed decimal key column");
  PartCols pc;
  memset(&pc, 0, sizeof(pc));
  auto out = std::make_shared<TableImpl>();
  out->ctx = ctx;
  out->schema = t.schema;
  out->num_rows = n;
  out->num_batches = t.num_batches > 0 ? 1 : 0;
  out->cols.assign(t.cols.size(), nullptr);
  for (size_t c = 0; c < t.cols.size(); ++c) {
    if (!t.cols[c]) continue;
    const DCol& s = *t.cols[c];
    const int w = phys_width(s.phys);
    if (w == 0 || s.null_count != 0) throw_internal("hash partitioning supports NULL-free fixed-width columns only (column '" + t.schema.fields[c].name + "')");
    if (pc.n_cols >= 16) throw_internal("hash partitioning supports at most 16 resident columns");
    auto d = std::make_shared<DCol>(s);
    d->data = ctx->alloc(std::max<size_t>((size_t)n * w, 16));
    d->validity.reset();
    d->dict_state = 0;
    d->dict_codes.reset();
    pc.width[pc.n_cols] = w;
    pc.src[pc.n_cols] = (const unsigned char*)s.data->ptr;
    pc.dst[pc.n_cols] = (unsigned char*)d->data->ptr;
    pc.n_cols++;
    out->cols[c] = d;
  }
  DBufP counts = ctx->alloc_zero((size_t)n_parts * 8), cursor = ctx->alloc((size_t)n_parts * 8);
  std::vector<unsigned long long> h((size_t)n_parts, 0);
  if (n > 0) {
    LAUNCH(ctx, k_part_hist, grid_for(ctx, n, PART_THREADS * 8), PART_THREADS, 0, kc.data->ptr, phys_width(kc.phys), n, n_parts,
           (unsigned long long*)counts->ptr);
    ctx->d2h_sync(h.data(), counts->ptr, (size_t)n_parts * 8);
  }
  offsets.assign((size_t)n_parts + 1, 0);
  for (int p = 0; p < n_parts; ++p) offsets[p + 1] = offsets[p] + (int64_t)h[p];
  std::vector<unsigned long long> start(offsets.begin(), offsets.end() - 1);
  ctx->h2d(cursor->ptr, start.data(), (size_t)n_parts * 8);
  ctx->sync();
  if (n > 0)
    LAUNCH(ctx, k_part_scatter, grid_for(ctx, n, PART_TILE), PART_THREADS, 0, kc.data->ptr, phys_width(kc.phys), n, n_parts,
           (unsigned long long*)cursor->ptr, pc);
  return out;
}

}  // namespace qgpu
