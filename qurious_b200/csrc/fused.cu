// Fused Scan -> pushed-down predicate -> (group-by) aggregate pipeline: ONE kernel over the resident
// columns of a table, no selection vector, no materialised temporaries.  This is the hot path of
// TPC-H Q6 / Q1 and of every HashAggregate/NoGroupingAggregate whose input is a
// (filtered) MemoryTable scan:
//   MemoryTable::scan + filter        qurious/src/datasource/memory.rs:69-98
//   Filter::execute                   qurious/src/physical/plan/filter.rs:28-44
//   BinaryExpr / CastExpr / Literal   qurious/src/physical/expr/{binary,cast,literal}.rs
//   NoGroupingAggregate::execute      qurious/src/physical/plan/aggregate/no_grouping.rs:30-62
//   HashAggregate / GroupAccumulator  qurious/src/physical/plan/aggregate/hash.rs:45-107,138-170
//   Sum/Avg/Count/Min/Max             qurious/src/physical/expr/aggregate/*.rs
//
// Design (DESIGN.md "Fused scan-aggregate kernel"):
//   * column tiles (1024 rows of every referenced column, in the narrowed resident layout) are staged
//     into shared memory with 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx), 2-4 stages
//     deep, one persistent CTA per SM; operands of the predicate / key / aggregate expressions are then
//     addressed dynamically in shared memory, so one kernel serves every plan of the supported shape;
//   * predicate  = conjunction of integer range tests (constants folded once on the host);
//   * aggregates = SUM / MIN / MAX / COUNT / AVG over products of affine terms (a + b*column), exact in
//     int64 per row -- proven from the column statistics, otherwise the generic i128 path runs -- with
//     128-bit group totals;
//   * grouping   = (a) DENSE: small key domain (dictionary codes / small integer ranges): every thread
//     owns a private accumulator table in shared memory (no atomics, no inter-lane traffic per row),
//     reduced once per CTA into the 128-bit global table; (b) HASH: HBM-resident open-addressing table
//     on the packed key, claimed with atomicCAS, accumulators updated with 64-bit atomics (+ carry).
// Anything outside this shape (NULLs, OR/!=, wide decimals, computed keys, ...) returns false and the
// generic interpreter path (ops.cu) runs instead: same results, lower speed.
#include <algorithm>
#include <unistd.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "comm.h"
#include "epilogue.h"
#include "launch.h"
#include "plan.h"

#include "fused_device.cuh"
#include "fused_jit.h"

namespace qgpu {

// DENSE keeps one CTA per SM (its private tables fill shared memory); the table-probing modes run two (HASH) or
// three (join probes) CTAs per SM to hide the latency of their random HBM/L2 accesses
template <int MODE>
__global__ void __launch_bounds__(F_NT + 32, MODE == FM_DENSE ? 1 : (MODE == FM_HASH ? 2 : 3)) k_fused_scan_agg(const __grid_constant__ FParams p) {  // FM_BUILD: like the probes
  fused_main<MODE, GenericBody<MODE>>(p);
}

#include "radix_agg.cuh"

template <uint64_t S0, uint64_t S1, uint64_t S2, uint64_t S3, uint32_t PACK>
__global__ void __launch_bounds__(F_NT + 32, 1) k_fused_scan_agg_spec(const __grid_constant__ FParams p) {
  fused_main<FM_DENSE, SpecBody<S0, S1, S2, S3, PACK>>(p);
}

// ---- registered shapes --------------------------------------------------------------------------------
// TPC-H Q1: 1 Date32 range; keys = 2 dictionary-coded Utf8 columns; SUM(qty), SUM(price),
// SUM(price*(100-disc)) chained on price, SUM(..*(100+tax)) chained, SUM(disc)  [AVGs share the sums]
#define SIG_Q1                                                                                                                     \
  sig_head(1, sig_list(WC_I32), 2, sig_list(WC_U8, WC_U8), 5),                                                                       \
      sig_accs3(sig_acc(FK_SUM, false, true, 1, sig_factor(WC_I64, true)), sig_acc(FK_SUM, false, true, 1, sig_factor(WC_I64, true)), \
                sig_acc(FK_SUM, true, false, 1, sig_factor(WC_I64, false))),                                                         \
      sig_accs3(sig_acc(FK_SUM, true, false, 1, sig_factor(WC_I64, false)), sig_acc(FK_SUM, false, true, 1, sig_factor(WC_I64, true))), 0
// TPC-H Q6: ranges on Date32, Decimal(15,2), Decimal(15,2); no keys; SUM(price * disc)
#define SIG_Q6                                                       \
  sig_head(3, sig_list(WC_I32, WC_I64, WC_I64), 0, 0, 1),            \
      sig_accs3(sig_acc(FK_SUM, false, true, 2, sig_factor(WC_I64, true), sig_factor(WC_I64, true))), 0, 0

static FSig make_sig(const FParams& P) {
  FSig g;
  uint64_t preds = 0, keys = 0;
  for (int k = 0; k < P.n_pred; ++k) preds |= (uint64_t)wk_to_code(P.pred[k].wk) << (3 * k);
  for (int k = 0; k < P.n_keys; ++k) keys |= (uint64_t)wk_to_code(P.keys[k].wk) << (3 * k);
  g.s[0] = sig_head(P.n_pred, preds, P.n_keys, keys, P.n_accs);
  g.s[1] = g.s[2] = g.s[3] = 0;
  for (int k = 0; k < P.n_accs; ++k) {
    const FAcc& a = P.accs[k];
    uint64_t f[3] = {0, 0, 0};
    for (int i = 0; i < a.n_factors; ++i) f[i] = sig_factor(wk_to_code(a.f[i].wk), a.f[i].plain != 0);
    g.s[1 + k / 3] |= sig_acc(a.kind, a.chain != 0, a.unit != 0, a.n_factors, f[0], f[1], f[2]) << (18 * (k % 3));
  }
  return g;
}

typedef void (*FusedKernel)(const FParams);
// PACK: accumulators kept as bit fields of the count word; usable only when every one of them is in `safe_mask`
template <uint32_t PACK, uint64_t S0, uint64_t S1, uint64_t S2, uint64_t S3>
static bool sig_matches(const FSig& g, uint32_t safe_mask, FusedKernel* out, uint32_t* pack) {
  if (g.s[0] != S0 || g.s[1] != S1 || g.s[2] != S2 || g.s[3] != S3 || (PACK & ~safe_mask) != 0) return false;
  *out = k_fused_scan_agg_spec<S0, S1, S2, S3, PACK>;
  *pack = PACK;
  return true;
}
// returns the specialised DENSE kernel registered for this shape (and the accumulators it packs), or nullptr
static FusedKernel find_specialised(const FParams& P, uint32_t safe_mask, uint32_t* pack) {
  *pack = 0;
  if (P.mode != FM_DENSE) return nullptr;
  const FSig g = make_sig(P);
  FusedKernel k = nullptr;
  if (sig_matches<0x11, SIG_Q1>(g, safe_mask, &k, pack)) return k;  // SUM(l_quantity), SUM(l_discount) ride in the count word
  if (sig_matches<0, SIG_Q1>(g, safe_mask, &k, pack)) return k;
  if (sig_matches<0, SIG_Q6>(g, safe_mask, &k, pack)) return k;
  if (getenv("QGPU_FUSED_DEBUG"))
    fprintf(stderr, "[qgpu] fused shape without a specialised kernel: %016llx %016llx %016llx %016llx\n", (unsigned long long)g.s[0],
            (unsigned long long)g.s[1], (unsigned long long)g.s[2], (unsigned long long)g.s[3]);
  return nullptr;
}

// ------------------------------------------------------------------------------------------------
// global table init + export to the GroupAccs layout consumed by finish_aggregate (ops.cu)
// ------------------------------------------------------------------------------------------------
struct FInit {
  long long v[F_MAXA + 2];
};
// layout 0: [k][slot] (PROBE), 1: [slot][k] (DENSE), 2: records of `rec` words {key = EMPTY, acc[0..na2), pad} (HASH)
__global__ void k_fused_init(unsigned long long* lo, unsigned long long* hi, int64_t n_slots, int na2, int layout, int rec, FInit init) {
  const int64_t total = layout == 2 ? n_slots * rec : n_slots * na2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    unsigned long long v;
    if (layout == 2) {
      const int w = (int)(i % rec);
      v = w == 0 ? F_EMPTY : (w <= na2 ? (unsigned long long)init.v[w - 1] : 0ull);
    } else {
      v = (unsigned long long)init.v[layout == 1 ? (int)(i % na2) : (int)(i / n_slots)];
    }
    lo[i] = v;
    if (hi) hi[i] = 0;
  }
}

__global__ void k_fused_occupied(const unsigned long long* __restrict__ lo, int64_t n_slots, int64_t cnt_base, int64_t cnt_stride,
                                 int64_t* __restrict__ flags) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n_slots; g += stride)
    flags[g] = lo[cnt_base + g * cnt_stride] != 0 ? 1 : 0;
}

struct FExport {
  int n_aggs;
  int acc_of[24];   // accumulator index per aggregate (-1: COUNT only)
  int kind_of[24];  // FK_* of that accumulator
  int wide[24];     // 1: the accumulator has a real hi word (DENSE sums / carry mode)
  unsigned long long* out_lo[24];
  unsigned long long* out_hi[24];
  // ungrouped aggregate over zero qualifying rows: MIN / MAX return the TYPED start value (min.rs / max.rs NATIVE::MAX / MIN,
  // quirk Q4) -- the kernels' working seed is INT64_MAX / INT64_MIN whatever the type
  unsigned long long sent_lo[24], sent_hi[24];
};
__global__ void k_fused_export(const unsigned long long* __restrict__ lo, const unsigned long long* __restrict__ hi, int64_t n_slots,
                               int64_t k_stride, int64_t g_stride, int n_accs, const int64_t* __restrict__ flags,
                               const int64_t* __restrict__ offs, FExport ex, unsigned long long* __restrict__ out_cnt,
                               long long* __restrict__ out_first) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n_slots; g += stride) {
    if (!flags[g]) continue;
    const int64_t o = offs[g];
    out_cnt[o] = lo[(int64_t)n_accs * k_stride + g * g_stride];
    out_first[o] = (long long)lo[(int64_t)(n_accs + 1) * k_stride + g * g_stride];
    for (int a = 0; a < ex.n_aggs; ++a) {
      const int k = ex.acc_of[a];
      if (k < 0) continue;
      const unsigned long long l = lo[(int64_t)k * k_stride + g * g_stride];
      unsigned long long h = 0;
      if (ex.wide[a]) h = hi[(int64_t)k * k_stride + g * g_stride];
      else if (ex.kind_of[a] != FK_SUMF) h = ((long long)l < 0) ? ~0ull : 0ull;
      ex.out_lo[a][o] = l;
      ex.out_hi[a][o] = h;
    }
  }
}

// Small tables (DENSE mode: <= 4096 slots): occupied flags, prefix sum, compaction and export in ONE CTA; the
// group count stays on the device (finish_aggregate fetches it together with its own flags).
__global__ void __launch_bounds__(1024) k_fused_export_small(const unsigned long long* __restrict__ lo,
                                                             const unsigned long long* __restrict__ hi, int n_slots, int64_t k_stride,
                                                             int64_t g_stride, int n_accs, int force_all, FExport ex,
                                                             unsigned long long* __restrict__ out_cnt, long long* __restrict__ out_first,
                                                             long long* __restrict__ n_groups_out) {
  __shared__ int pre[1024];
  const int tid = threadIdx.x;
  const int per = (n_slots + 1023) / 1024;
  const int base = tid * per;
  int c = 0;
  unsigned mask = 0;
  for (int i = 0; i < per; ++i) {
    const int g = base + i;
    const bool occ = g < n_slots && (force_all || lo[(int64_t)n_accs * k_stride + g * g_stride] != 0);
    c += occ;
    mask |= (unsigned)occ << i;
  }
  pre[tid] = c;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {
    const int v = tid >= d ? pre[tid - d] : 0;
    __syncthreads();
    pre[tid] += v;
    __syncthreads();
  }
  int o = pre[tid] - c;
  if (tid == 1023) *n_groups_out = pre[1023];
  for (int i = 0; i < per; ++i) {
    if (!((mask >> i) & 1)) continue;
    const int64_t g = base + i;
    const unsigned long long rows_g = lo[(int64_t)n_accs * k_stride + g * g_stride];
    out_cnt[o] = rows_g;
    out_first[o] = (long long)lo[(int64_t)(n_accs + 1) * k_stride + g * g_stride];
    for (int a = 0; a < ex.n_aggs; ++a) {
      const int k = ex.acc_of[a];
      if (k < 0) continue;
      unsigned long long l = lo[(int64_t)k * k_stride + g * g_stride];
      unsigned long long h = 0;
      if (ex.wide[a]) h = hi[(int64_t)k * k_stride + g * g_stride];
      else if (ex.kind_of[a] != FK_SUMF) h = ((long long)l < 0) ? ~0ull : 0ull;
      if (rows_g == 0 && (ex.kind_of[a] == FK_MIN || ex.kind_of[a] == FK_MAX)) {  // force_all: the ungrouped empty case
        l = ex.sent_lo[a];
        h = ex.sent_hi[a];
      }
      ex.out_lo[a][o] = l;
      ex.out_hi[a][o] = h;
    }
    ++o;
  }
}

// ------------------------------------------------------------------------------------------------
// dictionary encoding of low-cardinality Utf8 columns (<= 255 distinct values)
// ------------------------------------------------------------------------------------------------
#define DICT_CAP 1024
__device__ __forceinline__ uint64_t str_hash(const char* s, int len) {
  uint64_t x = 0xcbf29ce484222325ULL ^ (uint64_t)len;
  for (int i = 0; i < len; ++i) x = (x ^ (unsigned char)s[i]) * 0x100000001b3ULL;
  return fmix64(x);
}
// slots[s] = (tag32 << 32) | (rep_row + 1); row_slot[row] = s
__device__ __forceinline__ uint32_t dict_global_slot(const int32_t* __restrict__ offs, const char* __restrict__ data, int64_t row,
                                                     int32_t o0, int32_t len, unsigned long long* __restrict__ slots,
                                                     unsigned int* __restrict__ n_dict, int* __restrict__ abort_flag) {
  const uint64_t h = str_hash(data + o0, len);
  const uint32_t tag = (uint32_t)(h >> 32);
  uint32_t s = (uint32_t)h & (DICT_CAP - 1);
  const unsigned long long mine = ((unsigned long long)tag << 32) | (unsigned long long)(row + 1);
  // the walk is bounded: once more than 255 values exist the column is not dictionary material (abort), and rows already
  // in flight must not keep filling -- or circling -- the 1024-slot table
  for (int probes = 0; probes < DICT_CAP; ++probes) {
    if (*(volatile int*)abort_flag) return 0;
    unsigned long long cur = *(volatile unsigned long long*)&slots[s];
    if (cur == 0) {
      cur = atomicCAS(&slots[s], 0ull, mine);
      if (cur == 0) {
        if (atomicAdd(n_dict, 1u) >= 255u) *abort_flag = 1;
        return s;
      }
    }
    if ((uint32_t)(cur >> 32) == tag) {
      const int64_t rep = (int64_t)(cur & 0xffffffffull) - 1;
      const int32_t r0 = offs[rep], rl = offs[rep + 1] - r0;
      bool eq = rl == len;
      for (int i = 0; eq && i < len; ++i) eq = data[r0 + i] == data[o0 + i];
      if (eq) return s;
    }
    s = (s + 1) & (DICT_CAP - 1);
  }
  *abort_flag = 1;  // table full
  return 0;
}
// Values of <= 8 bytes (flags, status codes, most segment / mode names' prefixes do not count: the WHOLE value must fit) are
// their own key: a per-CTA shared-memory table maps (length, packed bytes) -> global slot, so that after a CTA's first
// meeting with a value its rows cost one coalesced offsets / data read, one shared-memory probe and a 2-byte store -- no
// dependent chain through the global table and the representative row (0.65 ms -> see DESIGN 3.3 for 60 M rows).
#define DICT_LCAP 512
#define DICT_L_EMPTY 0xffffffffu
#define DICT_L_BUSY 0xfffffffeu
__global__ void __launch_bounds__(256) k_dict_insert(const int32_t* __restrict__ offs, const char* __restrict__ data, int64_t n,
                                                     unsigned long long* __restrict__ slots, uint16_t* __restrict__ row_slot,
                                                     unsigned int* __restrict__ n_dict, int* __restrict__ abort_flag) {
  __shared__ unsigned long long l_key[DICT_LCAP];
  __shared__ unsigned int l_val[DICT_LCAP];  // (len << 16) | global slot
  for (int i = threadIdx.x; i < DICT_LCAP; i += blockDim.x) l_val[i] = DICT_L_EMPTY;
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  // four rows per iteration: their offsets and (short) values are loaded before the first table probe, so that the
  // dependent chain offsets -> bytes -> probe of one row overlaps the others'; the abort flag (too many distinct values)
  // is polled once per iteration, not per row
  constexpr int U = 4;
  for (int64_t row0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row0 < n; row0 += stride * U) {
    if (*(volatile int*)abort_flag) return;
    int32_t o0[U], len[U];
    unsigned long long key[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t row = row0 + (int64_t)u * stride;
      o0[u] = 0;
      len[u] = -1;
      if (row < n) {
        o0[u] = offs[row];
        len[u] = offs[row + 1] - o0[u];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      key[u] = 0;
      if (len[u] >= 0 && len[u] <= 8)
        for (int i = 0; i < len[u]; ++i) key[u] |= (unsigned long long)(unsigned char)data[o0[u] + i] << (8 * i);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (len[u] < 0) continue;
      const int64_t row = row0 + (int64_t)u * stride;
      uint32_t s;
      if (len[u] <= 8) {
        uint32_t li = (uint32_t)(fmix64(key[u] + (unsigned long long)len[u]) >> 40) & (DICT_LCAP - 1);
        bool done = false;
        for (int probe = 0; probe < 8 && !done; ++probe, li = (li + 1) & (DICT_LCAP - 1)) {
          unsigned int v = *(volatile unsigned int*)&l_val[li];
          if (v == DICT_L_EMPTY) {
            v = atomicCAS(&l_val[li], DICT_L_EMPTY, DICT_L_BUSY);
            if (v == DICT_L_EMPTY) {  // claimed: resolve through the global table once, then publish
              s = dict_global_slot(offs, data, row, o0[u], len[u], slots, n_dict, abort_flag);
              l_key[li] = key[u];
              __threadfence_block();
              *(volatile unsigned int*)&l_val[li] = ((unsigned int)len[u] << 16) | s;
              done = true;
              break;
            }
          }
          if (v == DICT_L_BUSY) break;  // being published by another thread: take the global path, no spinning
          if ((v >> 16) == (unsigned int)len[u] && *(volatile unsigned long long*)&l_key[li] == key[u]) {
            s = v & 0xffffu;
            done = true;
          }
        }
        if (!done) s = dict_global_slot(offs, data, row, o0[u], len[u], slots, n_dict, abort_flag);
      } else {
        s = dict_global_slot(offs, data, row, o0[u], len[u], slots, n_dict, abort_flag);
      }
      row_slot[row] = (uint16_t)s;
    }
  }
}
__global__ void k_dict_codes(const uint16_t* __restrict__ row_slot, const uint8_t* __restrict__ slot_code, int64_t n,
                             uint8_t* __restrict__ codes) {
  __shared__ uint8_t lut[DICT_CAP];
  for (int i = threadIdx.x; i < DICT_CAP; i += blockDim.x) lut[i] = slot_code[i];
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) codes[i] = lut[row_slot[i]];
}

// Build (once per resident column) the dictionary encoding; returns false when the column has more
// than 255 distinct values, NULLs, or rows beyond the u32 representative-row range.
static bool ensure_dict(Ctx* ctx, DCol& col) {
  if (col.dict_state != 0) return col.dict_state > 0;
  SpecSuspend cached_work(ctx);
  col.dict_state = -1;
  if (col.phys != PH_STR || col.null_count != 0 || col.length == 0 || col.length >= 0xfffffff0LL) return false;
  const int64_t n = col.length;
  DBufP slots = ctx->alloc_zero(DICT_CAP * 8);
  DBufP row_slot = ctx->alloc((size_t)n * 2);
  DBufP flags = ctx->alloc_zero(8);
  LAUNCH(ctx, k_dict_insert, grid_for(ctx, n, 256 * 4), 256, 0, (const int32_t*)col.offsets->ptr, (const char*)col.data->ptr, n,
         (unsigned long long*)slots->ptr, (uint16_t*)row_slot->ptr, (unsigned int*)flags->ptr, (int*)((char*)flags->ptr + 4));
  struct { unsigned int n; int abort_; } h;
  ctx->d2h_sync(&h, flags->ptr, 8);
  if (h.abort_ || h.n > 255) return false;
  std::vector<unsigned long long> hs(DICT_CAP);
  ctx->d2h_sync(hs.data(), slots->ptr, DICT_CAP * 8);
  // codes in ascending representative-row order (deterministic); fetch the strings of the representatives
  std::vector<std::pair<long long, int>> reps;
  for (int s = 0; s < DICT_CAP; ++s)
    if (hs[s]) reps.push_back({(long long)(hs[s] & 0xffffffffull) - 1, s});
  std::sort(reps.begin(), reps.end());
  std::vector<uint8_t> slot_code(DICT_CAP, 0);
  std::vector<long long> rep_rows;
  for (size_t i = 0; i < reps.size(); ++i) {
    slot_code[reps[i].second] = (uint8_t)i;
    rep_rows.push_back(reps[i].first);
  }
  DBufP idx = ctx->alloc(rep_rows.size() * 8);
  ctx->h2d(idx->ptr, rep_rows.data(), rep_rows.size() * 8);
  ctx->sync();
  DColP taken = take_column(ctx, col, (const int64_t*)idx->ptr, (int64_t)rep_rows.size());
  std::vector<int32_t> toffs(rep_rows.size() + 1);
  ctx->d2h_sync(toffs.data(), taken->offsets->ptr, toffs.size() * 4);
  std::vector<char> tdata((size_t)std::max<int64_t>(taken->str_bytes, 1));
  if (taken->str_bytes > 0) ctx->d2h_sync(tdata.data(), taken->data->ptr, (size_t)taken->str_bytes);
  col.dict_values.clear();
  for (size_t i = 0; i < rep_rows.size(); ++i) col.dict_values.emplace_back(tdata.data() + toffs[i], (size_t)(toffs[i + 1] - toffs[i]));
  DBufP sc = ctx->alloc(DICT_CAP);
  ctx->h2d(sc->ptr, slot_code.data(), DICT_CAP);
  ctx->sync();
  col.dict_codes = ctx->alloc((size_t)n);
  LAUNCH(ctx, k_dict_codes, grid_for(ctx, n, 256 * 8), 256, 0, (const uint16_t*)row_slot->ptr, (const uint8_t*)sc->ptr, n,
         (uint8_t*)col.dict_codes->ptr);
  col.dict_state = 1;
  return true;
}

// ------------------------------------------------------------------------------------------------
// host-side plan analysis
// ------------------------------------------------------------------------------------------------
namespace {

struct Unsupported {};  // thrown to leave the fused path (the generic path then runs)

const i128 LIM62 = (i128)1 << 62;

struct Slot {
  DCol* col;
  bool dict;
  i128 vmin, vmax;
};

struct Poly {                      // coef * prod(a_i + b_i * x_i)
  i128 coef = 1;
  std::vector<FFactor> fs;
  i128 lo = 1, hi = 1;             // value range
  DType type;
};

struct Analyzer {
  Ctx* ctx;
  const View& v;
  std::vector<Slot> slots;
  std::vector<int> slot_view_col;

  int slot_for(int view_col, bool want_dict) {
    if (view_col < 0 || view_col >= (int)v.cols.size()) throw Unsupported();
    const LazyCol& lc = v.cols[view_col];
    if (!lc.base || lc.idx) throw Unsupported();
    DCol& c = *lc.base;
    if (c.null_count != 0 || c.length != v.num_rows) throw Unsupported();
    for (size_t i = 0; i < slots.size(); ++i)
      if (slots[i].col == &c && slots[i].dict == want_dict) return (int)i;
    Slot s;
    s.col = &c;
    s.dict = want_dict;
    if (want_dict) {
      if (!ensure_dict(ctx, c)) throw Unsupported();
      s.vmin = 0;
      s.vmax = (i128)c.dict_values.size() - 1;
    } else {
      switch (c.phys) {
        case PH_I8: case PH_I16: case PH_I32: case PH_I64: case PH_D64: case PH_U8: case PH_U16: case PH_U32: break;
        default: throw Unsupported();
      }
      ensure_stats(ctx, c);
      if (!c.has_stats) throw Unsupported();
      s.vmin = c.vmin;
      s.vmax = c.vmax;
    }
    if ((int)slots.size() >= F_MAXC) throw Unsupported();
    slots.push_back(s);
    slot_view_col.push_back(view_col);
    return (int)slots.size() - 1;
  }
  int f64_slot_for(int view_col) {
    if (view_col < 0 || view_col >= (int)v.cols.size()) throw Unsupported();
    const LazyCol& lc = v.cols[view_col];
    if (!lc.base || lc.idx) throw Unsupported();
    DCol& c = *lc.base;
    if (c.null_count != 0 || c.phys != PH_F64 || c.length != v.num_rows) throw Unsupported();
    for (size_t i = 0; i < slots.size(); ++i)
      if (slots[i].col == &c) return (int)i;
    if ((int)slots.size() >= F_MAXC) throw Unsupported();
    slots.push_back({&c, false, 0, 0});
    slot_view_col.push_back(view_col);
    return (int)slots.size() - 1;
  }

  static bool const_i128(const Compiled& c, i128* out) {
    if (!c.is_const || !c.const_val.valid) return false;
    switch (class_of(c.result_type)) {
      case VC_INT: *out = (i128)(int64_t)c.const_val.lo; return true;
      case VC_UINT: case VC_BOOL: *out = (i128)(uint64_t)c.const_val.lo; return true;
      case VC_DEC: *out = val_i128(c.const_val); return true;
      default: return false;
    }
  }

  static void check_range(i128 lo, i128 hi) {
    if (lo <= -LIM62 || hi >= LIM62) throw Unsupported();
  }
  static void mul_range(i128 alo, i128 ahi, i128 blo, i128 bhi, i128* lo, i128* hi) {
    check_range(alo, ahi);
    check_range(blo, bhi);
    const i128 c[4] = {alo * blo, alo * bhi, ahi * blo, ahi * bhi};
    *lo = std::min(std::min(c[0], c[1]), std::min(c[2], c[3]));
    *hi = std::max(std::max(c[0], c[1]), std::max(c[2], c[3]));
    check_range(*lo, *hi);
  }
  // the declared integer width must not wrap (narrow ints wrap in the reference; we only take the proven-safe case)
  static void check_width(const DType& t, i128 lo, i128 hi) {
    if (t.is_signed_int() || t.is_date()) {
      const int bits = arrow_width(t) * 8;
      const i128 lim = (i128)1 << (bits - 1);
      if (lo < -lim || hi >= lim) throw Unsupported();
    } else if (t.is_unsigned_int()) {
      const int bits = arrow_width(t) * 8;
      if (lo < 0 || hi >= ((i128)1 << bits)) throw Unsupported();
    }
  }

  Poly analyze(const ExprNode& n) {
    auto c = compile_expr(n, v.schema);
    if (c->deferred_err) throw Unsupported();
    Poly p;
    p.type = c->result_type;
    const VClass vc = class_of(c->result_type);
    if (vc != VC_INT && vc != VC_UINT && vc != VC_DEC) throw Unsupported();
    i128 cv;
    if (c->is_const) {
      if (!const_i128(*c, &cv)) throw Unsupported();
      check_range(cv, cv);
      p.coef = cv;
      p.lo = p.hi = cv;
      return p;
    }
    switch (n.kind) {
      case QGPU_IR_COLUMN: {
        const int s = slot_for(n.col_index, false);
        FFactor f;
        memset(&f, 0, sizeof(f));
        f.col = s;
        f.a = 0;
        f.b = 1;
        p.fs.push_back(f);
        p.lo = slots[s].vmin;
        p.hi = slots[s].vmax;
        check_range(p.lo, p.hi);
        return p;
      }
      case QGPU_IR_NEGATIVE: {
        Poly q = analyze(*n.children[0]);
        q.coef = -q.coef;
        const i128 lo = -q.hi, hi = -q.lo;
        q.lo = lo;
        q.hi = hi;
        q.type = p.type;
        check_width(p.type, q.lo, q.hi);
        return q;
      }
      case QGPU_IR_BINARY: {
        Poly l = analyze(*n.children[0]);
        Poly r = analyze(*n.children[1]);
        if (n.op == 10) {
          p.coef = l.coef * r.coef;
          check_range(p.coef, p.coef);
          p.fs = l.fs;
          p.fs.insert(p.fs.end(), r.fs.begin(), r.fs.end());
          mul_range(l.lo, l.hi, r.lo, r.hi, &p.lo, &p.hi);
          check_width(p.type, p.lo, p.hi);
          return p;
        }
        if (n.op != 8 && n.op != 9) throw Unsupported();
        // affine (+/-): each side is a constant or coef * (a + b x) with a single factor
        i128 ml = 1, mr = 1;
        if (vc == VC_DEC) {
          ml = pow10_i128(p.type.scale - l.type.scale);
          mr = pow10_i128(p.type.scale - r.type.scale);
        }
        auto affine = [&](const Poly& q, i128 m, i128* a, i128* b, int* col) {
          if (q.fs.size() > 1) throw Unsupported();
          if (q.fs.empty()) {
            *a = q.coef * m;
            *b = 0;
            *col = -1;
          } else {
            *a = q.coef * m * q.fs[0].a;
            *b = q.coef * m * q.fs[0].b;
            *col = q.fs[0].col;
          }
          check_range(*a, *a);
          check_range(*b, *b);
        };
        i128 al, bl, ar, br;
        int cl, cr;
        affine(l, ml, &al, &bl, &cl);
        affine(r, mr, &ar, &br, &cr);
        if (cl >= 0 && cr >= 0 && cl != cr) throw Unsupported();
        const int sgn = n.op == 8 ? 1 : -1;
        const i128 a = al + sgn * ar, b = bl + sgn * br;
        const int col = cl >= 0 ? cl : cr;
        check_range(a, a);
        check_range(b, b);
        i128 lo = l.lo * ml, hi = l.hi * ml, rlo = r.lo * mr, rhi = r.hi * mr;
        check_range(lo, hi);
        check_range(rlo, rhi);
        p.lo = sgn > 0 ? lo + rlo : lo - rhi;
        p.hi = sgn > 0 ? hi + rhi : hi - rlo;
        check_range(p.lo, p.hi);
        check_width(p.type, p.lo, p.hi);
        if (col < 0 || b == 0) {
          p.coef = a;
          p.lo = p.hi = a;
          return p;
        }
        FFactor f;
        memset(&f, 0, sizeof(f));
        f.col = col;
        f.a = (int64_t)a;
        f.b = (int64_t)b;
        p.coef = 1;
        p.fs.push_back(f);
        return p;
      }
      default: throw Unsupported();
    }
  }
};

void flatten_and(const ExprNode& n, std::vector<const ExprNode*>& out) {
  if (n.kind == QGPU_IR_BINARY && n.op == 6) {
    flatten_and(*n.children[0], out);
    flatten_and(*n.children[1], out);
  } else {
    out.push_back(&n);
  }
}

struct Range {
  i128 lo, hi;
};

}  // namespace

namespace {
// Everything the analysis of one Aggregate <- (Filter)* <- Scan subtree produces; cached on the plan node and
// re-used by later executes while the table's resident columns are unchanged.
struct FusedPlan {
  FParams P;
  std::vector<std::shared_ptr<Compiled>> keys;
  std::vector<AggSpec> specs;
  std::vector<int> acc_of;
  int total_bits = 0;
  size_t smem_bytes = 0;
  int grid = 0;
  FusedKernel spec = nullptr;
  JitKernel jit;            // DENSE shapes without a registered kernel: the same SpecBody instantiated at run time (fused_jit.cu)
  int64_t learned_cap = 0;  // HASH: capacity that held every group last time (skips the growth retries)
  // RADIX (radix_agg.cuh): usable when the plan is a HASH-mode plan with integer-like keys and few operand values
  bool radix_ok = false;
  bool radix_failed = false;  // overflowed or declined (few groups) once on this table: stay on FM_HASH
  std::shared_ptr<void> exchange;  // RadixExchange: state of a multi-GPU exchange in flight + its receive buffers
  std::shared_ptr<void> dense;     // DenseRun: persistent accumulator table + epilogue description (DENSE mode)
  RParams R;                  // comps / comp_of / kind_of filled by the analysis
  int key_bits[F_MAXK] = {0, 0, 0, 0}, key_shift[F_MAXK] = {0, 0, 0, 0}, key_width[F_MAXK] = {0, 0, 0, 0};
  int64_t key_min[F_MAXK] = {0, 0, 0, 0}, key_max[F_MAXK] = {0, 0, 0, 0};  // value range of every key column (this table)
  std::vector<DColP> key_src;  // source column of every key (type / phys of the decoded key column)
  // validity of the cache
  std::vector<const DCol*> col_ids;
  int64_t n_rows = 0, n_batches = 0;
  bool usable = false;
};
}  // namespace

namespace {
// probe-side analysis for the fused join-probe + aggregate pipeline: no group keys in the kernel (the group of a
// probe row is its matching build row); aggregate arguments are given re-based onto the probe scan's schema
struct ProbeOpts {
  bool emit = false;  // join output pairs instead of aggregation: no aggregate list
  int probe_key_col = -1;
  std::vector<std::unique_ptr<ExprNode>> agg_exprs;
};
}  // namespace

static bool analyze_fused(PlanNode& agg, const std::vector<const ExprNode*>& predicates, const View& v, FusedPlan& fp,
                          const ProbeOpts* probe = nullptr) {
  Ctx* ctx = agg.ctx;
  SpecSuspend cached_work(ctx);  // the analysis is cached on the plan node
  auto agg_expr = [&](size_t i) -> const ExprNode& { return probe ? *probe->agg_exprs[i] : *agg.aggs[i].expr; };
  const bool emit = probe && probe->emit;
  if (!emit) {
    if ((int)agg.aggs.size() > 24 || agg.aggs.empty() || (int)agg.group_exprs.size() > F_MAXK) return false;
    if (agg.schema.fields.size() != agg.group_exprs.size() + agg.aggs.size()) return false;
  }

  FParams& P = fp.P;
  memset(&P, 0, sizeof(P));
  std::vector<std::shared_ptr<Compiled>>& keys = fp.keys;
  std::vector<AggSpec>& specs = fp.specs;
  std::vector<int>& acc_of = fp.acc_of;
  keys.clear();
  specs.clear();
  fp.key_src.clear();
  acc_of.assign(agg.aggs.size(), -1);
  std::vector<i128> acc_maxabs, acc_lo;  // bounds of one row's value per accumulator
  bool dense_ok = true;
  i128 dense_groups = 1;
  int total_bits = 0;
  Analyzer A{ctx, v, {}, {}};
  try {
    // same validation (and the same errors) as the generic path
    if (!probe)
      for (auto& e : agg.group_exprs) keys.push_back(compile_expr(*e, v.schema));
    for (size_t i = 0; i < agg.aggs.size(); ++i) {
      const AggDesc& a = agg.aggs[i];
      AggSpec s;
      s.op = a.op;
      s.arg = compile_expr(agg_expr(i), v.schema);
      s.return_type = a.return_type;
      s.expr_type = a.expr_type;
      specs.push_back(s);
    }
    validate_agg_types(specs);
    for (auto& k : keys) check_hash_key_type(k->result_type);

    // ---- predicate: conjunction of range tests ------------------------------------------------------
    std::vector<const ExprNode*> terms;
    for (auto* pe : predicates) flatten_and(*pe, terms);
    std::vector<std::pair<int, Range>> ranges;
    for (const ExprNode* t : terms) {
      auto ct = compile_expr(*t, v.schema);  // type-checks the comparison exactly like the generic path
      if (ct->result_type.id != QGPU_T_BOOL) return false;
      if (t->kind != QGPU_IR_BINARY || t->op > 5) return false;
      const ExprNode* cn = t->children[0].get();
      const ExprNode* kn = t->children[1].get();
      int op = t->op;
      if (cn->kind != QGPU_IR_COLUMN) {
        std::swap(cn, kn);
        static const int flip[6] = {0, 1, 4, 5, 2, 3};  // c op x  ==  x flip(op) c
        op = flip[op];
      }
      if (cn->kind != QGPU_IR_COLUMN) return false;
      auto kc = compile_expr(*kn, v.schema);
      if (!kc->is_const || !kc->const_val.valid || kc->deferred_err) return false;
      const DType ct_type = v.schema.fields[cn->col_index].type;
      Range r{-(LIM62 * 2), LIM62 * 2 - 1};  // int64 domain
      int slot;
      if (ct_type.id == QGPU_T_UTF8) {
        if (op != 0) return false;
        slot = A.slot_for(cn->col_index, true);
        const std::string lit(kc->blob.data() + kc->const_val.lo, (size_t)kc->const_val.hi);
        const auto& dv = A.slots[slot].col->dict_values;
        auto it = std::find(dv.begin(), dv.end(), lit);
        if (it == dv.end()) r = Range{1, 0};
        else r = Range{(i128)(it - dv.begin()), (i128)(it - dv.begin())};
      } else {
        i128 c;
        if (!Analyzer::const_i128(*kc, &c)) return false;
        slot = A.slot_for(cn->col_index, false);
        switch (op) {
          case 0: r = Range{c, c}; break;
          case 2: r.lo = c + 1; break;
          case 3: r.lo = c; break;
          case 4: r.hi = c - 1; break;
          case 5: r.hi = c; break;
          default: return false;  // != is not a range
        }
      }
      bool merged = false;
      for (auto& pr : ranges)
        if (pr.first == slot) {
          pr.second.lo = std::max(pr.second.lo, r.lo);
          pr.second.hi = std::min(pr.second.hi, r.hi);
          merged = true;
        }
      if (!merged) ranges.push_back({slot, r});
    }
    if ((int)ranges.size() > F_MAXP) return false;
    for (auto& pr : ranges) {
      FPred& fp = P.pred[P.n_pred++];
      fp.col = pr.first;
      i128 lo = std::max(pr.second.lo, -(LIM62 * 2)), hi = std::min(pr.second.hi, LIM62 * 2 - 1);
      if (lo > hi) {  // empty: no row passes
        fp.lo = 1;
        fp.span = 0;
        fp.col = pr.first;
        // (x - 1) <= 0 only for x == 1; make it impossible by pairing with a second contradictory test
        if (P.n_pred >= F_MAXP) return false;
        FPred& fq = P.pred[P.n_pred++];
        fq.col = pr.first;
        fq.lo = 2;
        fq.span = 0;
      } else {
        fp.lo = (int64_t)lo;
        fp.span = (uint64_t)(hi - lo);
      }
    }

    // ---- group keys: bare columns, packed into one code ----------------------------------------------
    for (size_t i = 0; i < keys.size(); ++i) {
      if (!keys[i]->is_column_ref) return false;
      const DType kt = keys[i]->result_type;
      const int slot = A.slot_for(keys[i]->column_ref, kt.id == QGPU_T_UTF8);
      const i128 range = A.slots[slot].vmax - A.slots[slot].vmin + 1;
      FKey& fk = P.keys[P.n_keys++];
      fk.col = slot;
      fk.base = (int64_t)A.slots[slot].vmin;
      int bits = 0;
      while (((i128)1 << bits) < range) ++bits;
      // DENSE: mixed-radix index; HASH: bit-packed code (filled below once the mode is known)
      fk.mult = (uint64_t)dense_groups;  // provisional (dense)
      fk.pad = bits;
      fp.key_bits[i] = bits;
      fp.key_min[i] = (int64_t)A.slots[slot].vmin;
      fp.key_max[i] = (int64_t)A.slots[slot].vmax;
      fp.key_width[i] = A.slots[slot].dict ? 0 : phys_width(A.slots[slot].col->phys);
      fp.key_src.push_back(v.cols[keys[i]->column_ref].base);
      if (dense_ok) {
        dense_groups *= range;
        if (dense_groups > 4096) dense_ok = false;
      }
      total_bits += bits;
    }
    if (total_bits > 64) return false;
    if (probe) {  // the one kernel-side "key" is the probe join key: code = its value
      const int slot = A.slot_for(probe->probe_key_col, false);
      FKey& fk = P.keys[P.n_keys++];
      fk.col = slot;
      fk.base = 0;
      fk.mult = 1;
      dense_ok = false;
    }

    // ---- accumulators ---------------------------------------------------------------------------------
    std::vector<std::string> acc_sig;
    for (size_t i = 0; i < specs.size(); ++i) {
      AggSpec& s = specs[i];
      if (s.arg->deferred_err) return false;
      if (s.op == QGPU_AGG_COUNT) {
        // COUNT(x): x must be provably non-NULL for every row
        if (s.arg->is_const) {
          if (!s.arg->const_val.valid) return false;
        } else if (s.arg->is_column_ref) {
          const LazyCol& lc = v.cols[s.arg->column_ref];
          if (!lc.base || lc.idx || lc.base->null_count != 0) return false;
        } else {
          A.analyze(agg_expr(i));  // arithmetic over non-NULL columns is never NULL
        }
        continue;
      }
      FAcc fa;
      memset(&fa, 0, sizeof(fa));
      i128 maxabs = 0, minval = -1;
      const VClass vc = class_of(s.arg->result_type);
      if (vc == VC_FLT) {
        if (!(s.op == QGPU_AGG_SUM || s.op == QGPU_AGG_AVG) || !s.arg->is_column_ref || s.arg->result_type.id != QGPU_T_FLOAT64)
          return false;
        fa.kind = FK_SUMF;
        fa.n_factors = 1;
        fa.f[0].col = A.f64_slot_for(s.arg->column_ref);
        fa.f[0].a = 0;
        fa.f[0].b = 1;
        fa.coef = 1;
      } else {
        Poly p = A.analyze(agg_expr(i));
        if (p.fs.size() > F_MAXF) return false;
        fa.kind = (s.op == QGPU_AGG_MIN) ? FK_MIN : (s.op == QGPU_AGG_MAX ? FK_MAX : FK_SUM);
        fa.coef = (int64_t)p.coef;
        fa.n_factors = (int)p.fs.size();
        for (size_t f = 0; f < p.fs.size(); ++f) fa.f[f] = p.fs[f];
        // every prefix product must stay inside int64
        i128 lo = p.coef, hi = p.coef;
        for (size_t f = 0; f < p.fs.size(); ++f) {
          const Slot& sl = A.slots[p.fs[f].col];
          i128 flo = (i128)p.fs[f].a + (i128)p.fs[f].b * sl.vmin, fhi = (i128)p.fs[f].a + (i128)p.fs[f].b * sl.vmax;
          if (flo > fhi) std::swap(flo, fhi);
          Analyzer::mul_range(lo, hi, flo, fhi, &lo, &hi);
        }
        maxabs = std::max(lo < 0 ? -lo : lo, hi < 0 ? -hi : hi);
        minval = lo;
      }
      std::string sig((const char*)&fa, sizeof(fa));
      int found = -1;
      for (size_t j = 0; j < acc_sig.size(); ++j)
        if (acc_sig[j] == sig) found = (int)j;
      if (found < 0) {
        if (P.n_accs >= F_MAXA) return false;
        found = P.n_accs;
        P.accs[P.n_accs++] = fa;
        acc_sig.push_back(sig);
        acc_maxabs.push_back(maxabs);
        acc_lo.push_back(minval);
      }
      acc_of[i] = found;
    }
  } catch (Unsupported&) {
    return false;
  }
  // chain detection: accumulator k = accumulator k-1 times extra factors (prefix test on the full factor lists)
  {
    std::vector<FAcc> full(P.accs, P.accs + P.n_accs);
    // RADIX tuple components: the distinct operand values (SUM(v), MIN(v), MAX(v) share one)
    memset(&fp.R, 0, sizeof(fp.R));
    fp.R.n_comp = 1;
    fp.radix_ok = !probe && !keys.empty();
    for (int k = 0; k < P.n_accs && fp.radix_ok; ++k) {
      RComp c;
      memset(&c, 0, sizeof(c));
      c.is_f64 = full[k].kind == FK_SUMF;
      c.n_factors = full[k].n_factors;
      c.coef = full[k].coef;
      for (int f = 0; f < full[k].n_factors; ++f) c.f[f] = full[k].f[f];
      int found = -1;
      for (int j = 0; j + 1 < fp.R.n_comp; ++j)
        if (memcmp(&fp.R.comp[j], &c, sizeof(c)) == 0) found = j + 1;
      if (found < 0) {
        if (fp.R.n_comp >= R_MAXCOMP) {
          fp.radix_ok = false;
          break;
        }
        fp.R.comp[fp.R.n_comp - 1] = c;
        found = fp.R.n_comp++;
      }
      fp.R.comp_of[k] = found;
      fp.R.kind_of[k] = full[k].kind;
      fp.R.cmask[found - 1] |= 1 << full[k].kind;
      fp.R.acc_at[found - 1][full[k].kind] = k;
    }
    fp.R.n_accs = P.n_accs;
    for (int k = 1; k < P.n_accs; ++k) {
      const FAcc& prv = full[k - 1];
      FAcc& cur = P.accs[k];
      if (full[k].kind == FK_SUMF || prv.kind == FK_SUMF) continue;
      if (prv.n_factors == 0 || prv.n_factors > full[k].n_factors || prv.coef != full[k].coef) continue;
      bool prefix = true;
      for (int f = 0; f < prv.n_factors && prefix; ++f)
        prefix = prv.f[f].col == full[k].f[f].col && prv.f[f].a == full[k].f[f].a && prv.f[f].b == full[k].f[f].b;
      if (!prefix) continue;
      cur.chain = 1;
      cur.n_factors = full[k].n_factors - prv.n_factors;
      for (int f = 0; f < cur.n_factors; ++f) cur.f[f] = full[k].f[prv.n_factors + f];
    }
  }

  // ---- columns + shared-memory layout -------------------------------------------------------------------
  const int64_t n_rows = v.num_rows;
  P.n_rows = n_rows;
  P.n_tiles = (n_rows + F_T - 1) / F_T;
  P.n_cols = (int)A.slots.size();
  uint32_t stage_bytes = 0;
  for (int c = 0; c < P.n_cols; ++c) {
    const Slot& s = A.slots[c];
    FCol& fc = P.cols[c];
    if (s.dict) {
      fc.ptr = (const unsigned char*)s.col->dict_codes->ptr;
      fc.width = 1;
      fc.kind = 1;
    } else {
      fc.ptr = (const unsigned char*)s.col->data->ptr;
      fc.width = (uint32_t)phys_width(s.col->phys);
      fc.kind = (s.col->phys == PH_U8 || s.col->phys == PH_U16 || s.col->phys == PH_U32) ? 1 : 0;
    }
    if (((uintptr_t)fc.ptr & 15) != 0) return false;
    fc.smem_off = stage_bytes;
    stage_bytes += (uint32_t)(((size_t)F_T * fc.width + 127) & ~(size_t)127);
  }
  if (P.n_cols == 0) return false;  // nothing to stream (COUNT(*) without predicate): generic path
  auto resolve = [&](int col, uint32_t* off, uint32_t* wk) {
    *off = P.cols[col].smem_off;
    *wk = P.cols[col].width | (P.cols[col].kind ? 256u : 0u);
    if (P.cols[col].width == 8) *wk = 8;
  };
  for (int k = 0; k < P.n_pred; ++k) resolve(P.pred[k].col, &P.pred[k].off, &P.pred[k].wk);
  for (int k = 0; k < P.n_keys; ++k) resolve(P.keys[k].col, &P.keys[k].off, &P.keys[k].wk);
  for (int k = 0; k < P.n_accs; ++k) {
    FAcc& a = P.accs[k];
    for (int f = 0; f < a.n_factors; ++f) {
      resolve(a.f[f].col, &a.f[f].off, &a.f[f].wk);
      a.f[f].plain = (a.f[f].a == 0 && a.f[f].b == 1) ? 1 : 0;
    }
    a.unit = (a.kind != FK_SUMF && !a.chain && a.coef == 1 && a.n_factors > 0 && a.f[0].plain) ? 1 : 0;
  }
  for (int c = 0; c + 1 < fp.R.n_comp; ++c)
    for (int f = 0; f < fp.R.comp[c].n_factors; ++f) {
      resolve(fp.R.comp[c].f[f].col, &fp.R.comp[c].f[f].off, &fp.R.comp[c].f[f].wk);
      fp.R.comp[c].f[f].plain = (fp.R.comp[c].f[f].a == 0 && fp.R.comp[c].f[f].b == 1) ? 1 : 0;
    }
  P.stage_bytes = stage_bytes;
  const int NA2 = P.n_accs + 2;
  const int grid_max = ctx->sm_count;
  const int64_t tiles_per_cta = (P.n_tiles + grid_max - 1) / grid_max;
  // DENSE needs: small domain, private tables + >= 2 stages in shared memory, int64-safe per-thread partial sums
  size_t priv_bytes = 0;
  if (dense_ok) {
    priv_bytes = (size_t)dense_groups * NA2 * F_NT * 8;
    if (priv_bytes + 128 + 2 * (size_t)stage_bytes > (size_t)F_SMEM_MAX) dense_ok = false;
    for (i128 m : acc_maxabs)
      if (m * (i128)(tiles_per_cta * F_R + 1) >= LIM62) dense_ok = false;
  }
  P.mode = probe ? (emit ? FM_EMIT : FM_PROBE) : (dense_ok ? FM_DENSE : FM_HASH);
  if (P.mode == FM_PROBE) {
    priv_bytes = 0;
    for (i128 m : acc_maxabs)
      if (m * (i128)n_rows >= LIM62) P.carry = 1;
  }
  if (P.mode == FM_HASH) {
    priv_bytes = 0;
    // bit-packed code
    int shift = 0;
    for (int k = 0; k < P.n_keys; ++k) {
      const int bits = P.keys[k].pad;
      P.keys[k].mult = shift >= 64 ? 0 : (1ull << shift);
      fp.key_shift[k] = shift;
      if (fp.key_width[k] == 0) fp.radix_ok = false;  // dictionary-coded Utf8 key: FM_HASH only
      shift += bits;
    }
    for (i128 m : acc_maxabs)
      if (m * (i128)n_rows >= LIM62) P.carry = 1;
  }
  for (int k = 0; k < P.n_keys; ++k) P.keys[k].pad = 0;
  if (P.mode != FM_HASH || P.carry) fp.radix_ok = false;
  // accumulators that may ride in the count word: non-negative SUMs whose per-thread partial sums (<= rows_per_thread
  // rows) provably fit a bit field.  Their private lines disappear: fewer shared-memory read-modify-writes per row and
  // room for one more TMA stage.
  uint32_t safe_mask = 0;
  memset(P.pack_shift, 0, sizeof(P.pack_shift));
  memset(P.pack_bits, 0, sizeof(P.pack_bits));
  P.pack_mask = 0;
  P.pack_cnt_bits = 0;
  auto bits_of = [](i128 x) { int b = 1; while (b < 63 && ((i128)1 << b) <= x) ++b; return b; };
  const i128 rows_per_thread = (i128)tiles_per_cta * F_R + 1;
  int need_bits[F_MAXA];
  if (P.mode == FM_DENSE && !getenv("QGPU_FUSED_NOPACK")) {
    for (int k = 0; k < P.n_accs; ++k) {
      need_bits[k] = 64;
      if (P.accs[k].kind != FK_SUM || acc_lo[k] < 0) continue;
      const i128 bound = acc_maxabs[k] * rows_per_thread;
      if (bound >= ((i128)1 << 48)) continue;
      need_bits[k] = bits_of(bound);
      safe_mask |= 1u << k;
    }
  }
  uint32_t pack = 0;
  // QGPU_FUSED_GENERIC: the interpreted body; QGPU_FUSED_NOAOT: skip the ahead-of-time shapes (the run-time compiler then
  // specialises Q1 / Q6 too: experiments and bench.py's roofline_jit)
  fp.spec = (getenv("QGPU_FUSED_GENERIC") || getenv("QGPU_FUSED_NOAOT")) ? nullptr : find_specialised(P, safe_mask, &pack);
  if (fp.spec && pack) {
    int shift = bits_of(rows_per_thread);
    const int cnt_bits = shift;
    for (int k = 0; k < P.n_accs; ++k)
      if ((pack >> k) & 1u) {
        P.pack_shift[k] = (uint8_t)shift;
        P.pack_bits[k] = (uint8_t)need_bits[k];
        shift += need_bits[k];
      }
    if (shift <= 62) {
      P.pack_mask = pack;
      P.pack_cnt_bits = (uint32_t)cnt_bits;
    } else {  // the fields do not fit one word: the unpacked instantiation of the same shape
      memset(P.pack_shift, 0, sizeof(P.pack_shift));
      memset(P.pack_bits, 0, sizeof(P.pack_bits));
      fp.spec = find_specialised(P, 0, &pack);
    }
  }
  if (!fp.spec && P.mode == FM_DENSE && !getenv("QGPU_FUSED_GENERIC")) {
    // no registered kernel for this shape: instantiate SpecBody for its signature at run time.  Every accumulator that
    // may ride in the count word does, as long as the fields fit the word (fewer private lines, more TMA stages).
    uint32_t jpack = 0;
    int shift = bits_of(rows_per_thread);
    const int cnt_bits = shift;
    uint8_t jshift[F_MAXA] = {0}, jbits[F_MAXA] = {0};
    for (int k = 0; k < P.n_accs; ++k)
      if (((safe_mask >> k) & 1u) && shift + need_bits[k] <= 62) {
        jpack |= 1u << k;
        jshift[k] = (uint8_t)shift;
        jbits[k] = (uint8_t)need_bits[k];
        shift += need_bits[k];
      }
    const FSig g = make_sig(P);
    fp.jit = jit_specialised_dense(agg.ctx->device, g.s, jpack);
    if (fp.jit && jpack) {
      P.pack_mask = jpack;
      P.pack_cnt_bits = (uint32_t)cnt_bits;
      memcpy(P.pack_shift, jshift, sizeof(jshift));
      memcpy(P.pack_bits, jbits, sizeof(jbits));
    }
  }
  // private-table lines per slot: the unpacked accumulators, then the count word, then the first row
  P.priv_lines = 0;
  for (int k = 0; k < NA2; ++k) {
    if (k < P.n_accs && ((P.pack_mask >> k) & 1u)) {
      P.priv_line_of[k] = 0xff;
    } else {
      P.priv_k_of[P.priv_lines] = (uint8_t)k;
      P.priv_line_of[k] = (uint8_t)P.priv_lines++;
    }
  }
  if (P.mode == FM_DENSE) priv_bytes = (size_t)dense_groups * P.priv_lines * F_NT * 8;
  int ctas_per_sm = 1;
  int stages = (int)(((size_t)F_SMEM_MAX - 128 - priv_bytes) / stage_bytes);
  stages = std::min(stages, 4);
  if (const char* se = getenv("QGPU_FUSED_STAGES")) stages = std::min(stages, std::max(2, atoi(se)));  // experiments
  if (P.mode != FM_DENSE) {
    const int s2 = (int)std::min<size_t>(((size_t)F_SMEM_MAX / 2 - 1024 - 128) / stage_bytes, 4);
    if (s2 >= 2) {
      stages = s2;
      ctas_per_sm = 2;
    }
    // join probes: a third resident CTA (27 warps/SM; the kernels are held to 72 registers) hides more of the probes'
    // L2 latency than deeper staging does -- Q3 SF10: PROBE 0.387 -> 0.355 ms, EMIT 0.196 -> 0.161 ms
    const int want = getenv("QGPU_FUSED_CTAS") ? atoi(getenv("QGPU_FUSED_CTAS")) : (probe ? 3 : 2);
    if (want == 3) {
      const int sw = (int)std::min<size_t>(((size_t)F_SMEM_MAX / 3 - 1024 - 128) / stage_bytes, 4);
      int fit = 0;
      if (sw >= 2) {
        const size_t smem3 = 128 + (size_t)sw * stage_bytes;
        const void* fn = P.mode == FM_EMIT ? (const void*)k_fused_scan_agg<FM_EMIT>
                                           : (P.mode == FM_PROBE ? (const void*)k_fused_scan_agg<FM_PROBE> : (const void*)k_fused_scan_agg<FM_HASH>);
        CUDA_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F_SMEM_MAX));
        CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fit, fn, F_NT + 32, smem3));
      }
      if (fit >= 3) {
        stages = sw;
        ctas_per_sm = 3;
      }
    }
  }
  if (stages < 2) return false;
  P.stages = stages;
  P.priv_off = 128 + (uint32_t)stages * stage_bytes;
  P.dense_groups = (int)dense_groups;
  fp.smem_bytes = (size_t)P.priv_off + priv_bytes;
  fp.grid = (int)std::min<int64_t>(P.n_tiles, (int64_t)grid_max * ctas_per_sm);
  fp.total_bits = total_bits;
  return true;
}



// ------------------------------------------------------------------------------------------------
// RADIX mode driver (kernels: radix_agg.cuh).  run_radix returns false when the radix path declines (few groups:
// the L2-resident FM_HASH table is the better plan) or overflowed (estimate off / skewed bucket): FM_HASH then runs.
// The same stages, cut at the level-1 scatter, form the multi-GPU exchange (radix_exchange_* below).
// ------------------------------------------------------------------------------------------------
namespace {
// small device state block: hll[4096] u32 | hist1[256] u32 | tpre[257] u32 (+pad) | off1[257] u64 | cur1[256] u64 | n_out u64 | overflow
constexpr size_t RO_HLL = 0, RO_HIST1 = RO_HLL + R_HLL_M * 4, RO_TPRE = RO_HIST1 + R_P1 * 4, RO_OFF1 = RO_TPRE + (R_P1 + 4) * 4,
                 RO_CUR1 = RO_OFF1 + (R_P1 + 1) * 8, RO_NOUT = RO_CUR1 + R_P1 * 8, RO_OVF = RO_NOUT + 8, RO_BYTES = RO_OVF + 8;
constexpr size_t R_SKETCH_BYTES = RO_TPRE;  // hll + hist1: what the ranks exchange
}  // namespace

static DBufP radix_state_block(Ctx* ctx, RParams& R) {
  DBufP st = ctx->alloc_zero(RO_BYTES);
  char* sp = (char*)st->ptr;
  R.hll = (unsigned int*)(sp + RO_HLL);
  R.hist1 = (unsigned int*)(sp + RO_HIST1);
  R.tpre = (unsigned int*)(sp + RO_TPRE);
  R.off1 = (unsigned long long*)(sp + RO_OFF1);
  R.cur1 = (unsigned long long*)(sp + RO_CUR1);
  R.n_out = (unsigned long long*)(sp + RO_NOUT);
  R.overflow = (int*)(sp + RO_OVF);
  return st;
}

// level-2 fan-out, shared-memory table capacity and staging rows of the final pass for `groups` groups in n_tuples rows.
// final pass shared memory: 18 B per table slot (key, row count, first staged row, occupied list) + per staged row 8 B
// per operand value and 4 B (slot, rank); the table stays below ~45 % load, the staging area keeps ~15 % headroom
static bool radix_choose(const RParams& R, double groups, int64_t n_tuples, int l1_buckets, int* b2_out, int* cap_out, int* row_cap_out) {
  const size_t smem_budget = (size_t)F_SMEM_MAX - 256;
  const size_t row_bytes = 8 * (size_t)(R.n_comp - 1) + 4;
  int b2 = 0, cap = 0, row_cap = 0;
  for (b2 = 1; b2 <= R_MAXB2; ++b2) {
    const double nb = (double)((int64_t)std::max(l1_buckets, 1) << b2);  // populated final buckets
    const double g_b = groups / nb, r_b = (double)n_tuples / nb;
    cap = 64;
    while (cap < 4096 && g_b > 0.45 * cap) cap *= 2;
    if (g_b > 0.45 * cap) continue;
    const size_t table = (size_t)(cap + 1) * 18 + 16;
    if (table + 1024 >= smem_budget) continue;
    row_cap = (int)std::min<size_t>((smem_budget - table) / row_bytes, (size_t)1 << 18);
    if ((double)row_cap >= r_b * 1.15 + 6.0 * sqrt(r_b) + 64.0) break;
  }
  if (b2 > R_MAXB2) return false;
  if (const char* fb = getenv("QGPU_RADIX_B2")) {  // experiments: a larger level-2 fan-out than the sizes ask for
    const int want = std::min(R_MAXB2, atoi(fb));
    if (want > b2) b2 = want;
  }
  if (const char* tc = getenv("QGPU_RADIX_TEST_CAP")) {  // tests: provoke the overflow -> FM_HASH fallback
    cap = std::max(64, std::min(4096, atoi(tc)));
    b2 = 1;
    row_cap = (int)std::min<size_t>((smem_budget - (size_t)(cap + 1) * 18 - 16) / row_bytes, (size_t)1 << 18);
  }
  *b2_out = b2;
  *cap_out = cap;
  *row_cap_out = row_cap;
  return true;
}

static size_t radix_scatter_smem(const RParams& R) { return (size_t)R.n_comp * R_T * 8 + 512 * 8 + 512 * 4 * 2 + (size_t)R_T * 2; }

// input stages of the TMA-pipelined scatter (k_radix_scatter_tma): LEVEL 1 stages the raw columns of a 4096-row tile,
// LEVEL 2 the tuple components (+ 2 tuples: an 8 B aligned source is copied from the 16 B boundary below it).  false when
// two stages do not fit the shared memory (wide tuples: the register-staged kernel runs) or QGPU_RADIX_SCATTER=regs.
static size_t radix_tma_smem(const RStage& st) { return 128 + 2 * (size_t)st.stage_bytes + 2 * (size_t)R_T * 2 + 512 * 8 + 512 * 4 * 2; }
static bool radix_tma_stage(int level, const FParams& P, const RParams& R, RStage* st) {
  const char* e = getenv("QGPU_RADIX_SCATTER");
  if (e && strcmp(e, "regs") == 0) return false;
  memset(st, 0, sizeof(*st));
  uint32_t off = 0;
  if (level == 1) {
    st->n_in = (uint32_t)P.n_cols;
    for (int c = 0; c < P.n_cols; ++c) {
      st->col_of[c] = (uint32_t)c;
      st->off[c] = off;
      st->bytes_per_row[c] = P.cols[c].width;
      off += (uint32_t)(((size_t)R_T * P.cols[c].width + 127) & ~(size_t)127);
    }
  } else if (R.pair12) {  // codes | (value 1, value 2) pairs
    st->n_in = 2;
    st->off[0] = 0;
    st->bytes_per_row[0] = 8;
    off = (uint32_t)(((size_t)(R_T + 2) * 8 + 127) & ~(size_t)127);
    st->off[1] = off;
    st->bytes_per_row[1] = 16;
    off += (uint32_t)(((size_t)R_T * 16 + 127) & ~(size_t)127);
  } else {
    st->n_in = (uint32_t)R.n_comp;
    for (int c = 0; c < R.n_comp; ++c) {
      st->off[c] = off;
      st->bytes_per_row[c] = 8;
      off += (uint32_t)(((size_t)(R_T + 2) * 8 + 127) & ~(size_t)127);
    }
  }
  st->stage_bytes = off;
  st->n_stages = 2;
  return radix_tma_smem(*st) + 256 <= (size_t)F_SMEM_MAX;
}

// level-1 histogram + sketch: the TMA-pipelined kernel whenever two stages of the key / predicate columns fit
static void radix_launch_hist1(Ctx* ctx, const FParams& P, const RParams& R) {
  RStage st;
  memset(&st, 0, sizeof(st));
  bool used[F_MAXC] = {false};
  for (int k = 0; k < P.n_pred; ++k) used[P.pred[k].col] = true;
  for (int k = 0; k < P.n_keys; ++k) used[P.keys[k].col] = true;
  uint32_t off = 0;
  for (int c = 0; c < P.n_cols; ++c) {
    if (!used[c]) continue;
    st.col_of[st.n_in++] = (uint32_t)c;
    st.off[c] = off;
    st.bytes_per_row[c] = P.cols[c].width;
    off += (uint32_t)(((size_t)R_T * P.cols[c].width + 127) & ~(size_t)127);
  }
  st.stage_bytes = off;
  const size_t fixed = 128 + (size_t)R_P1 * 4 + (size_t)R_HLL_M * 4;
  const size_t budget = ((size_t)F_SMEM_MAX - 2048) / 2;  // two CTAs per SM
  const char* e = getenv("QGPU_RADIX_SCATTER");
  if (st.n_in > 0 && off > 0 && fixed + 2 * (size_t)off <= budget && !(e && strcmp(e, "regs") == 0)) {
    st.n_stages = (uint32_t)std::min<size_t>(4, (budget - fixed) / off);
    const size_t smem = fixed + (size_t)st.n_stages * off;
    const int64_t tiles = (P.n_rows + R_T - 1) / R_T;
    CUDA_CHECK(cudaFuncSetAttribute(k_radix_hist1_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LAUNCH(ctx, k_radix_hist1_tma, (int)std::max<int64_t>(1, std::min<int64_t>(tiles, (int64_t)ctx->sm_count * 2)), RH_NT + 32, smem, P, st, R.hist1,
           R.hll);
    return;
  }
  const int grid1 = (int)std::max<int64_t>(1, std::min<int64_t>(P.n_tiles, (int64_t)ctx->sm_count * 8));
  LAUNCH(ctx, k_radix_hist1, grid1, R_NT, 0, P, R.hist1, R.hll);
}

// two operand values and the TMA scatter on both levels: the values travel as 16-byte pairs (RParams::pair12)
static bool radix_pair12(const FParams& P, const RParams& R) {
  const char* e = getenv("QGPU_RADIX_PAIR");
  if ((e && e[0] == '0') || R.n_comp != 3) return false;
  RParams t = R;
  t.pair12 = 1;
  RStage st;
  return radix_tma_stage(1, P, t, &st) && radix_tma_stage(2, P, t, &st);
}

static void radix_launch_scatter1(Ctx* ctx, const FParams& P, const RParams& R) {
  const int64_t tiles1 = (P.n_rows + R_T - 1) / R_T;
  RStage st;
  if (radix_tma_stage(1, P, R, &st)) {
    const size_t smem = radix_tma_smem(st);
    CUDA_CHECK(cudaFuncSetAttribute(k_radix_scatter_tma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LAUNCH(ctx, k_radix_scatter_tma<1>, (int)std::max<int64_t>(1, std::min<int64_t>(tiles1, (int64_t)ctx->sm_count)), R2_NT, smem, P, R, st);
    return;
  }
  const size_t sc_smem = radix_scatter_smem(R);
  CUDA_CHECK(cudaFuncSetAttribute(k_radix_scatter<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sc_smem));
  const int sc_per_sm = (2 * (sc_smem + 1024 + 64) <= (size_t)233472) ? 2 : 1;
  LAUNCH(ctx, k_radix_scatter<1>, (int)std::max<int64_t>(1, std::min<int64_t>(tiles1, (int64_t)ctx->sm_count * sc_per_sm)), R_SNT, sc_smem, P, R);
}

// level-1 tuples (R.tup_a, bucket offsets R.off1 / tile list R.tpre on the device) -> level 2 -> final pass -> View.
// returns false on overflow (*fail = the kernel's code)
static bool radix_tail(PlanNode& agg, const View& v, FusedPlan& fp, const FParams& P, const int* key_shift, const int* key_bits, RParams& R,
                       const DBufP& st, int64_t n_tuples, unsigned int n_tiles2, double est, View* out, int* fail) {
  Ctx* ctx = agg.ctx;
  const int b2 = R.b2, cap = R.cap, row_cap = R.row_cap;
  const size_t row_bytes = 8 * (size_t)(R.n_comp - 1) + 4;
  const int64_t n_buckets = (int64_t)R_P1 << b2;
  const int64_t out_cap = std::max<int64_t>(1, std::min<int64_t>(n_tuples, (int64_t)(est * 1.15) + 65536));
  R.out_cap = out_cap;
  // ONE slab for all level-2 tuple arrays: a single large block of a stable size is what the stream-ordered pool re-uses
  // without fragmenting (many ~8 GB blocks interleaved with the result columns made later executions re-map memory)
  const size_t comp_bytes = (((size_t)n_tuples * 8 + 64) + 255) & ~(size_t)255;
  DBufP slab_b = ctx->alloc(comp_bytes * (size_t)R.n_comp);
  for (int c = 0; c < R.n_comp; ++c) R.tup_b[c] = (unsigned long long*)((char*)slab_b->ptr + comp_bytes * (size_t)c);
  DBufP hist2 = ctx->alloc_zero((size_t)n_buckets * 4), off2 = ctx->alloc((size_t)(n_buckets + 1) * 8), cur2 = ctx->alloc((size_t)n_buckets * 8);
  R.hist2 = (unsigned int*)hist2->ptr;
  R.off2 = (unsigned long long*)off2->ptr;
  R.cur2 = (unsigned long long*)cur2->ptr;
  // ---- result layout.  Key columns are decoded from the packed code by the final pass itself.  When every aggregate is
  //      a copy (SUM / MIN / MAX of an int64-like value, COUNT), a sign extension (Decimal128) or a Float64 mean of one
  //      accumulator, the final pass also writes the RESULT columns (no NULLs can arise: every group has rows and the
  //      arguments are NULL-free in this mode) and finish_aggregate's extra pass over 5 x 100 M values disappears.
  RKeys rk;
  memset(&rk, 0, sizeof(rk));
  rk.n_keys = P.n_keys;
  std::vector<DColP> key_cols;
  for (int k = 0; k < P.n_keys; ++k) {
    auto d = std::make_shared<DCol>();
    d->type = fp.key_src[k]->type;
    d->phys = fp.key_src[k]->phys;
    d->null_count = 0;
    d->data = ctx->alloc(std::max<size_t>((size_t)out_cap * fp.key_width[k], 16));
    rk.shift[k] = key_shift[k];
    rk.bits[k] = key_bits[k];
    rk.width[k] = fp.key_width[k];
    rk.base[k] = P.keys[k].base;
    rk.out[k] = d->data->ptr;
    key_cols.push_back(d);
  }
  std::vector<int> acc_kind(fp.specs.size(), AK_COUNT);
  for (size_t i = 0; i < fp.specs.size(); ++i) {
    const int k = fp.acc_of[i];
    if (k < 0) continue;
    const int fk = P.accs[k].kind;
    const VClass vc = class_of(fp.specs[i].arg->result_type);
    if (fk == FK_SUMF) acc_kind[i] = AK_SUM_F64;
    else if (fk == FK_SUM) acc_kind[i] = vc == VC_DEC ? AK_SUM_DEC : AK_SUM_I64;
    else if (vc == VC_DEC) acc_kind[i] = fk == FK_MIN ? AK_MIN_DEC : AK_MAX_DEC;
    else if (vc == VC_UINT) acc_kind[i] = fk == FK_MIN ? AK_MIN_U64 : AK_MAX_U64;
    else acc_kind[i] = fk == FK_MIN ? AK_MIN_I64 : AK_MAX_I64;
  }
  bool direct = !getenv("QGPU_RADIX_NODIRECT");
  std::vector<int> fin_mode(fp.specs.size(), 0);
  {
    int per_acc[F_MAXA] = {0}, n_cnt = 0;
    for (size_t i = 0; i < fp.specs.size() && direct; ++i) {
      const AggSpec& a = fp.specs[i];
      DType produced = a.op == QGPU_AGG_COUNT ? mk_type(QGPU_T_INT64) : a.return_type;
      if (produced != agg.schema.fields[fp.keys.size() + i].type) direct = false;  // finish_aggregate raises the schema error
      const Phys op = out_phys_of(produced);
      const int ak = acc_kind[i];
      int mode = 0;
      if (a.op == QGPU_AGG_COUNT) mode = op == PH_I64 ? 1 : 0;
      else if (a.op == QGPU_AGG_AVG) mode = (ak == AK_SUM_F64 && op == PH_F64) ? 3 : 0;
      else if (op == PH_I128) mode = (ak == AK_SUM_DEC || ak == AK_MIN_DEC || ak == AK_MAX_DEC) ? 2 : 0;
      else if (op == PH_I64 || op == PH_U64)
        mode = (ak == AK_SUM_I64 || ak == AK_MIN_I64 || ak == AK_MAX_I64 || ak == AK_MIN_U64 || ak == AK_MAX_U64) ? 1 : 0;
      else if (op == PH_F64) mode = (ak == AK_SUM_F64 && a.op == QGPU_AGG_SUM) ? 1 : 0;
      if (!mode) direct = false;
      else if (fp.acc_of[i] < 0 ? ++n_cnt > 4 : ++per_acc[fp.acc_of[i]] > 2) direct = false;
      fin_mode[i] = mode;
    }
  }
  R.direct = direct ? 1 : 0;
  R.n_cnt_dst = 0;
  memset(R.fin_mode, 0, sizeof(R.fin_mode));
  std::vector<DColP> agg_cols;
  DBufP out_cnt;
  std::vector<DBufP> out_acc;
  if (direct) {
    int per_acc[F_MAXA] = {0};
    for (size_t i = 0; i < fp.specs.size(); ++i) {
      const AggSpec& a = fp.specs[i];
      auto col = std::make_shared<DCol>();
      col->type = a.op == QGPU_AGG_COUNT ? mk_type(QGPU_T_INT64) : a.return_type;
      col->phys = out_phys_of(col->type);
      col->null_count = 0;
      col->data = ctx->alloc((size_t)out_cap * phys_width(col->phys) + 64);
      const int k = fp.acc_of[i];
      if (k < 0) {
        R.cnt_dst[R.n_cnt_dst++] = (unsigned long long*)col->data->ptr;
      } else {
        R.fin_mode[k][per_acc[k]] = fin_mode[i];
        R.fin_dst[k][per_acc[k]++] = col->data->ptr;
      }
      agg_cols.push_back(col);
    }
  } else {
    out_cnt = ctx->alloc((size_t)out_cap * 8 + 64);
    R.out_cnt = (unsigned long long*)out_cnt->ptr;
    for (int k = 0; k < P.n_accs; ++k) {
      out_acc.push_back(ctx->alloc((size_t)out_cap * 8 + 64));
      R.out_acc[k] = (unsigned long long*)out_acc.back()->ptr;
    }
  }
  const size_t sc_smem = radix_scatter_smem(R);
  CUDA_CHECK(cudaFuncSetAttribute(k_radix_scatter<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sc_smem));
  const int sc_per_sm = (2 * (sc_smem + 1024 + 64) <= (size_t)233472) ? 2 : 1;
  RParams R2 = R;
  R2.world = 1;  // level 2 and the final pass are local
  LAUNCH(ctx, k_radix_hist2, (int)std::min<int64_t>(std::max<int64_t>(n_tiles2, 1), (int64_t)ctx->sm_count * 8), R_NT, 0, R2);
  LAUNCH(ctx, k_radix_scan2, R_P1, 1 << R_MAXB2, 0, R2.hist2, b2, R2.off1, R2.off2, R2.cur2);
  RStage st2;
  if (radix_tma_stage(2, P, R2, &st2)) {
    const size_t smem = radix_tma_smem(st2);
    CUDA_CHECK(cudaFuncSetAttribute(k_radix_scatter_tma<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LAUNCH(ctx, k_radix_scatter_tma<2>, (int)std::min<int64_t>(std::max<int64_t>(n_tiles2, 1), (int64_t)ctx->sm_count), R2_NT, smem, P, R2, st2);
  } else {
    LAUNCH(ctx, k_radix_scatter<2>, (int)std::min<int64_t>(std::max<int64_t>(n_tiles2, 1), (int64_t)ctx->sm_count * sc_per_sm), R_SNT, sc_smem,
           P, R2);
  }
  const int ag_grid = (int)std::min<int64_t>(n_buckets, (int64_t)ctx->sm_count);
  const size_t ag_smem = (size_t)(cap + 1) * 18 + (size_t)row_cap * row_bytes + 64;
  const char* agf = getenv("QGPU_RADIX_AGG");  // "sort": the sorting form of the final pass
  const bool list_form = !(agf && strcmp(agf, "sort") == 0);
#define QGPU_RADIX_AGG(NV)                                                                                             \
  case NV:                                                                                                             \
    if (list_form) {                                                                                                   \
      CUDA_CHECK(cudaFuncSetAttribute(k_radix_agg_list<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ag_smem)); \
      LAUNCH(ctx, k_radix_agg_list<NV>, ag_grid, R_AGG_NT, ag_smem, R2, rk);                                           \
    } else {                                                                                                           \
      CUDA_CHECK(cudaFuncSetAttribute(k_radix_agg<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ag_smem));    \
      LAUNCH(ctx, k_radix_agg<NV>, ag_grid, R_AGG_NT, ag_smem, R2, rk);                                                \
    }                                                                                                                  \
    break;
  switch (R.n_comp - 1) {  // operand values per tuple: unrolled in the kernel
    QGPU_RADIX_AGG(0)
    QGPU_RADIX_AGG(1)
    QGPU_RADIX_AGG(2)
    QGPU_RADIX_AGG(3)
    QGPU_RADIX_AGG(4)
    QGPU_RADIX_AGG(5)
    default: throw_internal("radix aggregate: too many operand values");
  }
#undef QGPU_RADIX_AGG
  unsigned long long fin[2];
  ctx->d2h_sync(fin, (char*)st->ptr + RO_NOUT, 16);
  ctx->trace("radix: level 2 + aggregate");
  const int64_t n_groups = (int64_t)fin[0];
  *fail = (int)fin[1];
  if (*fail != 0) return false;
  slab_b.reset();
  for (auto& kc : key_cols) {
    kc->length = n_groups;
    if (kc->phys == PH_D64) kc = materialize_arrow(ctx, {kc, nullptr}, n_groups);
  }
  GroupAccs accs;
  if (!direct) {
    accs.n_groups = n_groups;
    accs.unordered = true;
    for (size_t i = 0; i < fp.specs.size(); ++i) {
      const int k = fp.acc_of[i];
      accs.lo.push_back(k >= 0 ? out_acc[k] : out_cnt);
      accs.hi.push_back(nullptr);  // no carry in this mode: the high word is the sign extension
      accs.kind.push_back(acc_kind[i]);
      accs.cnt.push_back(out_cnt);
    }
  }
  agg.strategy = "fused_scan_agg[radix-partitioned: " + std::to_string(R_P1) + " x " + std::to_string(1 << b2) + " buckets, smem table " + std::to_string(cap) +
                 " slots / " + std::to_string(row_cap) + " rows, " + std::to_string(R.n_comp) + " x 8 B tuples, " +
                 std::to_string(P.n_cols) + " cols, " + std::to_string(P.n_pred) + " range preds, " + std::to_string(P.n_accs) +
                 " accs, est " + std::to_string((int64_t)est) + " groups, scatter " + (radix_tma_stage(1, P, R, &st2) ? "tma" : "regs") + "/" +
                 (radix_tma_stage(2, P, R2, &st2) ? "tma" : "regs") + (R.pair12 ? ", 16 B value pairs" : "") +
                 (direct ? ", result columns written by the final pass]" : "]");
  if (direct) {
    View o;
    o.schema = agg.schema;
    o.num_rows = n_groups;
    o.num_batches = 1;
    for (size_t i = 0; i < fp.keys.size(); ++i) {
      if (fp.keys[i]->result_type != agg.schema.fields[i].type)
        throw_arrow("column types must match schema types, expected " + agg.schema.fields[i].type.str() + " but found " +
                    fp.keys[i]->result_type.str() + " at column index " + std::to_string(i));
      o.cols.push_back({key_cols[i], nullptr});
    }
    for (auto& c : agg_cols) {
      c->length = n_groups;
      o.cols.push_back({c, nullptr});
    }
    *out = o;
    return true;
  }
  *out = finish_aggregate(ctx, v, fp.keys, fp.specs, agg.schema, accs, &key_cols, nullptr);
  ctx->trace("radix: finish_aggregate");
  return true;
}

static bool run_radix(PlanNode& agg, const View& v, FusedPlan& fp, View* out) {
  Ctx* ctx = agg.ctx;
  const FParams& P = fp.P;
  const char* mode_env = getenv("QGPU_RADIX");  // "off": never, "force": whenever the shape allows (tests)
  const bool force = mode_env && strcmp(mode_env, "force") == 0;
  if (mode_env && strcmp(mode_env, "off") == 0) return false;
  if (!force && P.n_rows < ((int64_t)1 << 24)) return false;
  RParams R = fp.R;
  R.world = 1;
  R.nopf = getenv("QGPU_RADIX_NOPF") ? atoi(getenv("QGPU_RADIX_NOPF")) : 0;
  // ---- pass 0: level-1 histogram + HyperLogLog sketch ---------------------------------------------------------
  DBufP st = radix_state_block(ctx, R);
  radix_launch_hist1(ctx, P, R);
  LAUNCH(ctx, k_radix_scan1, 1, R_P1, 0, R.hist1, R.off1, R.cur1, R.tpre);
  std::vector<unsigned char> hst(RO_CUR1);
  ctx->d2h_sync(hst.data(), st->ptr, RO_CUR1);
  const double est = hll_estimate((const unsigned int*)(hst.data() + RO_HLL));
  const int64_t n_tuples = (int64_t)((const unsigned long long*)(hst.data() + RO_OFF1))[R_P1];
  const unsigned int n_tiles2 = ((const unsigned int*)(hst.data() + RO_TPRE))[R_P1];
  ctx->trace("radix: hist1 + sketch");
  if (n_tuples == 0) return false;
  const double groups = std::min(est * 1.05 + 64.0, (double)n_tuples);
  if (!force && groups < 4.0e6) {
    // the hash table stays (mostly) L2-resident: FM_HASH, sized from the estimate so that it does not grow
    int64_t c = 1 << 12;
    while ((double)c < 2.2 * groups) c <<= 1;
    fp.learned_cap = std::max(fp.learned_cap, c);
    fp.radix_failed = true;  // same table, same answer: do not sketch again
    return false;
  }
  if (!radix_choose(R, groups, n_tuples, R_P1, &R.b2, &R.cap, &R.row_cap)) return false;
  R.pair12 = radix_pair12(P, R) ? 1 : 0;  // array 1 then spans the slab's second and third part
  const size_t comp_bytes = (((size_t)n_tuples * 8 + 64) + 255) & ~(size_t)255;
  DBufP slab_a = ctx->alloc(comp_bytes * (size_t)R.n_comp);  // one slab (see radix_tail)
  for (int c = 0; c < R.n_comp; ++c) R.tup_a[c] = (unsigned long long*)((char*)slab_a->ptr + comp_bytes * (size_t)c);
  radix_launch_scatter1(ctx, P, R);
  int fail = 0;
  if (!radix_tail(agg, v, fp, P, fp.key_shift, fp.key_bits, R, st, n_tuples, n_tiles2, est, out, &fail)) {
    fp.radix_failed = true;  // FM_HASH from now on (the results of this attempt are discarded)
    return false;
  }
  return true;
}

// ------------------------------------------------------------------------------------------------
// DENSE mode, launch-lean: scan kernel + ONE epilogue kernel (epilogue.cu), no host round trip.  The global accumulator
// table lives as long as the cached analysis: it is initialised once and re-initialised by every epilogue; the result
// columns of one execution come out of a single stream-ordered allocation and their metadata (group count, NULL
// counts, error code) is still in flight when this returns (View::pending).  With `sharded` the same two kernels are
// the whole multi-GPU step: the epilogue exchanges the state blocks over peer memory and merges them.
// ------------------------------------------------------------------------------------------------
namespace {
struct DenseRun {
  bool eligible = false;
  // TWO accumulator tables used alternately: the epilogue of execution i (epilogue stream) still reads and re-initialises
  // table i & 1 while the scan kernel of execution i + 1 (compute stream) already fills the other one
  bool inited[2] = {false, false};
  DBufP g_lo[2], g_hi[2];
  cudaEvent_t epi_done[2] = {nullptr, nullptr};  // recorded on the epilogue stream behind the epilogue that used table b
  uint64_t runs = 0;
  ~DenseRun() {
    for (cudaEvent_t e : epi_done)
      if (e) cudaEventDestroy(e);
  }
  std::vector<DBufP> keep;  // dictionary strings on the device
  std::vector<DType> key_types;
  std::vector<int> kinds;   // AccKind per aggregate
  EpiParams E;
};
}  // namespace

static int acc_kind_of(int fk, const DType& arg_type) {
  const VClass vc = class_of(arg_type);
  if (fk == FK_SUMF) return AK_SUM_F64;
  if (fk == FK_SUM) return vc == VC_DEC ? AK_SUM_DEC : AK_SUM_I64;
  if (vc == VC_DEC) return fk == FK_MIN ? AK_MIN_DEC : AK_MAX_DEC;
  if (vc == VC_UINT) return fk == FK_MIN ? AK_MIN_U64 : AK_MAX_U64;
  return fk == FK_MIN ? AK_MIN_I64 : AK_MAX_I64;
}

static DenseRun* prepare_dense(PlanNode& agg, FusedPlan& fp) {
  if (fp.dense) {
    DenseRun* d = (DenseRun*)fp.dense.get();
    return d->eligible ? d : nullptr;
  }
  auto d = std::make_shared<DenseRun>();
  fp.dense = d;
  Ctx* ctx = agg.ctx;
  SpecSuspend cached_work(ctx);
  const FParams& P = fp.P;
  const int nk = (int)fp.keys.size(), na = (int)fp.specs.size();
  if (P.mode != FM_DENSE || getenv("QGPU_NO_EPILOGUE")) return nullptr;
  if (nk > EPI_MAXK || na > EPI_MAXAGG || P.n_accs > EPI_MAXACC || P.dense_groups > EPI_MAXG || P.dense_groups < 1) return nullptr;
  EpiParams& E = d->E;
  memset(&E, 0, sizeof(E));
  E.src = EPI_SRC_DENSE;
  E.n_slots = P.dense_groups;
  E.n_accs = P.n_accs;
  for (int k = 0; k < P.n_accs + 2; ++k) {
    if (k < P.n_accs) E.init[k] = P.accs[k].kind == FK_MIN ? INT64_MAX : (P.accs[k].kind == FK_MAX ? INT64_MIN : 0);
    else E.init[k] = k == P.n_accs ? 0 : INT64_MAX;
  }
  for (int k = 0; k < P.n_accs; ++k) {
    E.acc_kind[k] = P.accs[k].kind;
    E.acc_wide[k] = P.accs[k].kind == FK_SUM ? 1 : 0;
  }
  for (int k = 0; k < nk; ++k) {
    d->key_types.push_back(fp.keys[k]->result_type);
    EpiKey& K = E.key[k];
    K.is_dict = fp.key_width[k] == 0 ? 1 : 0;
    K.base = P.keys[k].base;
    K.mult = (unsigned int)P.keys[k].mult;
    K.range = (unsigned int)((i128)fp.key_max[k] - (i128)fp.key_min[k] + 1);
    if (K.is_dict) {
      const DCol& src = *fp.key_src[k];
      std::vector<int32_t> offs(1, 0);
      std::string bytes;
      for (const std::string& sv : src.dict_values) {
        if (sv.size() > 16) return nullptr;  // state records carry Utf8 keys of <= 16 bytes: the multi-kernel path runs
        bytes += sv;
        offs.push_back((int32_t)bytes.size());
      }
      DBufP o = ctx->alloc(offs.size() * 4), b = ctx->alloc(std::max<size_t>(bytes.size(), 16));
      ctx->h2d(o->ptr, offs.data(), offs.size() * 4);
      if (!bytes.empty()) ctx->h2d(b->ptr, bytes.data(), bytes.size());
      ctx->sync();
      K.dict_offs = (const int32_t*)o->ptr;
      K.dict_data = (const char*)b->ptr;
      d->keep.push_back(o);
      d->keep.push_back(b);
    }
  }
  for (int i = 0; i < na; ++i) {
    const int k = fp.acc_of[i];
    E.agg_acc[i] = k;
    d->kinds.push_back(k < 0 ? (int)AK_COUNT : acc_kind_of(P.accs[k].kind, fp.specs[i].arg->result_type));
  }
  d->eligible = true;
  return d.get();
}

static View run_dense(PlanNode& agg, FusedPlan& fp, DenseRun& D, bool sharded, int64_t row_offset, int max_groups) {
  Ctx* ctx = agg.ctx;
  FParams P = fp.P;
  EpiParams E = D.E;
  epilogue_describe(ctx, D.key_types, fp.specs, D.kinds, agg.schema, E);  // type checks (RecordBatch::try_new) before any launch
  const int NA2 = P.n_accs + 2;
  const int b = (int)(D.runs++ & 1);
  if (!D.g_lo[b]) {
    D.g_lo[b] = ctx->alloc((size_t)P.dense_groups * NA2 * 8);
    D.g_hi[b] = ctx->alloc((size_t)P.dense_groups * NA2 * 8);
    D.g_lo[b]->free_stream = D.g_hi[b]->free_stream = ctx->epi_stream;  // the last epilogue re-initialises them
    CUDA_CHECK(cudaEventCreateWithFlags(&D.epi_done[b], cudaEventDisableTiming));
  } else {
    CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, D.epi_done[b], 0));  // the epilogue two executions ago has reset this table
  }
  if (!D.inited[b]) {
    FInit init;
    for (int k = 0; k < NA2; ++k) init.v[k] = E.init[k];
    LAUNCH(ctx, k_fused_init, grid_for(ctx, (int64_t)P.dense_groups * NA2, 256), 256, 0, (unsigned long long*)D.g_lo[b]->ptr,
           (unsigned long long*)D.g_hi[b]->ptr, (int64_t)P.dense_groups, NA2, 1, 0, init);
  }
  D.inited[b] = false;  // until the epilogue that re-initialises the table has been queued
  P.g_lo = (unsigned long long*)D.g_lo[b]->ptr;
  P.g_hi = (unsigned long long*)D.g_hi[b]->ptr;
  P.n_groups = nullptr;   // DENSE never touches them
  P.abort_flag = nullptr;
  E.g_lo = P.g_lo;
  E.g_hi = P.g_hi;
  const bool specialised = fp.spec != nullptr || fp.jit;
  if (fp.spec) {
    CUDA_CHECK(cudaFuncSetAttribute(fp.spec, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fp.smem_bytes));
    ::qgpu::KernelScope ks(ctx, "k_fused_scan_agg_spec", fp.grid);
    fp.spec<<<fp.grid, F_NT + 32, fp.smem_bytes, ctx->stream>>>(P);
    CUDA_CHECK(cudaGetLastError());
  } else if (fp.jit) {
    ::qgpu::KernelScope ks(ctx, "k_fused_scan_agg_jit", fp.grid);
    jit_launch(fp.jit, fp.grid, F_NT + 32, fp.smem_bytes, ctx->stream, &P);
  } else {
    CUDA_CHECK(cudaFuncSetAttribute(k_fused_scan_agg<FM_DENSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fp.smem_bytes));
    LAUNCH(ctx, k_fused_scan_agg<FM_DENSE>, fp.grid, F_NT + 32, fp.smem_bytes, P);
  }
  View out = epilogue_execute(ctx, E, D.key_types, fp.specs, agg.schema, sharded, row_offset, sharded ? max_groups : P.dense_groups, nullptr);
  CUDA_CHECK(cudaEventRecord(D.epi_done[b], ctx->epi_stream));
  D.inited[b] = true;
  agg.strategy = std::string("fused_scan_agg[dense-private") + (fp.spec ? "/shape-specialised, " : (fp.jit ? "/shape-specialised at run time, " : ", ")) +
                 (P.pack_mask ? std::to_string(__builtin_popcount(P.pack_mask)) + " accs packed into the count word, " : "") +
                 std::to_string(P.n_cols) + " cols, " + std::to_string(P.n_pred) + " range preds, " + std::to_string(P.n_accs) +
                 " accs, " + std::to_string(P.stages) + " TMA stages] + single-CTA epilogue" +
                 (sharded ? "[peer exchange + merge over " + std::to_string(ctx->comm ? ctx->comm->world : 1) + " ranks]" : "");
  return out;
}

static View run_fused(PlanNode& agg, const View& v, FusedPlan& fp) {
  Ctx* ctx = agg.ctx;
  ctx->trace(nullptr);
  FParams P = fp.P;
  std::vector<std::shared_ptr<Compiled>>& keys = fp.keys;
  std::vector<AggSpec>& specs = fp.specs;
  std::vector<int>& acc_of = fp.acc_of;
  const int total_bits = fp.total_bits;
  const size_t smem_bytes = fp.smem_bytes;
  const int grid = fp.grid;
  const int64_t n_rows = P.n_rows;
  const int NA2 = P.n_accs + 2;
  // ---- global accumulator table ---------------------------------------------------------------------------
  FInit init;
  for (int k = 0; k < NA2; ++k) {
    if (k < P.n_accs) init.v[k] = P.accs[k].kind == FK_MIN ? INT64_MAX : (P.accs[k].kind == FK_MAX ? INT64_MIN : 0);
    else init.v[k] = k == P.n_accs ? 0 : INT64_MAX;
  }
  CUDA_CHECK(cudaFuncSetAttribute(k_fused_scan_agg<FM_DENSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  CUDA_CHECK(cudaFuncSetAttribute(k_fused_scan_agg<FM_HASH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  DBufP g_lo, g_hi, flags;
  std::unique_ptr<Slab> slab;  // DENSE only
  int64_t rec = 0;        // HASH: words per group record (0: accumulators start at word 0)
  bool specialised = false;
  int64_t n_slots = 0, k_stride = 0, g_stride = 0;
  int64_t cap = 0;
  if (P.mode == FM_DENSE && !agg.defer) {
    if (DenseRun* D = prepare_dense(agg, fp)) return run_dense(agg, fp, *D, false, 0, 0);
  }
  if (P.mode == FM_DENSE) {
    n_slots = P.dense_groups;
    k_stride = 1;
    g_stride = NA2;
    g_lo = ctx->alloc((size_t)n_slots * NA2 * 8);
    g_hi = ctx->alloc((size_t)n_slots * NA2 * 8);
    // every zero-initialised buffer of this operator (flags, group count, the export's per-aggregate arrays) comes out
    // of ONE allocation and ONE memset: a re-executed Q1 step otherwise queues ~25 cudaMallocAsync + cudaMemsetAsync
    // pairs around the scan kernel
    slab.reset(new Slab(ctx, (5 + 2 * specs.size()) * Slab::need(std::max<size_t>((size_t)n_slots * 8, 16)), true));
    flags = slab->take(16);
    LAUNCH(ctx, k_fused_init, grid_for(ctx, n_slots * NA2, 256), 256, 0, (unsigned long long*)g_lo->ptr,
           (unsigned long long*)g_hi->ptr, n_slots, NA2, 1, 0, init);
    P.g_lo = (unsigned long long*)g_lo->ptr;
    P.g_hi = (unsigned long long*)g_hi->ptr;
    P.abort_flag = (int*)((char*)flags->ptr + 8);
    P.n_groups = (unsigned long long*)flags->ptr;
    FusedKernel spec = fp.spec;
    if (spec) {
      specialised = true;
      CUDA_CHECK(cudaFuncSetAttribute(spec, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
      ::qgpu::KernelScope ks(ctx, "k_fused_scan_agg_spec", grid);
      spec<<<grid, F_NT + 32, smem_bytes, ctx->stream>>>(P);
      CUDA_CHECK(cudaGetLastError());
    } else if (fp.jit) {
      specialised = true;
      ::qgpu::KernelScope ks(ctx, "k_fused_scan_agg_jit", grid);
      jit_launch(fp.jit, grid, F_NT + 32, smem_bytes, ctx->stream, &P);
    } else {
      LAUNCH(ctx, k_fused_scan_agg<FM_DENSE>, grid, F_NT + 32, smem_bytes, P);
    }
  } else {
    if (fp.radix_ok && !fp.radix_failed && !agg.defer) {
      View rv;
      if (run_radix(agg, v, fp, &rv)) return rv;
    }
    // capacity: bounded by the key domain and by the row count; grown x8 on overflow
    i128 domain = total_bits >= 63 ? ((i128)1 << 63) : ((i128)1 << total_bits);
    int64_t max_groups = (int64_t)std::min<i128>(domain, (i128)n_rows);
    int64_t need = 2;
    while (need < 2 * max_groups) need <<= 1;
    cap = std::min<int64_t>(need, std::max<int64_t>((int64_t)1 << 22, fp.learned_cap));
    while (true) {
      n_slots = cap + 1;
      rec = ((NA2 + 1 + 7) / 8) * 8;  // record words: key + accumulators + count + first row, padded to 64 B
      k_stride = 1;
      g_stride = rec;
      g_lo = ctx->alloc((size_t)n_slots * rec * 8);
      g_hi = P.carry ? ctx->alloc((size_t)n_slots * rec * 8) : nullptr;
      flags = ctx->alloc_zero(16);
      LAUNCH(ctx, k_fused_init, grid_for(ctx, n_slots * rec, 256), 256, 0, (unsigned long long*)g_lo->ptr,
             g_hi ? (unsigned long long*)g_hi->ptr : nullptr, n_slots, NA2, 2, (int)rec, init);
      P.g_lo = (unsigned long long*)g_lo->ptr;
      P.g_hi = g_hi ? (unsigned long long*)g_hi->ptr : nullptr;
      P.cap_mask = (uint64_t)(cap - 1);
      P.acc_base = 1;
      P.acc_kstride = 1;
      P.acc_gstride = rec;
      P.n_groups = (unsigned long long*)flags->ptr;
      P.abort_flag = (int*)((char*)flags->ptr + 8);
      LAUNCH(ctx, k_fused_scan_agg<FM_HASH>, grid, F_NT + 32, smem_bytes, P);
      const int aborted = ctx->read_scalar((const int*)((char*)flags->ptr + 8));
      if (!aborted) {
        fp.learned_cap = cap;
        break;
      }
      if (cap >= need) throw_internal("fused aggregate: hash table overflow (internal error)");
      cap = std::min<int64_t>(cap * 8, need);
    }
  }

  ctx->trace("scan-agg: init + kernel(s)");
  // ---- export: occupied slots -> dense group ids, accumulators -> GroupAccs ---------------------------------
  const bool grouped = !keys.empty();
  const bool small = P.mode == FM_DENSE;  // <= 4096 slots: one-CTA export, no host round trip for the group count
  int64_t n_groups = 0;
  DBufP occ, offs, n_groups_dev;
  if (small) {
    n_groups = grouped ? n_slots : 1;
    n_groups_dev = slab->take(8);
  } else {
    occ = ctx->alloc((size_t)n_slots * 8);
    offs = ctx->alloc((size_t)n_slots * 8);
    LAUNCH(ctx, k_fused_occupied, grid_for(ctx, n_slots, 256), 256, 0, (const unsigned long long*)g_lo->ptr + (rec ? 1 : 0), n_slots,
           (int64_t)P.n_accs * k_stride, g_stride, (int64_t*)occ->ptr);
    n_groups = exclusive_scan_i64(ctx, (const int64_t*)occ->ptr, (int64_t*)offs->ptr, n_slots);
  }
  GroupAccs accs;
  accs.n_groups = grouped ? n_groups : 1;
  if (small && grouped) accs.n_groups_dev = n_groups_dev;
  const int64_t ng_alloc = std::max<int64_t>(accs.n_groups, 1);
  auto zalloc = [&](size_t bytes) { return slab ? slab->take(bytes) : ctx->alloc_zero(bytes); };
  DBufP cnt = zalloc((size_t)ng_alloc * 8);
  DBufP first = zalloc((size_t)ng_alloc * 8);
  DBufP zero = zalloc((size_t)ng_alloc * 8);
  FExport ex;
  memset(&ex, 0, sizeof(ex));
  ex.n_aggs = (int)specs.size();
  for (size_t i = 0; i < specs.size(); ++i) {
    AggSpec& s = specs[i];
    const int k = acc_of[i];
    ex.acc_of[i] = k;
    int ak = AK_COUNT;
    if (k >= 0) {
      const int fk = P.accs[k].kind;
      ex.kind_of[i] = fk;
      ex.wide[i] = (fk == FK_SUM && (P.mode == FM_DENSE || P.carry)) ? 1 : 0;
      const VClass vc = class_of(s.arg->result_type);
      if (fk == FK_MIN || fk == FK_MAX) minmax_sentinel(s.arg->result_type, fk == FK_MIN, &ex.sent_lo[i], &ex.sent_hi[i]);
      if (fk == FK_SUMF) ak = AK_SUM_F64;
      else if (fk == FK_SUM) ak = vc == VC_DEC ? AK_SUM_DEC : AK_SUM_I64;
      else if (vc == VC_DEC) ak = fk == FK_MIN ? AK_MIN_DEC : AK_MAX_DEC;
      else if (vc == VC_UINT) ak = fk == FK_MIN ? AK_MIN_U64 : AK_MAX_U64;
      else ak = fk == FK_MIN ? AK_MIN_I64 : AK_MAX_I64;
      DBufP lo = zalloc((size_t)ng_alloc * 8), hi = zalloc((size_t)ng_alloc * 8);
      ex.out_lo[i] = (unsigned long long*)lo->ptr;
      ex.out_hi[i] = (unsigned long long*)hi->ptr;
      accs.lo.push_back(lo);
      accs.hi.push_back(hi);
    } else {
      accs.lo.push_back(zero);
      accs.hi.push_back(zero);
    }
    accs.kind.push_back(ak);
    accs.cnt.push_back(cnt);
  }
  accs.first_row = first;
  if (small) {
    // ungrouped: slot 0 is exported even when no row passed -- MIN/MAX then keep their sentinels (reference
    // quirk Q4), SUM/AVG are NULL because the row count is 0
    LAUNCH(ctx, k_fused_export_small, 1, 1024, 0, (const unsigned long long*)g_lo->ptr,
           g_hi ? (const unsigned long long*)g_hi->ptr : nullptr, (int)n_slots, k_stride, g_stride, P.n_accs, grouped ? 0 : 1, ex,
           (unsigned long long*)cnt->ptr, (long long*)first->ptr, (long long*)n_groups_dev->ptr);
  } else {
    if (n_groups > 0)
      LAUNCH(ctx, k_fused_export, grid_for(ctx, n_slots, 256), 256, 0, (const unsigned long long*)g_lo->ptr + (rec ? 1 : 0),
             g_hi ? (const unsigned long long*)g_hi->ptr + (rec ? 1 : 0) : nullptr, n_slots, k_stride, g_stride, P.n_accs,
             (const int64_t*)occ->ptr, (const int64_t*)offs->ptr, ex, (unsigned long long*)cnt->ptr, (long long*)first->ptr);
    if (!grouped && n_groups == 0) {
      // no row passed the filter: MIN/MAX keep their sentinels (reference quirk Q4), SUM/AVG are NULL (cnt == 0)
      for (size_t i = 0; i < specs.size(); ++i) {
        const int k = acc_of[i];
        if (k < 0 || P.accs[k].kind == FK_SUM || P.accs[k].kind == FK_SUMF) continue;
        unsigned long long h[2] = {ex.sent_lo[i], ex.sent_hi[i]};
        ctx->h2d(accs.lo[i]->ptr, &h[0], 8);
        ctx->h2d(accs.hi[i]->ptr, &h[1], 8);
        ctx->sync();
      }
    }
  }
  agg.strategy = std::string("fused_scan_agg[") + (P.mode == FM_DENSE ? "dense-private" : "hbm-hash") +
                 (fp.spec ? "/shape-specialised, " : (specialised ? "/shape-specialised at run time, " : ", ")) +
                 (P.pack_mask ? std::to_string(__builtin_popcount(P.pack_mask)) + " accs packed into the count word, " : "") +
                 std::to_string(P.n_cols) + " cols, " + std::to_string(P.n_pred) + " range preds, " + std::to_string(P.n_accs) +
                 " accs, " + std::to_string(P.stages) + " TMA stages]";
  ctx->trace("scan-agg: export");
  if (agg.defer) {
    agg.defer->set = true;
    agg.defer->input = v;
    agg.defer->keys = keys;
    agg.defer->specs = specs;
    agg.defer->accs = accs;
    return View();
  }
  View fin = finish_aggregate(ctx, v, keys, specs, agg.schema, accs);
  ctx->trace("scan-agg: finish_aggregate");
  return fin;
}

// ------------------------------------------------------------------------------------------------
// Fused hash-join probe + aggregate (TPC-H Q3's J2 + HashAggregate):
//   HashAggregate <- HashJoinExec(Inner, one integer key, no JoinFilter) <- [build: any sub-plan, probe: (Filter)* <- Scan]
//   HashJoinExec::{build_hash_table, probe_hash_table}   qurious/src/physical/plan/join/hash_join.rs:148-216,354-385
//   HashAggregate                                        qurious/src/physical/plan/aggregate/hash.rs:138-170
// The build side runs through the ordinary operators (it is small); its join key is inserted into an
// HBM-resident open-addressing table (slot = hash tag | build row, equality verified on the key column).
// When the build keys are UNIQUE and every group key is the join key or a build-side column, the group of a
// probe row is fully determined by its matching build row: the probe scan then streams through the same
// TMA-staged pipeline as the scan-aggregate kernel and accumulates straight into per-build-row accumulators
// (FM_PROBE) -- no join output, no second hash table.  Anything else returns false (generic operators run).
// ------------------------------------------------------------------------------------------------

// membership bitmap over the build key's value range (from the base column's statistics: the gathered subset lies
// inside it); skipped when the range is unknown or would need more than 256 MB
static DBufP join_key_bitmap(Ctx* ctx, const LazyCol& key, int64_t nb, int64_t* kmin, uint64_t* kspan) {
  *kmin = 0;
  *kspan = 0;
  if (nb <= 0 || !key.base || getenv("QGPU_NO_JOIN_BITMAP")) return nullptr;
  {
    SpecSuspend cached_work(ctx);
    ensure_stats(ctx, *key.base);
  }
  if (!key.base->has_stats) return nullptr;
  const i128 range = key.base->vmax - key.base->vmin + 1;
  if (range <= 0 || range > ((i128)1 << 31) || key.base->vmin < -(LIM62 * 2) || key.base->vmax > LIM62 * 2 - 1) return nullptr;
  *kmin = (int64_t)key.base->vmin;
  *kspan = (uint64_t)(range - 1);
  return ctx->alloc_zero((size_t)((range + 31) / 32) * 4 + 16);
}

__global__ void __launch_bounds__(256) k_join_build_unique(const void* __restrict__ bkey, int width, const uint32_t* __restrict__ validity,
                                                           int64_t n, unsigned long long* __restrict__ slots, uint64_t mask,
                                                           int* __restrict__ dup_flag, uint32_t* __restrict__ bitmap, int64_t kmin) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < n; row += stride) {
    if (validity && !((validity[row >> 5] >> (row & 31)) & 1u)) continue;  // NULL keys never match (hash_join.rs:177-216)
    const int64_t key = width == 8 ? ((const long long*)bkey)[row] : (int64_t)((const int*)bkey)[row];
    const uint64_t h = fmix64((uint64_t)key);
    const uint32_t tag = (uint32_t)(h >> 32);
    const unsigned long long mine = ((unsigned long long)tag << 32) | (unsigned long long)(row + 1);
    if (bitmap) {
      const uint64_t k = (uint64_t)key - (uint64_t)kmin;
      atomicOr(&bitmap[k >> 5], 1u << (k & 31));
    }
    uint64_t sl = h & mask;
    while (true) {
      unsigned long long cur = *(volatile unsigned long long*)&slots[sl];
      if (cur == 0) {
        cur = atomicCAS(&slots[sl], 0ull, mine);
        if (cur == 0) break;
      }
      if ((uint32_t)(cur >> 32) == tag) {
        const int64_t other = (int64_t)(cur & 0xffffffffull) - 1;
        const int64_t ok = width == 8 ? ((const long long*)bkey)[other] : (int64_t)((const int*)bkey)[other];
        if (ok == key) {  // duplicate build key: the fused path requires uniqueness
          *dup_flag = 1;
          break;
        }
      }
      sl = (sl + 1) & mask;
    }
  }
}

__global__ void k_occupied_rows(const int64_t* __restrict__ flags, const int64_t* __restrict__ offs, int64_t n, int64_t* __restrict__ rows) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    if (flags[i]) rows[offs[i]] = i;
}

namespace {
std::unique_ptr<ExprNode> clone_shift(const ExprNode& n, int shift, bool* ok) {
  auto c = std::make_unique<ExprNode>();
  c->kind = n.kind;
  c->col_index = n.col_index;
  c->lit_type = n.lit_type;
  c->lit_null = n.lit_null;
  c->lit_lo = n.lit_lo;
  c->lit_hi = n.lit_hi;
  c->lit_str = n.lit_str;
  c->op = n.op;
  c->cast_type = n.cast_type;
  c->n_when = n.n_when;
  if (n.kind == QGPU_IR_COLUMN) {
    if (n.col_index < shift) *ok = false;  // references the build side
    c->col_index = n.col_index - shift;
  }
  for (auto& ch : n.children) c->children.push_back(clone_shift(*ch, shift, ok));
  return c;
}
bool int_like_key(const DCol& c) { return c.phys == PH_I64 || c.phys == PH_I32 || c.phys == PH_D64; }
}  // namespace

// ------------------------------------------------------------------------------------------------
// The build side of a fused probe: HBM-resident linear-probing table over the build key (+ membership bitmap).
//   (i)  build side = (Filter)* <- Scan with range / dictionary predicates: ONE scan kernel (FM_BUILD) evaluates the
//        predicate and inserts the qualifying rows -- no selection vector, no gathered key column; the table holds
//        base-table rows;
//   (ii) anything else: the build side's operators run, its key column is gathered, k_join_build_unique inserts it.
// false: the fused probes cannot use this build side (non-integer key, duplicate keys, >= 2^32 rows).
// ------------------------------------------------------------------------------------------------
namespace {
struct JoinTable {
  View bv;          // build-side columns, addressed by the row ids stored in the table
  int64_t nb = 0;   // build rows (bound of the row ids)
  DColP bkey;       // key column indexed by build row (null when nb == 0)
  DBufP slots, bitmap, dup;
  int64_t cap = 0, kmin = 0;
  uint64_t kspan = 0;
  std::string how;
  void bind(FParams& P) const {
    P.jt_slots = (const unsigned long long*)slots->ptr;
    P.jt_mask = (uint64_t)(cap - 1);
    P.jt_bitmap = bitmap ? (const uint32_t*)bitmap->ptr : nullptr;
    P.jt_kmin = kmin;
    P.jt_kspan = kspan;
    P.bkey = nb > 0 ? bkey->data->ptr : slots->ptr;
    P.bkey_width = nb > 0 ? phys_width(bkey->phys) : 8;
  }
};
}  // namespace

static bool analyze_fused(PlanNode& agg, const std::vector<const ExprNode*>& predicates, const View& v, FusedPlan& fp,
                          const ProbeOpts* probe);
bool fused_unordered_join(PlanNode& join, View* out);

static bool fused_filtered_build(PlanNode& child, int key_col, JoinTable& jt) {
  Ctx* ctx = child.ctx;
  if (getenv("QGPU_NO_FUSED_BUILD")) return false;
  std::vector<const ExprNode*> predicates;
  PlanNode* n = &child;
  while (n->kind == PK_FILTER) {
    predicates.push_back(n->predicate.get());
    n = n->children[0].get();
  }
  if (n->kind != PK_SCAN) return false;
  if (n->predicate) predicates.push_back(n->predicate.get());
  if (predicates.empty()) return false;  // nothing to fuse: (ii) builds straight from the resident key column
  View pv = scan_view(*n);
  if (pv.num_batches == 0 || pv.num_rows == 0 || pv.num_rows >= 0xfffffff0LL) return false;
  if (key_col < 0 || key_col >= (int)pv.cols.size() || !pv.cols[key_col].base) return false;
  const DCol& kc = *pv.cols[key_col].base;
  if (!(kc.phys == PH_I64 || kc.phys == PH_I32 || kc.phys == PH_D64) || kc.null_count != 0) return false;
  std::shared_ptr<FusedPlan> fp = std::static_pointer_cast<FusedPlan>(child.fused_cache);
  bool fresh = false;
  if (fp) {
    fresh = fp->n_rows == pv.num_rows && fp->n_batches == pv.num_batches && fp->col_ids.size() == pv.cols.size();
    for (size_t i = 0; fresh && i < pv.cols.size(); ++i) fresh = fp->col_ids[i] == pv.cols[i].base.get() && !pv.cols[i].idx;
  }
  if (!fresh) {
    fp = std::make_shared<FusedPlan>();
    ProbeOpts po;
    po.emit = true;  // no aggregates: the kernel-side "key" is the build key
    po.probe_key_col = key_col;
    PlanNode shell;
    shell.ctx = ctx;
    try {
      fp->usable = analyze_fused(shell, predicates, pv, *fp, &po);
    } catch (QError&) {
      fp->usable = false;  // the generic operators raise the error in their own order
    }
    fp->n_rows = pv.num_rows;
    fp->n_batches = pv.num_batches;
    for (auto& c : pv.cols) fp->col_ids.push_back(c.base.get());
    child.fused_cache = fp;
  }
  if (!fp->usable) return false;
  jt.bv = pv;
  jt.nb = pv.num_rows;
  jt.bkey = pv.cols[key_col].base;
  jt.cap = 1024;
  while (jt.cap < 2 * jt.nb) jt.cap <<= 1;
  jt.slots = ctx->alloc_zero((size_t)jt.cap * 8);
  jt.dup = ctx->alloc_zero(8);
  jt.bitmap = join_key_bitmap(ctx, pv.cols[key_col], jt.nb, &jt.kmin, &jt.kspan);
  FParams P = fp->P;
  jt.bind(P);
  P.jt_wslots = (unsigned long long*)jt.slots->ptr;
  P.jt_wbitmap = jt.bitmap ? (uint32_t*)jt.bitmap->ptr : nullptr;
  P.abort_flag = (int*)jt.dup->ptr;
  P.n_groups = (unsigned long long*)((char*)jt.dup->ptr);
  CUDA_CHECK(cudaFuncSetAttribute(k_fused_scan_agg<FM_BUILD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fp->smem_bytes));
  LAUNCH(ctx, k_fused_scan_agg<FM_BUILD>, fp->grid, F_NT + 32, fp->smem_bytes, P);
  jt.how = "fused_filtered_build[" + std::to_string(P.n_cols) + " cols, " + std::to_string(P.n_pred) + " range preds]";
  child.strategy = jt.how;
  return true;
}

static bool build_join_table(PlanNode& join, int lk_col, JoinTable& jt) {
  Ctx* ctx = join.ctx;
  PlanNode& child = *join.children[0];
  if (!fused_filtered_build(child, lk_col, jt)) {
    if (!fused_unordered_join(child, &jt.bv)) jt.bv = join.child_view(0);
    jt.nb = jt.bv.num_rows;
    if (jt.nb >= 0xfffffff0LL) return false;
    jt.bkey = jt.nb > 0 ? materialize(ctx, jt.bv.cols[lk_col], jt.nb) : nullptr;
    if (jt.nb > 0 && !int_like_key(*jt.bkey)) return false;
    jt.cap = 1024;
    while (jt.cap < 2 * jt.nb) jt.cap <<= 1;
    jt.slots = ctx->alloc_zero((size_t)jt.cap * 8);
    jt.dup = ctx->alloc_zero(8);
    jt.bitmap = join_key_bitmap(ctx, jt.bv.cols[lk_col], jt.nb, &jt.kmin, &jt.kspan);
    if (jt.nb > 0)
      LAUNCH(ctx, k_join_build_unique, grid_for(ctx, jt.nb, 256), 256, 0, jt.bkey->data->ptr, phys_width(jt.bkey->phys),
             jt.bkey->validity ? (const uint32_t*)jt.bkey->validity->ptr : nullptr, jt.nb, (unsigned long long*)jt.slots->ptr,
             (uint64_t)(jt.cap - 1), (int*)jt.dup->ptr, jt.bitmap ? (uint32_t*)jt.bitmap->ptr : nullptr, jt.kmin);
    jt.how = child.strategy;
  }
  if (jt.nb > 0 && ctx->read_count((const int*)jt.dup->ptr, "join_build.dup")) return false;  // duplicate build keys: the generic join runs
  return true;
}

// Order-free Inner join used INSIDE a fused pipeline whose consumer does not depend on row order (the build
// side of try_fused_join_aggregate): HashJoin(Inner, one integer key, no JoinFilter, unique build keys) with a
// (Filter)* <- Scan probe side.  The probe scan streams through the TMA pipeline (FM_EMIT) and appends
// (build row, probe row) pairs with warp-aggregated atomics; the result is a View of index vectors (late
// materialisation), like the generic join's -- only the row ORDER differs, which is why this path is never used
// for a join whose output is returned to the caller (hash_join.rs:474-512 pins that order).
bool fused_unordered_join(PlanNode& join, View* out) {
  Ctx* ctx = join.ctx;
  if (join.kind != PK_HASH_JOIN || join.join_type != QGPU_JOIN_INNER || join.left_on.size() != 1 || join.has_join_filter) return false;
  const ExprNode& lk = *join.left_on[0];
  const ExprNode& rk = *join.right_on[0];
  if (lk.kind != QGPU_IR_COLUMN || rk.kind != QGPU_IR_COLUMN) return false;
  std::vector<const ExprNode*> predicates;
  PlanNode* n = join.children[1].get();
  while (n->kind == PK_FILTER) {
    predicates.push_back(n->predicate.get());
    n = n->children[0].get();
  }
  if (n->kind != PK_SCAN) return false;
  if (n->predicate) predicates.push_back(n->predicate.get());
  View pv = scan_view(*n);
  if (pv.num_batches == 0 || pv.num_rows == 0 || pv.num_rows >= ((int64_t)1 << 40)) return false;
  const int n_left = (int)join.children[0]->schema.fields.size();
  if (rk.col_index < 0 || rk.col_index >= (int)pv.cols.size() || lk.col_index < 0 || lk.col_index >= n_left) return false;
  if (join.children[0]->schema.fields[lk.col_index].type != pv.schema.fields[rk.col_index].type) return false;
  ProbeOpts po;
  po.emit = true;
  po.probe_key_col = rk.col_index;
  std::shared_ptr<FusedPlan> fp = std::static_pointer_cast<FusedPlan>(join.fused_cache);
  bool fresh = false;
  if (fp) {
    fresh = fp->n_rows == pv.num_rows && fp->n_batches == pv.num_batches && fp->col_ids.size() == pv.cols.size();
    for (size_t i = 0; fresh && i < pv.cols.size(); ++i) fresh = fp->col_ids[i] == pv.cols[i].base.get() && !pv.cols[i].idx;
  }
  if (!fresh) {
    fp = std::make_shared<FusedPlan>();
    PlanNode shell;  // analyze_fused only needs the context in emit mode
    shell.ctx = ctx;
    try {
      fp->usable = analyze_fused(shell, predicates, pv, *fp, &po);
    } catch (QError&) {
      fp->usable = false;
    }
    fp->n_rows = pv.num_rows;
    fp->n_batches = pv.num_batches;
    for (auto& c : pv.cols) fp->col_ids.push_back(c.base.get());
    join.fused_cache = fp;
  }
  if (!fp->usable) return false;
  // build side (a filtered scan builds the table in one kernel; recursively order-free when it is itself such a join)
  JoinTable jt;
  if (!build_join_table(join, lk.col_index, jt)) return false;
  View& bv = jt.bv;
  const int64_t nb = jt.nb;
  ctx->trace("  emit-join: build side + table");
  FParams P = fp->P;
  auto b_idx = std::make_shared<IdxVec>();
  auto p_idx = std::make_shared<IdxVec>();
  b_idx->buf = ctx->alloc((size_t)pv.num_rows * 8);
  p_idx->buf = ctx->alloc((size_t)pv.num_rows * 8);
  DBufP flags = ctx->alloc_zero(16);
  P.g_lo = (unsigned long long*)b_idx->buf->ptr;
  P.g_hi = (unsigned long long*)p_idx->buf->ptr;
  P.n_groups = (unsigned long long*)flags->ptr;
  P.abort_flag = (int*)((char*)flags->ptr + 8);
  jt.bind(P);
  CUDA_CHECK(cudaFuncSetAttribute(k_fused_scan_agg<FM_EMIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fp->smem_bytes));
  LAUNCH(ctx, k_fused_scan_agg<FM_EMIT>, fp->grid, F_NT + 32, fp->smem_bytes, P);
  const int64_t n_out = (int64_t)ctx->read_count((const unsigned long long*)flags->ptr, "join_emit.pairs");
  ctx->trace("  emit-join: probe kernel");
  b_idx->length = p_idx->length = n_out;
  View v;
  v.schema = join.schema;
  v.num_rows = n_out;
  v.num_batches = n_out > 0 ? 1 : 0;
  std::vector<std::pair<IdxP, IdxP>> cb, cp;
  for (const LazyCol& c : bv.cols) v.cols.push_back(apply_selection(ctx, c, b_idx, &cb));
  for (const LazyCol& c : pv.cols) v.cols.push_back(apply_selection(ctx, c, p_idx, &cp));
  ctx->trace("  emit-join: output view");
  if (v.cols.size() != join.schema.fields.size()) return false;
  join.strategy = "fused_join_probe_emit[unordered, unique-build " + std::to_string(nb) + " rows -> " + std::to_string(n_out) + " pairs]";
  *out = v;
  return true;
}

static bool join_aggregate_body(PlanNode& agg, View* out) {
  Ctx* ctx = agg.ctx;
  if (agg.defer || agg.group_exprs.empty() || agg.aggs.empty() || (int)agg.aggs.size() > 24) return false;
  PlanNode* join = agg.children[0].get();
  if (join->kind != PK_HASH_JOIN || join->join_type != QGPU_JOIN_INNER || join->left_on.size() != 1 || join->has_join_filter) return false;
  const ExprNode& lk = *join->left_on[0];
  const ExprNode& rk = *join->right_on[0];
  if (lk.kind != QGPU_IR_COLUMN || rk.kind != QGPU_IR_COLUMN) return false;
  // probe side: (Filter)* <- Scan
  std::vector<const ExprNode*> predicates;
  PlanNode* n = join->children[1].get();
  while (n->kind == PK_FILTER) {
    predicates.push_back(n->predicate.get());
    n = n->children[0].get();
  }
  if (n->kind != PK_SCAN) return false;
  if (n->predicate) predicates.push_back(n->predicate.get());
  View pv = scan_view(*n);
  if (pv.num_batches == 0 || pv.num_rows == 0 || pv.num_rows >= ((int64_t)1 << 40)) return false;
  const int n_left = (int)join->children[0]->schema.fields.size();
  const int n_right = (int)pv.cols.size();
  if (rk.col_index < 0 || rk.col_index >= n_right || lk.col_index < 0 || lk.col_index >= n_left) return false;
  // group keys: the join key (either side) or a build-side column
  std::vector<int> key_build_col;  // build column that carries each group key's value
  for (auto& g : agg.group_exprs) {
    if (g->kind != QGPU_IR_COLUMN) return false;
    if (g->col_index < n_left) key_build_col.push_back(g->col_index);
    else if (g->col_index - n_left == rk.col_index) key_build_col.push_back(lk.col_index);
    else return false;
  }
  ProbeOpts po;
  po.probe_key_col = rk.col_index;
  bool ok = true;
  for (auto& a : agg.aggs) po.agg_exprs.push_back(clone_shift(*a.expr, n_left, &ok));
  if (!ok) return false;

  // ---- probe-side analysis (cached like the scan-aggregate one) -------------------------------------------
  std::shared_ptr<FusedPlan> fp = std::static_pointer_cast<FusedPlan>(agg.fused_cache);
  bool fresh = false;
  if (fp) {
    fresh = fp->n_rows == pv.num_rows && fp->n_batches == pv.num_batches && fp->col_ids.size() == pv.cols.size();
    for (size_t i = 0; fresh && i < pv.cols.size(); ++i) fresh = fp->col_ids[i] == pv.cols[i].base.get() && !pv.cols[i].idx;
  }
  if (!fresh) {
    fp = std::make_shared<FusedPlan>();
    try {
      fp->usable = analyze_fused(agg, predicates, pv, *fp, &po);
    } catch (QError&) {
      fp->usable = false;  // let the generic operators raise the error in their own order
    }
    fp->n_rows = pv.num_rows;
    fp->n_batches = pv.num_batches;
    for (auto& c : pv.cols) fp->col_ids.push_back(c.base.get());
    agg.fused_cache = fp;
  }
  if (!fp->usable) return false;

  // ---- build side: ordinary operators, then the join table ------------------------------------------------
  ctx->trace(nullptr);
  // the probe key must compare like the build key: same logical type (the generic join raises otherwise)
  if (join->children[0]->schema.fields[lk.col_index].type != pv.schema.fields[rk.col_index].type) return false;
  JoinTable jt;
  if (!build_join_table(*join, lk.col_index, jt)) return false;
  View& bv = jt.bv;
  const int64_t nb = jt.nb;

  ctx->trace("join-agg: build side + join table");
  // ---- probe + aggregate ---------------------------------------------------------------------------------------
  FParams P = fp->P;
  const int NA2 = P.n_accs + 2;
  const int64_t n_slots = std::max<int64_t>(nb, 1);
  FInit init;
  for (int k = 0; k < NA2; ++k) {
    if (k < P.n_accs) init.v[k] = P.accs[k].kind == FK_MIN ? INT64_MAX : (P.accs[k].kind == FK_MAX ? INT64_MIN : 0);
    else init.v[k] = k == P.n_accs ? 0 : INT64_MAX;
  }
  DBufP g_lo = ctx->alloc((size_t)n_slots * NA2 * 8);
  DBufP g_hi = P.carry ? ctx->alloc((size_t)n_slots * NA2 * 8) : nullptr;
  DBufP flags = ctx->alloc_zero(16);
  LAUNCH(ctx, k_fused_init, grid_for(ctx, n_slots * NA2, 256), 256, 0, (unsigned long long*)g_lo->ptr,
         g_hi ? (unsigned long long*)g_hi->ptr : nullptr, n_slots, NA2, 0, 0, init);
  P.g_lo = (unsigned long long*)g_lo->ptr;
  P.g_hi = g_hi ? (unsigned long long*)g_hi->ptr : nullptr;
  P.n_groups = (unsigned long long*)flags->ptr;
  P.abort_flag = (int*)((char*)flags->ptr + 8);
  jt.bind(P);
  P.acc_base = 0;
  P.acc_kstride = n_slots;
  P.acc_gstride = 1;
  CUDA_CHECK(cudaFuncSetAttribute(k_fused_scan_agg<FM_PROBE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fp->smem_bytes));
  LAUNCH(ctx, k_fused_scan_agg<FM_PROBE>, fp->grid, F_NT + 32, fp->smem_bytes, P);

  ctx->trace("join-agg: probe+aggregate kernel");
  // ---- groups = build rows that were matched at least once -----------------------------------------------------
  DBufP occ = ctx->alloc((size_t)n_slots * 8), offs = ctx->alloc((size_t)n_slots * 8);
  LAUNCH(ctx, k_fused_occupied, grid_for(ctx, n_slots, 256), 256, 0, (const unsigned long long*)g_lo->ptr, n_slots,
         (int64_t)P.n_accs * n_slots, (int64_t)1, (int64_t*)occ->ptr);
  const int64_t n_groups = exclusive_scan_i64(ctx, (const int64_t*)occ->ptr, (int64_t*)offs->ptr, n_slots);
  std::vector<AggSpec>& specs = fp->specs;
  std::vector<int>& acc_of = fp->acc_of;
  GroupAccs accs;
  accs.n_groups = n_groups;
  const int64_t ng_alloc = std::max<int64_t>(n_groups, 1);
  // one allocation + one memset for the export's zero-initialised arrays (see Slab)
  std::unique_ptr<Slab> slab;
  if ((size_t)ng_alloc * 8 * (3 + 2 * specs.size()) <= ((size_t)256 << 20))
    slab.reset(new Slab(ctx, (3 + 2 * specs.size()) * Slab::need((size_t)ng_alloc * 8), true));
  auto zalloc = [&](size_t bytes) { return slab ? slab->take(bytes) : ctx->alloc_zero(bytes); };
  DBufP cnt = zalloc((size_t)ng_alloc * 8), first = zalloc((size_t)ng_alloc * 8), zero = zalloc((size_t)ng_alloc * 8);
  FExport ex;
  memset(&ex, 0, sizeof(ex));
  ex.n_aggs = (int)specs.size();
  for (size_t i = 0; i < specs.size(); ++i) {
    const int k = acc_of[i];
    ex.acc_of[i] = k;
    int ak = AK_COUNT;
    if (k >= 0) {
      const int fk = P.accs[k].kind;
      ex.kind_of[i] = fk;
      ex.wide[i] = (fk == FK_SUM && P.carry) ? 1 : 0;
      const VClass vc = class_of(specs[i].arg->result_type);
      if (fk == FK_SUMF) ak = AK_SUM_F64;
      else if (fk == FK_SUM) ak = vc == VC_DEC ? AK_SUM_DEC : AK_SUM_I64;
      else if (vc == VC_DEC) ak = fk == FK_MIN ? AK_MIN_DEC : AK_MAX_DEC;
      else if (vc == VC_UINT) ak = fk == FK_MIN ? AK_MIN_U64 : AK_MAX_U64;
      else ak = fk == FK_MIN ? AK_MIN_I64 : AK_MAX_I64;
      DBufP lo = zalloc((size_t)ng_alloc * 8), hi = zalloc((size_t)ng_alloc * 8);
      ex.out_lo[i] = (unsigned long long*)lo->ptr;
      ex.out_hi[i] = (unsigned long long*)hi->ptr;
      accs.lo.push_back(lo);
      accs.hi.push_back(hi);
    } else {
      accs.lo.push_back(zero);
      accs.hi.push_back(zero);
    }
    accs.kind.push_back(ak);
    accs.cnt.push_back(cnt);
  }
  accs.first_row = first;
  // many groups: keep the build-row order of the groups (deterministic) instead of ranking them by first probe row -- the
  // reference's group order is unspecified (hash.rs:98) and the ranking is a 64-bit radix sort of n_groups pairs
  accs.unordered = n_groups > 4096;
  auto rows = std::make_shared<IdxVec>();
  rows->length = n_groups;
  rows->buf = ctx->alloc((size_t)ng_alloc * 8);
  if (n_groups > 0) {
    LAUNCH(ctx, k_fused_export, grid_for(ctx, n_slots, 256), 256, 0, (const unsigned long long*)g_lo->ptr,
           g_hi ? (const unsigned long long*)g_hi->ptr : nullptr, n_slots, n_slots, (int64_t)1, P.n_accs, (const int64_t*)occ->ptr,
           (const int64_t*)offs->ptr, ex, (unsigned long long*)cnt->ptr, (long long*)first->ptr);
    LAUNCH(ctx, k_occupied_rows, grid_for(ctx, n_slots, 256), 256, 0, (const int64_t*)occ->ptr, (const int64_t*)offs->ptr, n_slots,
           (int64_t*)rows->buf->ptr);
  }
  ctx->trace("join-agg: export");
  // key values travel with the group (gid-indexed): the build columns at the matched build rows
  std::vector<std::shared_ptr<Compiled>> keys;
  Schema js = join->schema;
  for (auto& g : agg.group_exprs) keys.push_back(compile_expr(*g, js));
  std::vector<DColP> key_cols;
  std::vector<std::pair<IdxP, IdxP>> cache;
  for (size_t i = 0; i < keys.size(); ++i) {
    check_hash_key_type(keys[i]->result_type);
    LazyCol lc = apply_selection(ctx, bv.cols[key_build_col[i]], rows, &cache);
    DColP kc = materialize(ctx, lc, n_groups);
    if (kc->phys == PH_D64) kc = materialize_arrow(ctx, {kc, nullptr}, n_groups);
    key_cols.push_back(kc);
  }
  agg.strategy = "fused_join_probe_agg[unique-build " + std::to_string(nb) + " rows, " + std::to_string(P.n_cols) + " probe cols, " +
                 std::to_string(P.n_pred) + " range preds, " + std::to_string(P.n_accs) + " accs, " + std::to_string(P.stages) +
                 " TMA stages] <- build[" + jt.how + "]";
  join->strategy = "fused-into-aggregate <- [" + jt.how + ", probe scan]";
  ctx->trace("join-agg: key columns");
  View dummy;
  dummy.schema = js;
  // asynchronous execution of a replayed pipeline: every count is already on the host (speculated), so the finalise
  // flags may stay in flight too -- the whole step then never waits for the device
  const bool defer_flags = ctx->async_ok && ctx->spec && ctx->spec->replay;
  *out = finish_aggregate(ctx, dummy, keys, specs, agg.schema, accs, &key_cols, nullptr, defer_flags);
  ctx->trace("join-agg: finish_aggregate");
  return true;
}

// identity of every table scanned below `n` (resident column buffers, row / batch counts): learned counts are only
// replayed over exactly the tables they were learned on
static uint64_t subtree_signature(PlanNode& n) {
  uint64_t h = 0x9e3779b97f4a7c15ULL * (uint64_t)n.kind;
  auto mix = [&](uint64_t x) { h = (h ^ x) * 0xff51afd7ed558ccdULL; h ^= h >> 29; };
  if (n.kind == PK_BROADCAST) {  // what EVERY rank contributed (the node runs now, once per call: exchange.cu)
    mix(broadcast_signature(n));
    return h;
  }
  if (n.kind == PK_SCAN && n.table) {
    n.table->consolidate();
    mix((uint64_t)(uintptr_t)n.table.get());
    mix((uint64_t)n.table->num_rows);
    mix((uint64_t)n.table->num_batches);
    for (auto& c : n.table->cols) mix((uint64_t)(uintptr_t)(c && c->data ? c->data->ptr : nullptr));
  }
  for (auto& c : n.children) mix(subtree_signature(*c));
  return h;
}

uint64_t local_subtree_signature(PlanNode& n) { return subtree_signature(n); }

View run_speculated(PlanNode& owner, uint64_t signature, const std::function<View()>& body) {
  Ctx* ctx = owner.ctx;
  if (ctx->spec) return body();  // already inside an enclosing scope
  if (!owner.spec) owner.spec = std::make_shared<Speculation>();
  Speculation& sp = *owner.spec;
  if (signature != owner.spec_sig) sp.have = false;
  owner.spec_sig = signature;
  for (int attempt = 0; attempt < 2; ++attempt) {
    SpecScope scope(ctx, &sp);
    const bool replay = sp.replay;
    try {
      View v = body();
      if (!replay) {
        sp.have = !sp.learned.empty();
        return v;
      }
      if (v.pending ? scope.defer_verify(v.pending) : scope.verify()) return v;
    } catch (SpeculationMiss&) {
      if (!replay) throw_internal("speculation miss outside a replay (internal error)");
    } catch (QError&) {
      if (!replay) throw;
    }
    sp.have = false;
  }
  throw_internal("speculated execution did not settle");
}

// Speculative re-execution: the first execution of the join pipeline learns its device-side counts (selection sizes,
// join output size, duplicate flags, group count) with one host round trip each; later executions over the same
// tables take them from the cache, never wait, and verify all of them after the pipeline's single final
// synchronisation.  A wrong guess (cannot happen over immutable tables; detected anyway) re-runs in learning mode.
bool try_fused_join_aggregate(PlanNode& agg, View* out) {
  Ctx* ctx = agg.ctx;
  if (ctx->spec) return join_aggregate_body(agg, out);  // already inside an enclosing scope
  if (!agg.spec) agg.spec = std::make_shared<Speculation>();
  Speculation& sp = *agg.spec;
  const uint64_t sig = subtree_signature(agg);
  if (sig != agg.spec_sig) sp.have = false;
  agg.spec_sig = sig;
  for (int attempt = 0; attempt < 2; ++attempt) {
    SpecScope scope(ctx, &sp);
    const bool replay = sp.replay;
    bool ok = false;
    try {
      View v;
      ok = join_aggregate_body(agg, &v);
      if (!replay) {
        sp.have = ok && !sp.learned.empty();
        if (ok) *out = v;
        return ok;
      }
      if (ok && v.pending ? scope.defer_verify(v.pending) : scope.verify()) {
        if (ok) *out = v;
        return ok;
      }
    } catch (SpeculationMiss&) {
      if (!replay) throw_internal("speculation miss outside a replay (internal error)");
    } catch (QError&) {
      if (!replay) throw;  // a replay's error may be an artefact of a wrong guess: learn again, which raises the real one
    }
    sp.have = false;
  }
  return false;
}

// analysis (cached on the node) of Aggregate <- (Filter <-)* Scan; null when the subtree has another shape
static std::shared_ptr<FusedPlan> fused_plan_for(PlanNode& agg, View* vout) {
  std::vector<const ExprNode*> predicates;
  PlanNode* n = agg.children[0].get();
  while (n->kind == PK_FILTER) {
    predicates.push_back(n->predicate.get());
    n = n->children[0].get();
  }
  if (n->kind != PK_SCAN) return nullptr;
  if (n->predicate) predicates.push_back(n->predicate.get());
  View v = scan_view(*n);
  if (v.num_batches == 0 || v.num_rows == 0 || v.num_rows >= ((int64_t)1 << 40)) return nullptr;
  std::shared_ptr<FusedPlan> fp = std::static_pointer_cast<FusedPlan>(agg.fused_cache);
  bool fresh = false;
  if (fp) {
    fresh = fp->n_rows == v.num_rows && fp->n_batches == v.num_batches && fp->col_ids.size() == v.cols.size();
    for (size_t i = 0; fresh && i < v.cols.size(); ++i) fresh = fp->col_ids[i] == v.cols[i].base.get() && !v.cols[i].idx;
  }
  if (!fresh) {
    fp = std::make_shared<FusedPlan>();
    fp->usable = analyze_fused(agg, predicates, v, *fp);
    fp->n_rows = v.num_rows;
    fp->n_batches = v.num_batches;
    for (auto& c : v.cols) fp->col_ids.push_back(c.base.get());
    agg.fused_cache = fp;
  }
  *vout = v;
  return fp;
}

// Sharded (multi-GPU) variant: scan + epilogue with the peer exchange and the merge inside the epilogue.  false: this
// rank's plan is not DENSE-eligible (the caller packs the generic accumulators into the same state block instead).
bool try_fused_scan_aggregate_sharded(PlanNode& agg, int64_t row_offset, int max_groups, View* out) {
  View v;
  std::shared_ptr<FusedPlan> fp = fused_plan_for(agg, &v);
  if (!fp || !fp->usable || fp->P.mode != FM_DENSE) return false;
  DenseRun* D = prepare_dense(agg, *fp);
  if (!D) return false;
  *out = run_dense(agg, *fp, *D, true, row_offset, max_groups);
  return true;
}

bool try_fused_scan_aggregate(PlanNode& agg, View* out) {
  View v;
  std::shared_ptr<FusedPlan> fp = fused_plan_for(agg, &v);
  if (!fp || !fp->usable) return false;
  *out = run_fused(agg, v, *fp);
  return true;
}

// ------------------------------------------------------------------------------------------------
// Multi-GPU exchange of a high-cardinality group-by (BASELINE.json configs[3]): the radix aggregate cut
// at its level-1 scatter.  Level-1 bucket b (top 8 hash bits) belongs to rank b % world; every rank scatters its
// tuples STRAIGHT INTO THE OWNER'S tuple arrays with peer-to-peer stores over NVLink (one kernel partitions and
// exchanges; no send buffers, no NCCL all-to-all), then runs level 2 + the final pass on the buckets it owns.
//   sketch   local 256-bin histogram + HyperLogLog registers                     (caller: all-gather, 17 KB per rank)
//   prepare  owner-side bucket offsets, this rank's write cursors inside every owner's arrays, receive buffers
//            (cudaMalloc, exported as CUDA IPC handles)                           (caller: all-gather of the handles)
//   scatter  k_radix_scatter<1> with peer destinations                            (caller: barrier)
//   finish   level 2 + final pass over the received tuples -> the aggregate's result (consumed by the next execute)
// ------------------------------------------------------------------------------------------------
namespace {
struct ExHandle {  // what a rank publishes per tuple component
  uint64_t raw_ptr, pid;
  cudaIpcMemHandle_t ipc;
};
struct RadixExchange {
  Ctx* ctx = nullptr;
  FParams P;   // the plan re-based on the GLOBAL key ranges: a key packs to the same code (and hash) on every rank
  int key_shift[F_MAXK] = {0, 0, 0, 0}, key_bits[F_MAXK] = {0, 0, 0, 0};
  RParams R;
  DBufP st;
  View view;
  int world = 0, rank = 0;
  double est_owned = 0;
  int64_t n_owned = 0;
  unsigned int n_tiles2 = 0;
  void* recv[R_MAXCOMP] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t recv_cap = 0;  // bytes per component
  int recv_pair = 0;    // buffer 1 sized for 16-byte pairs (RParams::pair12)
  std::vector<std::pair<std::string, void*>> opened;
  ~RadixExchange() {
    for (auto& o : opened) cudaIpcCloseMemHandle(o.second);
    for (void* p : recv)
      if (p) cudaFree(p);
  }
};
}  // namespace

static std::shared_ptr<FusedPlan> exchange_plan(PlanNode& root, PlanNode** agg_out, View* v) {
  PlanNode* agg = find_aggregate_node(root);
  std::shared_ptr<FusedPlan> fp = fused_plan_for(*agg, v);
  if (!fp || !fp->usable || !fp->radix_ok)
    throw_internal("exchange needs a radix-eligible group-by over a table scan (integer-like keys, at most 5 operand values, "
                   "64-bit sums)");
  *agg_out = agg;
  return fp;
}

void radix_exchange_keystats(PlanNode& root, int64_t* stats, int32_t* n_keys) {
  PlanNode* agg;
  View v;
  std::shared_ptr<FusedPlan> fp = exchange_plan(root, &agg, &v);
  *n_keys = fp->P.n_keys;
  for (int k = 0; k < fp->P.n_keys; ++k) {
    stats[2 * k] = fp->key_min[k];
    stats[2 * k + 1] = fp->key_max[k];
  }
}

int radix_exchange_sketch(PlanNode& root, const int64_t* global_stats, void** dev_buf, int64_t* bytes) {
  PlanNode* agg;
  View v;
  std::shared_ptr<FusedPlan> fp = exchange_plan(root, &agg, &v);
  Ctx* ctx = agg->ctx;
  auto ex = std::static_pointer_cast<RadixExchange>(fp->exchange);
  if (!ex) {
    ex = std::make_shared<RadixExchange>();
    ex->ctx = ctx;
    fp->exchange = ex;
  }
  // re-base the packed key on the global value ranges (min over ranks, max over ranks)
  ex->P = fp->P;
  int shift = 0;
  for (int k = 0; k < ex->P.n_keys; ++k) {
    const i128 lo = global_stats[2 * k], hi = global_stats[2 * k + 1];
    if (lo > fp->key_min[k] || hi < fp->key_max[k]) throw_internal("exchange: the global key range does not cover this rank's keys");
    const i128 range = hi - lo + 1;
    int bits = 0;
    while (((i128)1 << bits) < range) ++bits;
    ex->P.keys[k].base = (int64_t)lo;
    ex->P.keys[k].mult = shift >= 64 ? 0 : (1ull << shift);
    ex->key_shift[k] = shift;
    ex->key_bits[k] = bits;
    shift += bits;
  }
  *dev_buf = nullptr;
  *bytes = (int64_t)R_SKETCH_BYTES;
  if (shift > 64) return 1;  // the keys do not pack into 64 bits over the global ranges (the same answer on every rank)
  ex->R = fp->R;
  ex->view = v;
  ex->st = radix_state_block(ctx, ex->R);
  const FParams& P = ex->P;
  radix_launch_hist1(ctx, P, ex->R);
  *dev_buf = ex->st->ptr;
  return 0;
}

int radix_exchange_prepare(PlanNode& root, const void* gathered_host, int world, int rank, void* handles_out, int32_t* n_handles) {
  PlanNode* agg;
  View v;
  std::shared_ptr<FusedPlan> fp = exchange_plan(root, &agg, &v);
  Ctx* ctx = agg->ctx;
  auto ex = std::static_pointer_cast<RadixExchange>(fp->exchange);
  if (!ex || !ex->st) throw_internal("exchange: prepare without sketch");
  if (world < 1 || world > 8 || rank < 0 || rank >= world) throw_internal("exchange: world must be in [1, 8]");
  ex->world = world;
  ex->rank = rank;
  RParams& R = ex->R;
  // ---- merge the sketches: HyperLogLog registers by max, histograms kept per source rank --------------------------------
  std::vector<unsigned int> hll(R_HLL_M, 0);
  std::vector<const unsigned int*> hist((size_t)world);
  for (int s = 0; s < world; ++s) {
    const unsigned int* blk = (const unsigned int*)((const char*)gathered_host + (size_t)s * R_SKETCH_BYTES);
    for (int i = 0; i < R_HLL_M; ++i) hll[i] = std::max(hll[i], blk[RO_HLL / 4 + i]);
    hist[s] = blk + RO_HIST1 / 4;
  }
  const double est = hll_estimate(hll.data());
  // ---- bucket b belongs to rank b % world: owner-side layout and this rank's cursors ------------------------------------
  std::vector<unsigned long long> off1(R_P1 + 1, 0), cur1(R_P1, 0), owner_fill((size_t)world, 0);
  std::vector<unsigned int> tpre(R_P1 + 1, 0);
  std::vector<int64_t> owned((size_t)world, 0);
  int64_t total = 0;
  for (int b = 0; b < R_P1; ++b) {
    const int o = b % world;
    unsigned long long tot_b = 0, before_me = 0;
    for (int s = 0; s < world; ++s) {
      if (s < rank) before_me += hist[s][b];
      tot_b += hist[s][b];
    }
    cur1[b] = owner_fill[o] + before_me;  // where my tuples of bucket b start inside the owner's arrays
    owner_fill[o] += tot_b;
    owned[o] += (int64_t)tot_b;
    total += (int64_t)tot_b;
    const unsigned long long mine_b = o == rank ? tot_b : 0;  // as an owner: non-owned buckets are empty here
    off1[b + 1] = off1[b] + mine_b;
    tpre[b + 1] = tpre[b] + (unsigned int)((mine_b + R_T - 1) / R_T);
  }
  // eligibility must be the same decision on every rank: test every owner's share
  int ok = 1;
  for (int o = 0; o < world && ok; ++o) {
    if (owned[o] == 0) continue;
    const double g_o = std::min(est * ((double)owned[o] / (double)std::max<int64_t>(total, 1)) * 1.05 + 64.0, (double)owned[o]);
    int b2, cap, rc;
    if (!radix_choose(R, g_o, owned[o], (R_P1 - o + world - 1) / world, &b2, &cap, &rc)) ok = 0;
  }
  *n_handles = R.n_comp;
  if (!ok) return 1;
  ex->n_owned = owned[rank];
  ex->n_tiles2 = tpre[R_P1];
  ex->est_owned = est * ((double)owned[rank] / (double)std::max<int64_t>(total, 1));
  const double g_me = std::min(ex->est_owned * 1.05 + 64.0, (double)std::max<int64_t>(owned[rank], 1));
  if (!radix_choose(R, g_me, std::max<int64_t>(owned[rank], 1), (R_P1 - rank + world - 1) / world, &R.b2, &R.cap, &R.row_cap)) return 1;
  char* sp = (char*)ex->st->ptr;
  ctx->h2d(sp + RO_OFF1, off1.data(), (R_P1 + 1) * 8);
  ctx->h2d(sp + RO_TPRE, tpre.data(), (R_P1 + 1) * 4);
  ctx->h2d(sp + RO_CUR1, cur1.data(), R_P1 * 8);
  ctx->sync();
  // ---- receive buffers: plain cudaMalloc (exportable through CUDA IPC), kept across executions ----------------------------
  R.pair12 = radix_pair12(ex->P, R) ? 1 : 0;  // the same on every rank (plan shape + environment); buffer 1 then holds 16 B pairs
  const size_t need = (size_t)owned[rank] * 8 + 256;
  if (need > ex->recv_cap || ex->recv_pair != R.pair12) {
    ex->recv_pair = R.pair12;
    ctx->sync();
    for (int c = 0; c < R_MAXCOMP; ++c) {
      if (ex->recv[c]) CUDA_CHECK(cudaFree(ex->recv[c]));
      ex->recv[c] = nullptr;
    }
    ex->recv_cap = need + need / 8 + ((size_t)1 << 20);
    for (int c = 0; c < R.n_comp; ++c) CUDA_CHECK(cudaMalloc(&ex->recv[c], ex->recv_cap * ((R.pair12 && c == 1) ? 2 : 1)));
  }
  ExHandle* hs = (ExHandle*)handles_out;
  for (int c = 0; c < R.n_comp; ++c) {
    memset(&hs[c], 0, sizeof(ExHandle));
    hs[c].raw_ptr = (uint64_t)(uintptr_t)ex->recv[c];
    hs[c].pid = (uint64_t)getpid();
    if (world > 1) CUDA_CHECK(cudaIpcGetMemHandle(&hs[c].ipc, ex->recv[c]));
  }
  return 0;
}

void radix_exchange_scatter(PlanNode& root, const void* all_handles) {
  PlanNode* agg;
  View v;
  std::shared_ptr<FusedPlan> fp = exchange_plan(root, &agg, &v);
  Ctx* ctx = agg->ctx;
  auto ex = std::static_pointer_cast<RadixExchange>(fp->exchange);
  if (!ex || ex->world == 0) throw_internal("exchange: scatter without prepare");
  RParams& R = ex->R;
  const ExHandle* hs = (const ExHandle*)all_handles;
  R.world = ex->world;
  for (int o = 0; o < ex->world; ++o)
    for (int c = 0; c < R.n_comp; ++c) {
      const ExHandle& h = hs[(size_t)o * R.n_comp + c];
      void* ptr = nullptr;
      if (o == ex->rank) {
        ptr = ex->recv[c];
      } else if (h.pid == (uint64_t)getpid()) {
        ptr = (void*)(uintptr_t)h.raw_ptr;  // ranks emulated inside one process (tests): the pointer is directly usable
      } else {
        const std::string key((const char*)&h.ipc, sizeof(h.ipc));
        for (auto& op : ex->opened)
          if (op.first == key) ptr = op.second;
        if (!ptr) {
          CUDA_CHECK(cudaIpcOpenMemHandle(&ptr, h.ipc, cudaIpcMemLazyEnablePeerAccess));
          ex->opened.push_back({key, ptr});
        }
      }
      R.peer_a[o][c] = (unsigned long long*)ptr;
    }
  for (int c = 0; c < R.n_comp; ++c) R.tup_a[c] = (unsigned long long*)ex->recv[c];
  radix_launch_scatter1(ctx, ex->P, R);
  ctx->sync();  // my stores have left; the caller's barrier then makes every rank's stores complete
  ctx->trace("exchange: scatter over peer memory");
}

int radix_exchange_finish(PlanNode& root) {
  PlanNode* agg;
  View v;
  std::shared_ptr<FusedPlan> fp = exchange_plan(root, &agg, &v);
  auto ex = std::static_pointer_cast<RadixExchange>(fp->exchange);
  if (!ex || ex->world == 0) throw_internal("exchange: finish without scatter");
  View out;
  int fail = 0;
  if (!radix_tail(*agg, ex->view, *fp, ex->P, ex->key_shift, ex->key_bits, ex->R, ex->st, ex->n_owned, ex->n_tiles2, ex->est_owned, &out,
                  &fail))
    return fail ? fail : 1;
  agg->strategy = "exchange[rank " + std::to_string(ex->rank) + "/" + std::to_string(ex->world) + ", p2p scatter] -> " + agg->strategy;
  agg->merged_override = std::make_shared<View>(out);
  return 0;
}

}  // namespace qgpu
