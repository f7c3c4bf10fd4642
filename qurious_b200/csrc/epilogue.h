// Single-CTA epilogue of the DENSE fused scan-aggregate (fused.cu): everything that used to follow the scan kernel
// as 5-6 small launches, ~25 stream-ordered allocations and a host round trip -- export of the occupied slots,
// first-occurrence ranking, finalisation (AVG division, NULL-ness), key columns, re-initialisation of the global
// accumulator table -- as ONE kernel behind the scan, with the result metadata travelling to a pinned slot
// asynchronously (qgpu_internal.h: Pending).
//
// With a communicator (comm.h) the SAME kernel is the multi-GPU merge of row-range shards: it packs this
// rank's groups into a state block, stores the block into every peer's symmetric buffer over NVLink, waits on the
// peers' epoch flags, merges all blocks exactly (128-bit integer adds in rank order: bit-identical to one GPU over
// the whole table) and finalises -- aggregate -> exchange -> merge -> finalise without a collective call or a host
// round trip in between.
//   HashAggregate / NoGroupingAggregate   qurious/src/physical/plan/aggregate/hash.rs:89-107, no_grouping.rs:48-61
//   accumulator evaluate()                qurious/src/physical/expr/aggregate/*.rs (finalize.cuh)
#pragma once
#include "finalize.cuh"

namespace qgpu {

constexpr int EPI_NT = 256;     // small enough (registers) to share an SM with a CTA of the next scan
constexpr int EPI_MAXG = 4096;   // groups (and gathered records) one epilogue CTA handles
constexpr int EPI_MAXK = 8, EPI_MAXACC = 8, EPI_MAXAGG = 24;
constexpr int EPI_HDR = 8;       // state block header words: magic, n records, n_keys, n_accs, overflow, record words
#define EPI_MAGIC 0x5147505545504931ULL  // "QGPUEPI1"

// metadata block (words): [0] groups, [1] EvalErr, [2] merge / exchange error, [3 + a] NULL count of aggregate a,
// [3 + EPI_MAXAGG + k] NULL count of key column k
constexpr int EPI_META_WORDS = 3 + EPI_MAXAGG + EPI_MAXK;
// state block: EPI_HDR header words, then one record per group:
//   n_keys x {tag, w1, w2} | first row (global) | n_aggs x {lo, hi, count}
//   tag: 0 = NULL key, 1 = fixed-width value (w1 = lo, w2 = hi, sign-extended), 2 + len = Utf8 of len <= 16 bytes
// Key VALUES travel, so blocks of different shards match by value: no dictionary or statistics need to agree, and a
// shard whose local plan ran on the generic operators (k_pack_records, shard.cu) produces the same block.
static inline int epi_rec_words(int n_keys, int n_aggs) { return 3 * n_keys + 1 + 3 * n_aggs; }
enum { EPI_SRC_DENSE = 0, EPI_SRC_PACKED = 1 };
enum { EPI_ERR_NONE = 0, EPI_ERR_TIMEOUT = 10, EPI_ERR_BLOCK = 11, EPI_ERR_OVERFLOW = 12, EPI_ERR_TOO_MANY = 13 };

struct EpiKey {
  int out_phys;           // Phys of the output key column (Arrow layout: PH_I8..PH_U64, PH_I128, PH_STR)
  int is_dict;            // the slot digit is the dictionary code of a Utf8 column (strings of <= 16 bytes)
  long long base;         // integer-like keys: value = base + digit
  unsigned int mult, range;  // digit = (slot / mult) % range
  const int32_t* dict_offs;  // dictionary strings on the device
  const char* dict_data;
  void* out;
  int32_t* out_offsets;   // Utf8
  uint32_t* out_valid;    // may be null when the source cannot produce NULL keys (dense table)
};

struct EpiParams {
  // the scan kernel's global accumulator table: word (slot g, k) = g_lo[g * (n_accs + 2) + k]; k = n_accs: rows, n_accs + 1: first row
  unsigned long long* g_lo;
  unsigned long long* g_hi;
  int n_slots, n_accs, n_keys, n_aggs;
  int src;                             // EPI_SRC_DENSE: records come from the table above; EPI_SRC_PACKED: `rec` is already filled
  int smem_cap;                        // capacity the shared-memory arrays are laid out for (>= max_groups, g_max, gathered records)
  int grouped, max_groups, g_max, rw;  // max_groups: records one rank may send; g_max: output capacity; rw: record words
  long long row_offset;                // global index of this shard's first row (first-occurrence order across shards)
  long long init[EPI_MAXACC + 2];      // the table is re-initialised for the next execution
  int acc_kind[EPI_MAXACC];            // FK_SUM 0 / FK_MIN 1 / FK_MAX 2 / FK_SUMF 3
  int acc_wide[EPI_MAXACC];            // the accumulator has a real high word (DENSE 128-bit sums)
  EpiKey key[EPI_MAXK];
  FinSpec agg[EPI_MAXAGG];             // lo / hi / cnt / stride are set inside the kernel
  int agg_acc[EPI_MAXAGG];             // dense source: accumulator of aggregate a (-1: COUNT)
  unsigned long long sent_lo[EPI_MAXAGG], sent_hi[EPI_MAXAGG];  // typed MIN / MAX start values (ungrouped, zero rows)
  unsigned long long* rec;             // this rank's state block: EPI_HDR + max_groups * rw words
  unsigned long long* mrec;            // merged records (world > 1): g_max * rw words
  unsigned long long* meta;            // EPI_META_WORDS
  // exchange (world > 1): symmetric buffers of comm.h
  int world, rank;
  unsigned long long epoch;
  unsigned long long timeout_ns;
  unsigned long long* peer[8];
};

size_t epilogue_smem_bytes(int cap);
void launch_dense_epilogue(Ctx* ctx, const EpiParams& p);

// FinSpec static fields / accumulator kinds / typed MIN-MAX start values of every aggregate (+ the RecordBatch::try_new
// type checks finish_aggregate performs); kinds[a] = AccKind of aggregate a
void epilogue_describe(Ctx* ctx, const std::vector<DType>& key_types, const std::vector<AggSpec>& specs, const std::vector<int>& kinds,
                       const Schema& out_schema, EpiParams& E);
// Allocates the output columns and scratch of ONE epilogue launch (a single stream-ordered allocation), launches it and
// returns the result with its metadata pending (View::pending).  sharded: exchange + merge over ctx->comm.
// packed_rec: EPI_SRC_PACKED -- the state block k_pack_records filled (EPI_HDR + max_groups * rw words).
View epilogue_execute(Ctx* ctx, EpiParams E, const std::vector<DType>& key_types, const std::vector<AggSpec>& specs,
                      const Schema& out_schema, bool sharded, int64_t row_offset, int max_groups, DBufP packed_rec);

}  // namespace qgpu
