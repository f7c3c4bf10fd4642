// Operator implementations (generic, interpreter driven).  Fused fast paths live in fused.cu.
#pragma once
#include "expr.h"
#include "kernels.h"

namespace qgpu {

struct AggSpec {
  int op = 0;  // qgpu_agg_op
  std::shared_ptr<Compiled> arg;
  DType return_type;
  DType expr_type;
};

enum AccKind : int {
  AK_COUNT = 0, AK_SUM_I64, AK_SUM_DEC, AK_SUM_F64, AK_MIN_I64, AK_MAX_I64, AK_MIN_U64, AK_MAX_U64, AK_MIN_F64, AK_MAX_F64,
  AK_MIN_DEC, AK_MAX_DEC
};

// Per-group accumulator state handed from an accumulation strategy (generic atomics in ops.cu or
// the fused pipeline in fused.cu) to the common finalisation (AVG division, NULL-ness, output order).
struct GroupAccs {
  int64_t n_groups = 0;
  std::vector<int> kind;            // AccKind per aggregate
  std::vector<DBufP> lo, hi, cnt;   // per aggregate: u64[n_groups] each (cnt = number of non-NULL inputs); hi may be
                                    // null: the high word is then the sign extension of lo (0 for f64 / unsigned kinds)
  DBufP first_row;                  // i64[n_groups]: first input row of the group
  // optional: the actual group count is still on the device (int64); n_groups is then an upper bound and
  // finish_aggregate reads the count together with its own flags in ONE device->host copy
  DBufP n_groups_dev;
  // unordered: keep the producer's group order (no first-occurrence ranking; needs key_cols).  The reference's
  // order is unspecified (hash.rs:98); the radix-partitioned aggregate uses this for 10^8-group results.
  bool unordered = false;
  unsigned long long side_word = 0;  // out: the word following the key NULL counts (see finish_aggregate)
};
// key_cols (optional): the group key VALUES as gid-indexed device columns (sharded execution: the first row of a
// merged group may live on another shard, so keys travel with the state instead of being gathered from `input`);
// key_nulls: device u64[n_keys] NULL counts of those columns, fetched together with the other flags.
View finish_aggregate(Ctx* ctx, const View& input, std::vector<std::shared_ptr<Compiled>>& keys, std::vector<AggSpec>& aggs,
                      const Schema& out_schema, GroupAccs& accs, std::vector<DColP>* key_cols = nullptr,
                      const unsigned long long* key_nulls = nullptr, bool allow_pending = false);
// allow_pending: when the group count is already exact on the host, the error code and the aggregates' NULL counts
// may stay in flight (View::pending) instead of costing a host round trip here

// An aggregate stopped right before finalisation: per-group accumulator state + what finish_aggregate needs.
struct AggPending {
  bool set = false;
  View input;
  std::vector<std::shared_ptr<Compiled>> keys;
  std::vector<AggSpec> specs;
  GroupAccs accs;
};

// HashAggregate / NoGroupingAggregate (keys.empty()).  SURVEY 8a a7,a8,a10-a13.
View run_aggregate(Ctx* ctx, const View& input, std::vector<std::shared_ptr<Compiled>>& keys, std::vector<AggSpec>& aggs,
                   const Schema& out_schema, AggPending* defer = nullptr);

struct JoinFilterSpec {
  std::shared_ptr<Compiled> expr;  // compiled against `schema`
  Schema schema;
  std::vector<int> column_index;
  std::vector<int> column_side;  // 0 = Left/build, 1 = Right/probe
};

// HashJoinExec::execute.  SURVEY 8a a14, a15.
View run_hash_join(Ctx* ctx, const View& build, const View& probe, int join_type,
                   std::vector<std::shared_ptr<Compiled>>& left_on, std::vector<std::shared_ptr<Compiled>>& right_on,
                   JoinFilterSpec* filter, const Schema& out_schema);

// NestedLoopJoinExec::execute (nest_loop_join.rs:79-228).  SURVEY 8f #3.
View run_nested_loop_join(Ctx* ctx, const View& left, const View& right, int join_type, JoinFilterSpec* filter,
                          const Schema& out_schema);

// CrossJoin::execute (join/cross_join.rs:118-168): the cartesian product, left row major.  The reference emits one batch
// per (left batch, right batch, left row); with one batch per side that is exactly this order, otherwise the same rows
// in an order that depends on the inputs' batch boundaries.
View run_cross_join(Ctx* ctx, const View& left, const View& right, const Schema& out_schema);

Schema build_join_schema(const Schema& left, const Schema& right, int join_type);
// rows of [0, n) whose bit in the bitmap is (invert ? clear : set), ascending
IdxP rows_by_bit(Ctx* ctx, const DBufP& bits, int64_t n, int invert);

// MIN / MAX start value of the accumulator for argument type `at` (min.rs / max.rs: NATIVE::MAX / NATIVE::MIN of the array's
// native type) as the (lo, hi) words the accumulators and finish_aggregate use; floats travel as their total-order key.
// An ungrouped MIN/MAX over zero qualifying rows returns exactly this value.
void minmax_sentinel(const DType& at, bool is_min, unsigned long long* lo, unsigned long long* hi);

// physical layout of an output column of logical type `t` (canonical Arrow layout)
Phys out_phys_of(const DType& t);

// validation helpers shared with the fused paths
void validate_agg_types(const std::vector<AggSpec>& aggs);
void check_hash_key_type(const DType& t);

// ---- shared device hash-table machinery (ops.cu) ----------------------------------------------
struct KeyTable {
  DBufP slots;     // u64 per slot: (hash32 << 32) | (rep_row + 1); 0 = empty
  DBufP slot_gid;  // u32 per slot
  DBufP gid_rep;   // i64 per group: representative row (CAS winner)
  DBufP row_gid;   // u32 per input row: group id, 0xffffffff = no group (NULL key, joins only)
  int64_t capacity = 0;
  int64_t n_groups = 0;
};
// Group the rows of `v` by key equality.  skip_null_keys: rows with any NULL key get no group (joins).
KeyTable build_key_table(Ctx* ctx, const View& v, std::vector<std::shared_ptr<Compiled>>& keys, bool skip_null_keys,
                         DBufP* key_programs_out);

}  // namespace qgpu
