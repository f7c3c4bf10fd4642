// Physical plan nodes: the C++ host-side mirror of qurious/src/physical/plan/* for the hot path.
#pragma once
#include "ops.h"

namespace qgpu {

enum PlanKind { PK_SCAN = 1, PK_FILTER, PK_PROJECTION, PK_AGGREGATE, PK_HASH_JOIN, PK_SORT, PK_LIMIT, PK_NL_JOIN, PK_CROSS_JOIN, PK_BROADCAST, PK_FINAL_AGG };

struct AggDesc {
  int op = 0;
  std::shared_ptr<ExprNode> expr;
  DType return_type;
  DType expr_type;
};

struct PlanNode {
  Ctx* ctx = nullptr;
  int kind = 0;
  Schema schema;  // PhysicalPlan::schema()
  std::vector<std::shared_ptr<PlanNode>> children;
  // Scan
  std::shared_ptr<TableImpl> table;
  bool has_projection = false;
  std::vector<int> projection;
  std::shared_ptr<ExprNode> predicate;  // Scan filter / Filter predicate (may be null for Scan)
  // Projection
  std::vector<std::shared_ptr<ExprNode>> exprs;
  // Aggregate
  std::vector<std::shared_ptr<ExprNode>> group_exprs;
  std::vector<AggDesc> aggs;
  // HashJoin
  int join_type = QGPU_JOIN_INNER;
  std::vector<std::shared_ptr<ExprNode>> left_on, right_on;
  bool has_join_filter = false;
  std::shared_ptr<ExprNode> join_filter_expr;
  Schema join_filter_schema;
  std::vector<int> join_filter_index, join_filter_side;
  // Sort (exprs = sort expressions) / Limit
  std::vector<int> sort_desc, sort_nulls_first;
  int64_t sort_limit = -1;
  int64_t limit_fetch = -1, limit_skip = 0;
  // stats of the last execute
  double last_ms = 0;
  int64_t last_launches = 0;
  std::string strategy = "not-executed";
  std::string strategy_desc;
  std::shared_ptr<void> fused_cache;
  std::shared_ptr<Speculation> spec;  // learned device-side counts of this subtree's fused pipeline (fused.cu)
  uint64_t spec_sig = 0;              // identity of the scanned tables the counts were learned on
  // exchange operators (exchange.cu): Broadcast memo; FinalAggregate key / value columns and merge operators (0 SUM, 1 MIN, 2 MAX)
  std::shared_ptr<void> exchange_cache;
  std::vector<int> exchange_keys, exchange_cols, exchange_ops;
  // pruned Broadcast: rows travel only to the ranks whose probe-side key range (statistics of prune_table's column) holds their key
  std::shared_ptr<TableImpl> prune_table;
  int prune_key_col = -1, prune_probe_col = -1;
  bool order_free = false;  // the consumer does not depend on this node's output row order (qgpu_plan_set_order_free)
  // sharded execution (shard.cu): stop before finalisation / resume from merged states
  AggPending* defer = nullptr;
  std::shared_ptr<View> merged_override;
  std::shared_ptr<AggPending> shard_pending;  // kept between qgpu_plan_partial_state and qgpu_plan_execute_merged  // fused.cu: analysis of this Aggregate <- (Filter)* <- Scan subtree

  View execute();
  View child_view(int i);  // children[i]->execute() with the result metadata resolved
};

// the (unfiltered) view of a Scan node's table: consolidates appended batches, applies the projection
View scan_view(PlanNode& scan);

// shard.cu
PlanNode* find_aggregate_node(PlanNode& root);
void shard_partial_state(PlanNode& root, int64_t row_offset, int32_t max_groups, void* out_buf, int64_t cap_bytes);
int64_t shard_state_bytes(PlanNode& root, int32_t max_groups);
View shard_execute_merged(PlanNode& root, const void* gathered, int32_t n_states, int32_t max_groups);
View shard_execute_fused(PlanNode& root, int64_t row_offset, int32_t max_groups);

// fused.cu: multi-GPU exchange of a radix-partitioned group-by (see radix_exchange_* there)
void radix_exchange_keystats(PlanNode& root, int64_t* stats, int32_t* n_keys);
int radix_exchange_sketch(PlanNode& root, const int64_t* global_stats, void** dev_buf, int64_t* bytes);
int radix_exchange_prepare(PlanNode& root, const void* gathered_host, int world, int rank, void* handles_out, int32_t* n_handles);
void radix_exchange_scatter(PlanNode& root, const void* all_handles);
int radix_exchange_finish(PlanNode& root);

// exchange.cu: the two exchange operators of a distributed plan
View run_broadcast(PlanNode& node);
uint64_t broadcast_signature(PlanNode& node);  // executes the node (memoised per C-ABI call): identity of what every rank contributed
View run_final_aggregate(PlanNode& node);
// fused.cu: runs `body` under the speculation of `owner` (learned device-side counts replayed without host round trips and
// verified at the end; a wrong guess re-runs in learning mode).  `body` must not issue collectives.
View run_speculated(PlanNode& owner, uint64_t signature, const std::function<View()>& body);

// sort.cu
View run_sort(PlanNode& node, const View& input);
View run_limit(PlanNode& node, const View& input);

std::shared_ptr<TableImpl> hash_partition_table(TableImpl& t, int key_col, int n_parts, std::vector<int64_t>& offsets);

// fused.cu: returns true and fills `out` when the aggregate over this input can run as one fused
// scan+filter+aggregate pipeline kernel.
bool try_fused_scan_aggregate(PlanNode& agg, View* out);
bool try_fused_scan_aggregate_sharded(PlanNode& agg, int64_t row_offset, int max_groups, View* out);
// an Inner join with unique integer build keys as ONE probe-scan kernel emitting (build row, probe row) pairs in
// arbitrary order; only where the order does not matter (build sides of fused join-aggregates, order_free nodes)
bool fused_unordered_join(PlanNode& join, View* out);
// fused.cu: Aggregate <- HashJoin(Inner, unique build keys) <- [build plan, probe scan]: probe + aggregate in one kernel
bool try_fused_join_aggregate(PlanNode& agg, View* out);

}  // namespace qgpu

struct qgpu_plan {
  std::shared_ptr<qgpu::PlanNode> node;
};
