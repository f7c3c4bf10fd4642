/*
 * qgpu.h -- C ABI of libqgpu.so: B200-native physical operators for holicc/qurious.
 *
 * This is the drop-in boundary.  Every entry point is what a one-file Rust
 * `impl PhysicalPlan for Gpu*` shim would bind through `extern "C"`.
 * Data crosses as Arrow C Data Interface structs (arrow-rs: arrow::ffi::{FFI_ArrowSchema,
 * FFI_ArrowArray}, feature "ffi"); expressions cross as a postfix byte stream because the
 * reference's PhysicalExpr structs are opaque (`Arc<dyn PhysicalExpr>`, private fields,
 * qurious/src/physical/expr/mod.rs:33-35) and must be re-serialised from the LogicalExpr tree
 * by a sibling of DefaultQueryPlanner::create_physical_expr (qurious/src/planner/mod.rs:102-152).
 *
 * No torch / C++ types appear here.  All functions return 0 on success, else a qgpu_status;
 * the message is available through qgpu_last_error().  The library never aborts the process
 * and has NO CPU fallback: without a CUDA device qgpu_init fails with QGPU_ERR_CUDA.
 *
 * Threading: a qgpu_ctx may be used from one thread at a time (internal mutex); independent
 * contexts may run concurrently.  Calls are synchronous: they return when the output is ready.
 */
#ifndef QGPU_H
#define QGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- Arrow C Data Interface (https://arrow.apache.org/docs/format/CDataInterface.html) ---- */
#ifndef ARROW_C_DATA_INTERFACE
#define ARROW_C_DATA_INTERFACE
#define ARROW_FLAG_DICTIONARY_ORDERED 1
#define ARROW_FLAG_NULLABLE 2
#define ARROW_FLAG_MAP_KEYS_SORTED 4
struct ArrowSchema {
  const char* format;
  const char* name;
  const char* metadata;
  int64_t flags;
  int64_t n_children;
  struct ArrowSchema** children;
  struct ArrowSchema* dictionary;
  void (*release)(struct ArrowSchema*);
  void* private_data;
};
struct ArrowArray {
  int64_t length;
  int64_t null_count;
  int64_t offset;
  int64_t n_buffers;
  int64_t n_children;
  const void** buffers;
  struct ArrowArray** children;
  struct ArrowArray* dictionary;
  void (*release)(struct ArrowArray*);
  void* private_data;
};
#endif
#ifndef ARROW_C_STREAM_INTERFACE
#define ARROW_C_STREAM_INTERFACE
struct ArrowArrayStream {
  int (*get_schema)(struct ArrowArrayStream*, struct ArrowSchema* out);
  int (*get_next)(struct ArrowArrayStream*, struct ArrowArray* out);
  const char* (*get_last_error)(struct ArrowArrayStream*);
  void (*release)(struct ArrowArrayStream*);
  void* private_data;
};
#endif

/* ---- status codes: mapping to qurious::error::Error (qurious/src/error.rs:41-53) ---------- */
typedef enum qgpu_status {
  QGPU_OK = 0,
  QGPU_ERR_INTERNAL = 1, /* Error::InternalError / unsupported type / unimplemented!() paths */
  QGPU_ERR_ARROW = 2,    /* Error::ArrowError: type mismatch, divide by zero, cast failure */
  QGPU_ERR_CUDA = 3,
  QGPU_ERR_NCCL = 4,
  QGPU_ERR_OOM = 5
} qgpu_status;

/* ---- logical type ids used in the expression IR and in qgpu_agg_desc ---------------------- */
typedef enum qgpu_type_id {
  QGPU_T_NULL = 0, QGPU_T_BOOL = 1, QGPU_T_INT8 = 2, QGPU_T_INT16 = 3, QGPU_T_INT32 = 4,
  QGPU_T_INT64 = 5, QGPU_T_UINT8 = 6, QGPU_T_UINT16 = 7, QGPU_T_UINT32 = 8, QGPU_T_UINT64 = 9,
  QGPU_T_FLOAT32 = 10, QGPU_T_FLOAT64 = 11, QGPU_T_UTF8 = 12, QGPU_T_DATE32 = 13,
  QGPU_T_DATE64 = 14, QGPU_T_DECIMAL128 = 15,
  /* Time32(Second | Millisecond) / Time64(Microsecond | Nanosecond): the unit travels in qgpu_type.scale
   * (0 = s, 1 = ms, 2 = us, 3 = ns); hashable key types of the reference (utils/array.rs:198-201) */
  QGPU_T_TIME32 = 16, QGPU_T_TIME64 = 17
} qgpu_type_id;

typedef struct qgpu_type {
  uint8_t id;        /* qgpu_type_id */
  uint8_t precision; /* Decimal128 only */
  int8_t scale;      /* Decimal128 only */
} qgpu_type;

/* ---- expression IR: postfix byte stream, little-endian ------------------------------------
 *  QGPU_IR_COLUMN      u32 index                    physical/expr/column.rs:24-33
 *  QGPU_IR_LITERAL     type(3B) u8 is_null value    physical/expr/literal.rs:19-23 (ScalarValue)
 *                      value: bool/ints/dates = 8B, floats = f64 8B, decimal128 = 16B,
 *                             utf8 = u32 len + bytes; absent when is_null or type NULL
 *  QGPU_IR_BINARY      u8 op (Operator order of datatypes/operator.rs:4-20: Eq NotEq Gt GtEq Lt
 *                      LtEq And Or Add Sub Mul Div Mod)   physical/expr/binary.rs:30-71
 *  QGPU_IR_CAST        type(3B)                     physical/expr/cast.rs:32-38 (safe:false)
 *  QGPU_IR_CASE        u32 n_when; stack: when1 then1 .. whenN thenN else   case.rs:30-47
 *  QGPU_IR_IS_NULL / QGPU_IR_IS_NOT_NULL / QGPU_IR_NEGATIVE   is_null.rs / is_not_null.rs / negative.rs
 *  QGPU_IR_LIKE        u8 negated; stack: expr pattern     physical/expr/like.rs:28-41 (arrow like / nlike: % _ and \ escape)
 *  QGPU_IR_EXTRACT     u8 part (0 YEAR, 1 MONTH, 2 DAY)    physical/expr/function.rs + functions/datetime/extract.rs (Date32/Date64 -> Int64)
 *  QGPU_IR_SUBQUERY    u64 qgpu_plan* (a plan of the same context)   physical/expr/subquery.rs:15-20: the sub-plan runs once per
 *                      evaluation and its FIRST column stands in as an array operand; like the reference's arrow kernels the
 *                      evaluation fails (ArrowError) unless it has exactly as many rows as the input
 */
typedef enum qgpu_ir_op {
  QGPU_IR_COLUMN = 1, QGPU_IR_LITERAL = 2, QGPU_IR_BINARY = 3, QGPU_IR_CAST = 4, QGPU_IR_CASE = 5,
  QGPU_IR_IS_NULL = 6, QGPU_IR_IS_NOT_NULL = 7, QGPU_IR_NEGATIVE = 8, QGPU_IR_LIKE = 9, QGPU_IR_EXTRACT = 10, QGPU_IR_SUBQUERY = 11
} qgpu_ir_op;

/* aggregate operator: AggregateOperator of logical/expr/aggregate.rs:56-62 */
typedef enum qgpu_agg_op { QGPU_AGG_SUM = 0, QGPU_AGG_MIN = 1, QGPU_AGG_MAX = 2, QGPU_AGG_AVG = 3, QGPU_AGG_COUNT = 4 } qgpu_agg_op;

/* join type: JoinType of common/join_type.rs:4-11 */
typedef enum qgpu_join_type { QGPU_JOIN_LEFT = 0, QGPU_JOIN_RIGHT = 1, QGPU_JOIN_INNER = 2, QGPU_JOIN_FULL = 3, QGPU_JOIN_LEFT_SEMI = 4, QGPU_JOIN_LEFT_ANTI = 5 } qgpu_join_type;

typedef struct qgpu_ctx qgpu_ctx;
typedef struct qgpu_table qgpu_table; /* HBM-resident table: the GpuMemoryTable of SURVEY 8b */
typedef struct qgpu_expr qgpu_expr;   /* parsed PhysicalExpr */
typedef struct qgpu_plan qgpu_plan;   /* a PhysicalPlan node (owns references to its children) */

/* One aggregate: {Sum,Min,Max,Avg,Count}AggregateExpr (physical/expr/aggregate/{sum,min,max,avg,count}.rs).
 * `return_type` is the planner-inferred type (logical/expr/aggregate.rs:65-90);
 * `expr_type` is AvgAggregateExpr::expr_data_type (avg.rs:16-20), ignored otherwise. */
typedef struct qgpu_agg_desc {
  int32_t op; /* qgpu_agg_op */
  const qgpu_expr* expr;
  qgpu_type return_type;
  qgpu_type expr_type;
} qgpu_agg_desc;

/* JoinFilter (physical/plan/join/nest_loop_join.rs:29-40): expr over an intermediate batch whose
 * column i is column `column_index[i]` of side `column_side[i]` (0 = Left/build, 1 = Right/probe). */
typedef struct qgpu_join_filter {
  const qgpu_expr* expr;
  const struct ArrowSchema* schema;
  const int32_t* column_index;
  const int32_t* column_side;
  int32_t n_columns;
} qgpu_join_filter;

/* ---- context ------------------------------------------------------------------------------ */
/* devices: CUDA ordinals this context drives (n = 1 in round 1: one process per GPU). */
int qgpu_init(const int* devices, int n, qgpu_ctx** out);
void qgpu_shutdown(qgpu_ctx* ctx);
/* Last error of this context (or of a failed qgpu_init when ctx == NULL).  Never NULL. */
const char* qgpu_last_error(const qgpu_ctx* ctx);
/* Compatibility switches for reference quirks: name in {"avg_precision",
 * "empty_decimal_sum"}; value 1 reproduces the reference's failure, 0 (default) the intended value. */
int qgpu_set_compat(qgpu_ctx* ctx, const char* name, int value);
/* Run-time specialisation of the fused scan-aggregate kernel (csrc/fused_jit.cu): DENSE plans whose shape has no
 * ahead-of-time kernel get the same hand-written source instantiated for their shape signature through NVRTC (once per
 * process and shape; QGPU_JIT=0 disables it, the generic tile body then runs).  qgpu_jit_compile compiles a signature
 * without a GPU (build checks, tests): returns the CUBIN size in bytes, 0 when the compilation failed or NVRTC is not
 * installed, -1 on an internal error; the compiler log goes to log_buf (NUL terminated, truncated to cap). */
/* Monotonic counters of a context: "alloc_bytes" (bytes of every device allocation: all temporaries an execution
 * creates -- index vectors, gathered columns, accumulator tables), "gather_bytes" (payload bytes written by column
 * gathers: what late materialisation finally copies), "kernel_launches".  -1: unknown name. */
int64_t qgpu_counter(const qgpu_ctx* ctx, const char* name);
int64_t qgpu_jit_compile(const uint64_t* signature4, uint32_t pack, char* log_buf, int64_t cap);
/* Tuning knobs: "ingest_threads" (host worker threads of the staged ingest, 0 = min(hardware threads, 16); at most 16),
 * "ingest_host_narrow" (1: Decimal128(p <= 18) narrowed to int64 by the host workers while staging -- 8 instead of 16
 * bytes per value cross PCIe; 0: uploaded as 16-byte values and narrowed by one kernel; -1: automatic = on, except for
 * large page-locked sources when this rank has fewer than 8 worker threads: those are DMA'd directly);
 * "pool_reserve_mb": the context's device memory pool keeps at least this much reserved (default 4096; grown now, kept by
 * qgpu_release_cached_memory), so that no execution waits for the driver to map memory in the middle of a step. */
int qgpu_set_option(qgpu_ctx* ctx, const char* name, int64_t value);
/* number of kernel launches issued by this context since creation (bench.py: gpu_launches) */
int64_t qgpu_kernel_launches(const qgpu_ctx* ctx);
/* The context keeps large freed device blocks (>= 256 MB, up to 96 GB) for re-use by the next execution of the same plan;
 * this returns them to the driver (e.g. before another library allocates on the same GPU). */
int qgpu_release_cached_memory(qgpu_ctx* ctx);
/* The context's compute stream (a cudaStream_t) so that callers can record their own CUDA events
 * around calls (bench.py times steps with events on THIS stream). */
void* qgpu_ctx_stream(const qgpu_ctx* ctx);
/* Per-kernel device timing: while enabled every kernel launch is bracketed by CUDA events on the
 * compute stream (on == 1), or only the launches of at least `on` thread blocks (on >= 2: the large-grid kernels a
 * roofline is quoted for, without perturbing a sub-millisecond step).  qgpu_profile_report synchronises, writes "kernel\tlaunches\ttotal_ms\tmax_ms\n"
 * lines for the launches since the last report into buf (NUL terminated, truncated to cap) and
 * returns the number of bytes the full report needs. */
int qgpu_profile_enable(qgpu_ctx* ctx, int on);
int64_t qgpu_profile_report(qgpu_ctx* ctx, char* buf, int64_t cap);

/* ---- tables: replaces MemoryTable (datasource/memory.rs:20-45) ---------------------------- */
int qgpu_table_create(qgpu_ctx* ctx, const struct ArrowSchema* schema, qgpu_table** out);
/* MemoryTable::insert (memory.rs:104-111): appends one RecordBatch (struct array).  The batch is validated and
 * RETAINED (ownership moves to the table; batch->release is called after the upload, or by qgpu_table_free): all
 * retained batches are uploaded at once when the table is first used or on qgpu_table_flush -- host worker threads
 * gather the many small batches the reference produces (1024 rows: datasource/file/csv.rs:34-72) through a pinned ring
 * into the final contiguous device columns, narrowing Decimal128(p <= 18) to int64 on the way (ingest.cu).  Only the
 * columns later referenced need to be uploaded: pass upload_columns = NULL for all, else n indices. */
int qgpu_table_append(qgpu_table* t, struct ArrowArray* batch, const int32_t* upload_columns, int32_t n);
/* The same for every batch of an Arrow C stream (arrow-rs: FFI_ArrowArrayStream over a RecordBatchReader; one FFI call
 * instead of one per 1024-row batch).  Takes ownership of the stream (released before returning); *out_batches (may be
 * NULL) = batches appended.  On an error the batches appended so far stay. */
int qgpu_table_append_stream(qgpu_table* t, struct ArrowArrayStream* stream, const int32_t* upload_columns, int32_t n,
                             int64_t* out_batches);
/* Upload the retained batches now (otherwise: first use). */
int qgpu_table_flush(qgpu_table* t);
/* File -> HBM ingest: replaces read_csv / `COPY t FROM 'x.tbl' (DELIMITER '|')` (datasource/file/csv.rs:16-72,
 * planner/sql.rs:324-375) for a table whose schema is declared (CREATE TABLE): the raw bytes go to the device and are
 * split and parsed there into the table's column types (csrc/csv.cu) -- no host-side Arrow batches at all.
 * CsvReadOptions (csv.rs:16-32): has_header, delimiter (0 = ','), quote / escape (must be 0: not supported on the
 * device -> QGPU_ERR_INTERNAL).  An empty field is NULL; a trailing delimiter yields one more (empty) field.
 * _csv: `text` is a HOST buffer of `len` bytes; _csv_file: the library reads the file itself (pread into pinned
 * memory).  *out_rows (may be NULL) = rows appended.  Parse errors: QGPU_ERR_ARROW with line and column. */
typedef struct qgpu_csv_options {
  uint8_t has_header, delimiter, quote, escape;
} qgpu_csv_options;
int qgpu_table_append_csv(qgpu_table* t, const void* text, int64_t len, const qgpu_csv_options* options,
                          const int32_t* upload_columns, int32_t n, int64_t* out_rows);
int qgpu_table_append_csv_file(qgpu_table* t, const char* path, const qgpu_csv_options* options,
                               const int32_t* upload_columns, int32_t n, int64_t* out_rows);
/* Same, but every buffer pointer inside `batch` is a DEVICE pointer on this context's GPU
 * (Arrow C Device Data Interface, device_type ARROW_DEVICE_CUDA); buffers are copied D2D. */
int qgpu_table_append_device(qgpu_table* t, struct ArrowArray* batch);
int64_t qgpu_table_num_rows(const qgpu_table* t);
int64_t qgpu_table_num_batches(const qgpu_table* t);
/* bytes of column `col` resident in HBM in the layout the kernels read (after narrowing) */
int64_t qgpu_table_column_bytes(const qgpu_table* t, int32_t col);
int qgpu_table_schema(const qgpu_table* t, struct ArrowSchema* out);
/* download as one RecordBatch (struct array) */
int qgpu_table_export(qgpu_table* t, struct ArrowArray* out_array, struct ArrowSchema* out_schema);
void qgpu_table_free(qgpu_table* t);

/* ---- expressions --------------------------------------------------------------------------- */
int qgpu_expr_parse(qgpu_ctx* ctx, const uint8_t* ir, size_t len, qgpu_expr** out);
void qgpu_expr_free(qgpu_expr* e);

/* ---- plan nodes: constructor argument lists mirror the reference's ------------------------ */
/* Scan::new(schema, datasource, projections, filter) (physical/plan/scan.rs:22-35) over a
 * GpuMemoryTable; execute == MemoryTable::scan(projection, filters) (memory.rs:69-98).
 * projection: column indices or NULL (n_projection ignored). filter may be NULL. */
int qgpu_plan_scan(qgpu_ctx* ctx, qgpu_table* datasource, const int32_t* projection, int32_t n_projection,
                   const qgpu_expr* filter, qgpu_plan** out);
/* Filter::new(input, predicate) (physical/plan/filter.rs:18-20) */
int qgpu_plan_filter(qgpu_ctx* ctx, qgpu_plan* input, const qgpu_expr* predicate, qgpu_plan** out);
/* Projection::new(schema, input, exprs) (physical/plan/projection.rs:17-19) */
int qgpu_plan_projection(qgpu_ctx* ctx, const struct ArrowSchema* schema, qgpu_plan* input,
                         const qgpu_expr* const* exprs, int32_t n_exprs, qgpu_plan** out);

/* Sort::new_with_limit(exprs, input, limit) (physical/plan/sort.rs:29-41,48-82; SURVEY 8f "next" #1): stable lexicographic
 * sort by the expressions with per-expression SortOptions {descending, nulls_first} (arrow SortOptions), ties keep the input
 * order (the reference appends the row index as a final key); limit >= 0 keeps the first `limit` rows (top-N), limit < 0
 * keeps all.  Always one output batch.  The planner passes nulls_first = true (planner/mod.rs:339-342). */
int qgpu_plan_sort(qgpu_ctx* ctx, qgpu_plan* input, const qgpu_expr* const* exprs, const int32_t* descending, const int32_t* nulls_first,
                   int32_t n_exprs, int64_t limit, qgpu_plan** out);
/* Limit::new(input, fetch, skip) (physical/plan/limit.rs:15-58): rows [skip, skip + fetch) of the input; fetch < 0 = no bound */
int qgpu_plan_limit(qgpu_ctx* ctx, qgpu_plan* input, int64_t fetch, int64_t skip, qgpu_plan** out);
/* HashAggregate::new(schema, input, group_exprs, aggregate_exprs) (aggregate/hash.rs:118-130);
 * n_group == 0  =>  NoGroupingAggregate::new(schema, input, aggr_expr) (aggregate/no_grouping.rs:16-22) */
int qgpu_plan_aggregate(qgpu_ctx* ctx, const struct ArrowSchema* schema, qgpu_plan* input,
                        const qgpu_expr* const* group_exprs, int32_t n_group,
                        const qgpu_agg_desc* aggs, int32_t n_aggs, qgpu_plan** out);
/* HashJoinExec::try_new(left, right, join_type, on, filter) (join/hash_join.rs:122-146).
 * Build side is always `left` (hash_join.rs:355-359).  n_on == 0 is an InternalError. */
int qgpu_plan_hash_join(qgpu_ctx* ctx, qgpu_plan* left, qgpu_plan* right, int32_t join_type,
                        const qgpu_expr* const* left_on, const qgpu_expr* const* right_on, int32_t n_on,
                        const qgpu_join_filter* filter, qgpu_plan** out);
/* NestedLoopJoinExec::try_new(left, right, join_type, filter) (join/nest_loop_join.rs:52-76): the planner's choice
 * for joins without equi-conditions (planner/mod.rs:316-320).  Matched pairs come out ordered by (right row, left
 * row); Left / Right / Full append the unmatched left rows, then the unmatched right rows. */
int qgpu_plan_nested_loop_join(qgpu_ctx* ctx, qgpu_plan* left, qgpu_plan* right, int32_t join_type,
                               const qgpu_join_filter* filter, qgpu_plan** out);
/* CrossJoin::new(left, right) (join/cross_join.rs:62-116): the cartesian product (what `FROM a, b` stays when no
 * equi-condition links a and b: optimizer/rule/eliminate_cross_join.rs).  Rows come out left-row major; the reference's
 * order additionally depends on its inputs' batch boundaries (one batch per (left batch, right batch, left row)). */
int qgpu_plan_cross_join(qgpu_ctx* ctx, qgpu_plan* left, qgpu_plan* right, qgpu_plan** out);
/* ---- exchange operators of a distributed plan (one process per GPU, qgpu_comm_init first) -----------------
 * The reference is single-process: these two nodes are what a distributed planner inserts around its operators.
 * Broadcast: every rank executes `child` over its shard; the rows of ALL ranks (rank order) are this node's output on
 * every rank -- the build side of a broadcast join.  Fixed-width columns (NULLs allowed).  order_free != 0: the consumer
 * does not depend on the child's row order (it feeds a hash table), so joins below may emit rows in any order.
 * FinalAggregate: `child` yields PARTIAL groups per rank; rows are hash-partitioned on the first key column, exchanged
 * all-to-all and re-aggregated (merge_ops[i] for value_columns[i]: 0 SUM -- partial sums and counts --, 1 MIN, 2 MAX);
 * every child column must be a key or a value; NULL-free fixed-width columns.  The result stays sharded: each final group
 * is returned by exactly one rank.  Without a communicator (or world 1) both nodes are the identity. */
int qgpu_plan_broadcast(qgpu_ctx* ctx, qgpu_plan* child, int32_t order_free, qgpu_plan** out);
/* Broadcast for the build side of an INNER equi-join whose probe side is `probe_table` on this rank, joined on
 * child column `key_column` = probe column `probe_key_column`: a build row travels only to the ranks whose probe-side key
 * range (column statistics) contains its key -- dynamic partition pruning; with range-sharded inputs every rank receives
 * ~1/world of the rows.  Implies order_free.  Falls back to the full broadcast when a range is unknown or a column has NULLs. */
int qgpu_plan_broadcast_pruned(qgpu_ctx* ctx, qgpu_plan* child, int32_t key_column, qgpu_table* probe_table,
                               int32_t probe_key_column, qgpu_plan** out);
int qgpu_plan_final_aggregate(qgpu_ctx* ctx, qgpu_plan* child, const int32_t* key_columns, int32_t n_keys,
                              const int32_t* value_columns, const int32_t* merge_ops, int32_t n_values, qgpu_plan** out);
/* PhysicalPlan::schema (physical/plan/mod.rs:26) */
int qgpu_plan_schema(const qgpu_plan* p, struct ArrowSchema* out);
/* PhysicalPlan::execute (physical/plan/mod.rs:27): runs the whole subtree on the GPU (intermediate
 * results stay in HBM) and returns the materialised Vec<RecordBatch> as a stream of host batches. */
int qgpu_plan_execute(qgpu_plan* p, struct ArrowArrayStream* out);
/* Same, but the result stays in HBM as a new table (used to chain operators and by bench.py's
 * device-resident leg).  *out_batches receives the number of RecordBatches the reference would
 * have returned (0 for HashAggregate over zero input batches). */
int qgpu_plan_execute_device(qgpu_plan* p, qgpu_table** out, int64_t* out_batches);
/* device time (ms, CUDA events on the context stream) and kernel launches of the last execute */
int qgpu_plan_last_stats(const qgpu_plan* p, double* device_ms, int64_t* launches);
/* name of the execution strategy chosen for this node by the last execute ("fused_scan_agg", ...) */
const char* qgpu_plan_strategy(const qgpu_plan* p);
void qgpu_plan_free(qgpu_plan* p);

/* ---- sharded (multi-GPU) aggregates: SURVEY 8e -----------------------------------------------------------
 * The reference has no exchange operator (single process); this is the B200 build's distribution strategy for
 * aggregates over row-range shards.  `p` must be (Projection|Filter)* <- HashAggregate/NoGroupingAggregate <- ...
 * and every process runs the same plan over its own shard:
 *   1. qgpu_plan_state_bytes(p, max_groups, &n)          size of one state block
 *   2. qgpu_plan_partial_state(p, row_offset, max_groups, buf, n)
 *        runs the aggregate on this shard up to (not including) finalisation and writes the per-group state
 *        (packed key values, first row + row_offset, 128-bit accumulator words, counts) into the DEVICE buffer
 *        `buf` (stream-ordered on qgpu_ctx_stream);
 *   3. the caller all-gathers the blocks of all shards (ncclAllGather) into one device buffer;
 *   4. qgpu_plan_execute_merged(p, gathered, n_states, max_groups, out)
 *        merges the states exactly (integer/decimal results are bit-identical to one GPU over the whole table)
 *        and finishes the plan.  More than max_groups groups on a shard -> QGPU_ERR_INTERNAL (use repartition). */
/* Hash repartition (high-cardinality group-by / join inputs): `out` = the same rows grouped by
 * partition id = mix64(key) % n_parts (order inside a partition unspecified); offsets[n_parts + 1] = row offsets of
 * the partitions.  The key must be a NULL-free integer / date / Decimal(p<=18) column; every resident column must
 * be NULL-free and fixed-width.  The per-partition slices are contiguous: the caller sends slice p of every column
 * to rank p (ncclSend/ncclRecv == torch.distributed.all_to_all_single) using qgpu_table_column_device_buffer. */
int qgpu_table_hash_partition(qgpu_table* t, int32_t key_col, int32_t n_parts, qgpu_table** out, int64_t* offsets);
/* device pointer, byte size and value width of a resident fixed-width column's value buffer (valid while `t` lives) */
int qgpu_table_column_device_buffer(qgpu_table* t, int32_t col, void** ptr, int64_t* bytes, int32_t* value_width);

/* Multi-GPU exchange of a high-cardinality group-by (SURVEY 8e "hash-repartitioned ... over NVLink"; the reference is
 * single-process, so there is no reference interface behind these).  Plan: (Projection|Filter)* <- HashAggregate <-
 * (Filter)* <- Scan with integer-like keys.  One process per GPU, every rank runs the same sequence:
 *   0. qgpu_plan_exchange_keystats -> (min, max) of every group key on this rank (stats[2k], stats[2k+1]; up to 4 keys);
 *                                     the caller reduces them over the ranks (once per table: they do not change)
 *   1. qgpu_plan_exchange_sketch   (global (min, max) per key: keys then pack to the same code on every rank)
 *                                     -> device buffer (histogram of the key hash's top byte + HyperLogLog registers);
 *                                     the caller all-gathers the blocks (NCCL) and copies them to the host;
 *                                     *eligible = 0: the keys do not pack into 64 bits (same answer on every rank)
 *   2. qgpu_plan_exchange_prepare  (gathered blocks, world, rank) -> *n_handles opaque 80-byte handles of this rank's
 *                                     receive buffers (handles_out must hold 6 of them); *eligible = 0 when the plan
 *                                     cannot run this way (the same answer on every rank) -- use qgpu_table_hash_partition
 *   3. the caller all-gathers the handles (rank-major), then
 *      qgpu_plan_exchange_scatter  partitions this rank's rows and stores every tuple straight into its owner's
 *                                     buffers (peer-to-peer over NVLink); returns when this rank's stores are issued
 *   4. the caller runs a barrier, then
 *      qgpu_plan_exchange_finish   aggregates the tuples this rank received; *overflow != 0 -> a skewed bucket did not
 *                                     fit (fall back on every rank); otherwise the next qgpu_plan_execute /
 *                                     qgpu_plan_execute_device of `p` returns this rank's share of the groups.
 * The whole result is the concatenation of the ranks' results (groups are rank-disjoint). */
int qgpu_plan_exchange_keystats(qgpu_plan* p, int64_t* stats, int32_t* n_keys);
int qgpu_plan_exchange_sketch(qgpu_plan* p, const int64_t* global_stats, void** device_buf, int64_t* bytes, int32_t* eligible);
int qgpu_plan_exchange_prepare(qgpu_plan* p, const void* gathered_host, int32_t world, int32_t rank, void* handles_out,
                               int32_t* n_handles, int32_t* eligible);
int qgpu_plan_exchange_scatter(qgpu_plan* p, const void* all_handles_host);
int qgpu_plan_exchange_finish(qgpu_plan* p, int32_t* overflow);

int qgpu_plan_state_bytes(qgpu_plan* p, int32_t max_groups, int64_t* bytes);
int qgpu_plan_partial_state(qgpu_plan* p, int64_t row_offset, int32_t max_groups, void* device_buf, int64_t cap_bytes);
/* Declares that the consumer of `p` does not depend on its output ROW ORDER (e.g. the rows feed an all-gather or a hash
 * table).  HashJoinExec's order is deterministic in the reference (hash_join.rs:474-512) and is reproduced by default; an
 * order-free Inner join with unique integer build keys runs as one probe-scan kernel instead.  The flag is set on the
 * first join below `p`'s Projection / Filter operators. */
int qgpu_plan_set_order_free(qgpu_plan* p, int32_t on);

/* like qgpu_plan_execute_merged, the result stays in HBM (cf. qgpu_plan_execute_device) */
int qgpu_plan_execute_merged_device(qgpu_plan* p, const void* gathered_device_buf, int32_t n_states, int32_t max_groups,
                                    qgpu_table** out);
int qgpu_plan_execute_merged(qgpu_plan* p, const void* gathered_device_buf, int32_t n_states, int32_t max_groups,
                             struct ArrowArrayStream* out);

/* ---- asynchronous execution ---------------------------------------------------------------------------------------
 * Like qgpu_plan_execute_device, but returns as soon as the kernels are queued on the context stream: the result table's
 * row count / NULL counts may still be on their way from the device (plans that end in the fused dense aggregate: scan
 * kernel + single-CTA epilogue, no host round trip).  qgpu_table_wait (or the first accessor that needs the numbers:
 * qgpu_table_num_rows, _export, a Scan over the table ...) waits for them; errors of the producing kernels (decimal AVG
 * overflow, sharded-merge overflow, peer time-out) surface there.  Other plans resolve before returning. */
int qgpu_plan_execute_device_async(qgpu_plan* p, qgpu_table** out);
int qgpu_table_wait(qgpu_table* t);

/* ---- communicator: multi-GPU plumbing below the ABI (no reference counterpart) ---------------------------------
 * One process per GPU.  Rank 0 calls qgpu_comm_unique_id and hands the 128 bytes to every rank (any side channel); every
 * rank then calls qgpu_comm_init, which creates an NCCL communicator owned by the library (libnccl.so.2 is dlopen'ed here:
 * single-GPU hosts never need it), allocates this rank's SYMMETRIC peer buffer and maps every peer's buffer over
 * NVLink (CUDA IPC).  Failures are QGPU_ERR_NCCL.  qgpu_comm_init_local wires several contexts of ONE process together
 * without NCCL (tests: ranks emulated on one GPU; peer kernels, no NCCL collectives). */
int qgpu_comm_unique_id(void* out, int64_t cap /* >= 128 */);
int qgpu_comm_init(qgpu_ctx* ctx, const void* unique_id, int32_t rank, int32_t world);
int qgpu_comm_init_local(qgpu_ctx** ctxs, int32_t n);
int qgpu_comm_destroy(qgpu_ctx* ctx);
int qgpu_comm_world(const qgpu_ctx* ctx, int32_t* rank, int32_t* world);
/* plain NCCL collectives on the context stream over DEVICE buffers (hosts without their own NCCL binding) */
int qgpu_comm_all_gather(qgpu_ctx* ctx, const void* send_device, void* recv_device, int64_t bytes_per_rank);
int qgpu_comm_all_to_all(qgpu_ctx* ctx, const void* send_device, const int64_t* send_offsets, const int64_t* send_bytes,
                         void* recv_device, const int64_t* recv_offsets, const int64_t* recv_bytes);
int qgpu_comm_barrier(qgpu_ctx* ctx);

/* Whole sharded aggregate step in one call (replaces partial_state + all-gather + execute_merged): every rank runs
 * `p` over its row-range shard; the shard-local aggregate is followed by ONE kernel that stores this rank's state block
 * into every peer's symmetric buffer over NVLink, waits on the peers' epoch flags, merges all blocks exactly and
 * finalises -- no collective call, no host round trip.  Every rank returns the same (whole) result.  row_offset = global
 * index of the shard's first row; more than max_groups (<= 4096) groups on a shard -> QGPU_ERR_INTERNAL (use the
 * exchange / hash repartition for high-cardinality keys); a peer that never arrives -> QGPU_ERR_NCCL after
 * QGPU_PEER_TIMEOUT_MS (default 20 s).  Every rank must issue the same sequence of sharded executions. */
int qgpu_plan_execute_sharded(qgpu_plan* p, int64_t row_offset, int32_t max_groups, struct ArrowArrayStream* out);
/* the result stays in HBM; async != 0: see qgpu_plan_execute_device_async */
int qgpu_plan_execute_sharded_device(qgpu_plan* p, int64_t row_offset, int32_t max_groups, int32_t async, qgpu_table** out);

#ifdef __cplusplus
}
#endif
#endif /* QGPU_H */
