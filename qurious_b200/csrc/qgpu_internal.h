// Internal data model of libqgpu.so (not part of the ABI; the ABI is include/qgpu.h).
//
// HBM layout (DESIGN.md "Data layout"):
//   * a table is a set of immutable, contiguous, 256B-aligned device columns (Arrow layout:
//     values / i32 offsets + bytes / bit-packed validity); Decimal128(p<=18) columns whose values
//     are proven to fit are stored NARROWED to int64 (8 B/value) -- the kernels read that layout;
//   * operators exchange `View`s: lazy columns = (base column, optional int64 row-index vector,
//     -1 = NULL).  Filters and joins only produce index vectors; payload columns are gathered
//     once, when something finally needs them (late materialisation).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <functional>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/qgpu.h"

namespace qgpu {

typedef __int128 i128;
typedef unsigned __int128 u128;

// ----------------------------------------------------------------------------------------------
// errors
// ----------------------------------------------------------------------------------------------
struct QError : public std::runtime_error {
  int code;
  QError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
[[noreturn]] inline void throw_internal(const std::string& m) { throw QError(QGPU_ERR_INTERNAL, "InternalError: " + m); }
[[noreturn]] inline void throw_arrow(const std::string& m) { throw QError(QGPU_ERR_ARROW, "ArrowError: " + m); }

#define CUDA_CHECK(expr)                                                                          \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      int _c = (_e == cudaErrorMemoryAllocation) ? QGPU_ERR_OOM : QGPU_ERR_CUDA;                  \
      throw ::qgpu::QError(_c, std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " +    \
                                   __FILE__ + ":" + std::to_string(__LINE__));                    \
    }                                                                                             \
  } while (0)

// ----------------------------------------------------------------------------------------------
// logical types
// ----------------------------------------------------------------------------------------------
struct DType {
  uint8_t id = QGPU_T_NULL;
  uint8_t precision = 0;
  int8_t scale = 0;
  bool operator==(const DType& o) const { return id == o.id && precision == o.precision && scale == o.scale; }
  bool operator!=(const DType& o) const { return !(*this == o); }
  bool is_decimal() const { return id == QGPU_T_DECIMAL128; }
  bool is_signed_int() const { return id >= QGPU_T_INT8 && id <= QGPU_T_INT64; }
  bool is_unsigned_int() const { return id >= QGPU_T_UINT8 && id <= QGPU_T_UINT64; }
  bool is_int() const { return is_signed_int() || is_unsigned_int(); }
  bool is_float() const { return id == QGPU_T_FLOAT32 || id == QGPU_T_FLOAT64; }
  bool is_date() const { return id == QGPU_T_DATE32 || id == QGPU_T_DATE64; }
  bool is_time() const { return id == QGPU_T_TIME32 || id == QGPU_T_TIME64; }  // unit in `scale`: 0 s, 1 ms, 2 us, 3 ns
  std::string str() const;
};
inline DType mk_type(int id, int p = 0, int s = 0) {
  DType t;
  t.id = (uint8_t)id;
  t.precision = (uint8_t)p;
  t.scale = (int8_t)s;
  return t;
}
// byte width of one value in the canonical (Arrow) layout; 0 for bool (bit-packed) / utf8 / null
int arrow_width(const DType& t);

struct Field {
  std::string name;
  DType type;
  bool nullable = true;
  std::string metadata;  // raw Arrow-encoded metadata blob (may be empty)
};
struct Schema {
  std::vector<Field> fields;
  std::string metadata;  // raw Arrow-encoded metadata blob
};

// ----------------------------------------------------------------------------------------------
// device memory
// ----------------------------------------------------------------------------------------------
struct Ctx;
struct Comm;         // comm.h: NCCL communicator + symmetric peer buffers (multi-GPU)
struct MetaSlot;     // pending.h
struct DBuf {
  Ctx* ctx = nullptr;
  void* ptr = nullptr;
  size_t bytes = 0;
  std::shared_ptr<DBuf> parent;  // sub-buffer of a slab (Slab::take): the memory belongs to `parent`
  cudaStream_t free_stream = nullptr;  // freed in this stream's order instead of ctx->stream (buffers last used on the epilogue stream)
  DBuf(Ctx* c, size_t n);
  DBuf(std::shared_ptr<DBuf> slab, void* p, size_t n) : ctx(slab->ctx), ptr(p), bytes(n), parent(std::move(slab)) {}
  ~DBuf();
  DBuf(const DBuf&) = delete;
  DBuf& operator=(const DBuf&) = delete;
};
typedef std::shared_ptr<DBuf> DBufP;

// Many small temporaries of one operator from ONE allocation (and one memset): a re-executed Q1 step issued ~100
// cudaMallocAsync / cudaFreeAsync / cudaMemsetAsync calls for buffers of a few KB each.
struct Slab {
  DBufP buf;
  size_t off = 0;
  Slab(Ctx* ctx, size_t total_bytes, bool zero);
  static size_t need(size_t bytes) { return ((bytes + 255) / 256) * 256 + 256; }  // what take(bytes) consumes
  DBufP take(size_t bytes);
};

// physical storage class of a device column
enum Phys : uint8_t {
  PH_NULL = 0,  // no buffers (all NULL)
  PH_BIT,       // bit-packed booleans
  PH_I8, PH_I16, PH_I32, PH_I64,
  PH_U8, PH_U16, PH_U32, PH_U64,
  PH_F32, PH_F64,
  PH_I128,      // Decimal128 in Arrow layout (16 B)
  PH_D64,       // Decimal128 narrowed to int64 (8 B) -- value fits, proven at ingest
  PH_STR        // utf8: i32 offsets + bytes
};
inline int phys_width(Phys p) {
  switch (p) {
    case PH_I8: case PH_U8: return 1;
    case PH_I16: case PH_U16: return 2;
    case PH_I32: case PH_U32: case PH_F32: return 4;
    case PH_I64: case PH_U64: case PH_F64: case PH_D64: return 8;
    case PH_I128: return 16;
    default: return 0;
  }
}

struct DCol {
  DType type;
  Phys phys = PH_NULL;
  int64_t length = 0;
  int64_t null_count = 0;
  DBufP data;      // values (or utf8 bytes)
  DBufP offsets;   // utf8: (length+1) int32
  DBufP validity;  // bitmap (uint32 words), null when null_count == 0
  int64_t str_bytes = 0;
  bool str_bytes_is_bound = false;  // str_bytes >= offsets[length] (small gathers size the data buffer without a round trip)
  mutable int32_t max_str_len = -1;  // longest value of a Utf8 column, computed on first use (take_column)
  // lazily computed value range of the non-null values (ints / dates / decimals)
  bool has_stats = false;
  i128 vmin = 0, vmax = 0;
  // lazily built dictionary encoding of a low-cardinality Utf8 column (fused.cu: ensure_dict):
  // dict_codes[row] = index into dict_values; dict_state: 0 = not tried, 1 = encoded, -1 = too many values
  int dict_state = 0;
  DBufP dict_codes;  // uint8 per row
  std::vector<std::string> dict_values;
  int64_t bytes_resident() const {
    int64_t b = 0;
    if (data) b += (int64_t)data->bytes;
    if (offsets) b += (int64_t)offsets->bytes;
    if (validity) b += (int64_t)validity->bytes;
    return b;
  }
};
typedef std::shared_ptr<DCol> DColP;

// int64 row-index vector; -1 = NULL row (outer joins)
struct IdxVec {
  DBufP buf;
  int64_t length = 0;
  bool may_have_null = false;
  const int64_t* ptr() const { return buf ? (const int64_t*)buf->ptr : nullptr; }
};
typedef std::shared_ptr<IdxVec> IdxP;

struct LazyCol {
  DColP base;  // may be null for a column that was never uploaded (selective upload)
  IdxP idx;    // null = identity
};

// Result metadata that is still on its way from the device (row count, NULL counts, error code of the producing
// kernels): the single-CTA epilogue of the dense fused path writes a small block that is copied to a pinned slot
// behind the kernels; whoever needs the numbers calls resolve() -- one event wait, no stream synchronisation.  Until
// then View::num_rows / DCol::length are UPPER BOUNDS (buffers are sized for them).
struct MetaSlot {
  unsigned long long* host = nullptr;  // pinned, META_WORDS words
  cudaEvent_t ev = nullptr;
  bool busy = false;
};
constexpr int META_WORDS = 128;
struct Pending {
  Ctx* ctx = nullptr;
  MetaSlot* slot = nullptr;
  bool done = false;
  int err_code = 0;
  std::string err_msg;
  int64_t num_rows = 0;
  // patches the result columns from the metadata words and sets num_rows; may throw QError
  std::function<void(const unsigned long long*, Pending&)> apply;
  std::shared_ptr<Pending> also;  // resolved first (deferred verification of speculated counts)
  void resolve();
  ~Pending();
};
typedef std::shared_ptr<Pending> PendingP;
// queues the copy of `words` metadata words at `dev_meta` into a pinned slot on ctx->stream (+ an event)
PendingP make_pending(Ctx* ctx, const void* dev_meta, int words, std::function<void(const unsigned long long*, Pending&)> apply,
                      cudaStream_t stream = nullptr /* default: the compute stream */);

struct View {
  Schema schema;
  std::vector<LazyCol> cols;
  int64_t num_rows = 0;
  int64_t num_batches = 1;  // how many RecordBatches the reference would have returned
  PendingP pending;         // non-null: num_rows and the columns' lengths / NULL counts are bounds until resolve()
  void resolve() {
    if (!pending) return;
    PendingP p = pending;
    pending.reset();
    p->resolve();
    num_rows = p->num_rows;
  }
};

// ----------------------------------------------------------------------------------------------
// context
// ----------------------------------------------------------------------------------------------
struct Ctx {
  int device = 0;
  cudaStream_t stream = nullptr;       // compute stream
  cudaStream_t copy_stream = nullptr;  // H2D staging side stream
  // The single-CTA epilogues (epilogue.cu: finalisation / multi-GPU exchange + merge of the dense fused aggregate) run
  // here, behind an event of the compute stream: the next execution's scan kernel overlaps them.
  cudaStream_t epi_stream = nullptr;
  cudaEvent_t epi_ready = nullptr;
  cudaMemPool_t pool = nullptr;
  // The stream-ordered pool keeps at least this much reserved (grown at init, kept by qgpu_release_cached_memory): a plan
  // that allocates and frees hundreds of MB per execution otherwise keeps sending the pool back to the driver for physical
  // memory in the middle of a step (distributed Q3 at N = 2: median step 1.83 -> 1.40 ms, 5-38 ms outliers gone).
  // qgpu_set_option "pool_reserve_mb" / QGPU_POOL_RESERVE_MB.
  size_t pool_floor_bytes = (size_t)4 << 30;
  void reserve_pool();
  std::recursive_mutex mu;
  std::string last_error;
  int64_t launches = 0;
  // qgpu_counter: bytes of every device allocation (all temporaries: index vectors, gathered columns, tables) and the
  // payload bytes written by column gathers (take_column) since the context was created
  int64_t alloc_bytes_total = 0, gather_bytes_total = 0;
  int sm_count = 148;
  bool compat_avg_precision = false;
  bool compat_empty_decimal_sum = false;
  // pinned staging ring for pageable host buffers
  static const int kStageSlots = 2;
  size_t stage_bytes = (size_t)32 << 20;
  void* stage[kStageSlots] = {nullptr, nullptr};
  cudaEvent_t stage_ev[kStageSlots] = {nullptr, nullptr};
  int stage_next = 0;
  // staged multi-batch ingest (ingest.cu): the ring above cut into 2 slots per host worker thread
  size_t ingest_slot_bytes = (size_t)2 << 20;
  int ingest_threads = 0;       // 0: min(hardware threads, 16) (QGPU_INGEST_THREADS); qgpu_set_option "ingest_threads"
  int ingest_host_narrow = -1;  // -1: on (QGPU_INGEST_HOST_NARROW); qgpu_set_option "ingest_host_narrow"
  std::vector<cudaEvent_t> ingest_ev;
  // small pinned scratch for D2H of scalars / flags
  void* pinned_scratch = nullptr;
  size_t pinned_scratch_bytes = 1 << 16;

  // per-kernel CUDA-event profiling (off by default; qgpu_profile_enable)
  bool profiling = false;
  long long prof_min_blocks = 0;  // qgpu_profile_enable(ctx, n >= 2): only launches of >= n blocks are bracketed
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
  std::vector<const char*> prof_names;
  int prof_used = 0;
  std::string prof_text;
  int prof_begin(const char* name);
  void prof_end(int slot);
  std::string prof_report();  // "name\tlaunches\ttotal_ms\tmax_ms\n" per kernel; resets the log

  unsigned long long exec_epoch = 0;  // one per C-ABI call (guard): exchange nodes memoise their output within a call
  bool async_ok = false;       // inside qgpu_plan_execute_device_async: operators may leave result metadata pending
  size_t epi_smem_set = 0;     // epilogue.cu: dynamic shared memory opt-in done on this device up to this size
  std::shared_ptr<Comm> comm;  // set by qgpu_comm_init / qgpu_comm_init_local
  // pinned result-metadata slots of asynchronous executions (pending.h)
  std::vector<MetaSlot*> meta_slots;

  // QGPU_TRACE=1: synchronising wall-clock marks printed to stderr (debug aid, never on in benchmarks)
  void trace(const char* what);
  double trace_t0 = 0;
  int trace_on = -1;

  DBufP alloc(size_t bytes);
  // exact-size free list of large device blocks (kernels.cu: DBuf)
  std::vector<std::pair<size_t, void*>> big_free;
  size_t big_free_bytes = 0;
  void release_big_blocks();
  DBufP alloc_zero(size_t bytes);
  void h2d(void* dst, const void* src, size_t bytes);        // pageable or pinned host -> device
  void d2h_sync(void* dst, const void* src, size_t bytes);   // device -> host, waits
  void sync();
  template <typename T>
  T read_scalar(const T* dptr) {
    T v;
    d2h_sync(&v, dptr, sizeof(T));
    return v;
  }
  // A device-side COUNT (selection size, join output size, group count, duplicate / error flag) the host needs to size
  // the next buffers and launches.  Outside a speculation scope this is read_scalar (one host round trip).  Inside one
  // (Speculation, below) the first execution of a cached plan learns the values; re-executions over the same immutable
  // tables take the learned value immediately, queue a copy of the actual device word for later verification and keep
  // the stream busy -- the whole pipeline then synchronises once, at its end.
  template <typename T>
  T read_count(const T* dptr, const char* site);  // site: names the call site; a replay that reaches another one re-learns
  struct Speculation* spec = nullptr;
};

struct Speculation {
  static const int kMax = 96;
  bool have = false;    // `learned` holds the counts of a complete earlier execution
  bool replay = false;  // this execution takes learned counts and verifies them at the end
  std::vector<unsigned long long> learned;  // raw bits (zero-extended), in call order
  std::vector<const char*> sites;           // call site of every learned count
  size_t cursor = 0;
  DBufP check;          // replay: the actual device words, one per read
};
struct SpeculationMiss {};  // a replay reached a count it did not learn (other call site / more reads): re-run in learning mode
// Work whose result is CACHED across executions (plan analysis, dictionary encoding, column statistics) runs outside the
// speculation: its counts are read once, not once per execution
struct SpecSuspend {
  Ctx* ctx;
  Speculation* saved;
  explicit SpecSuspend(Ctx* c) : ctx(c), saved(c->spec) { c->spec = nullptr; }
  ~SpecSuspend() { ctx->spec = saved; }
};

template <typename T>
T Ctx::read_count(const T* dptr, const char* site) {
  static_assert(sizeof(T) <= 8, "counts are at most 64 bits");
  Speculation* sp = spec;
  if (!sp) return read_scalar(dptr);
  if (!sp->replay) {
    const T v = read_scalar(dptr);
    unsigned long long bits = 0;
    memcpy(&bits, &v, sizeof(T));
    sp->learned.push_back(bits);
    sp->sites.push_back(site);
    return v;
  }
  if (sp->cursor >= sp->learned.size() || sp->cursor >= (size_t)Speculation::kMax || sp->sites[sp->cursor] != site) throw SpeculationMiss();
  CUDA_CHECK(cudaMemcpyAsync((char*)sp->check->ptr + 8 * sp->cursor, dptr, sizeof(T), cudaMemcpyDeviceToDevice, stream));
  T v;
  memcpy(&v, &sp->learned[sp->cursor], sizeof(T));
  sp->cursor++;
  return v;
}

// RAII: runs the enclosed host code under `sp` (learning or replaying); verify() after the pipeline's final
// synchronisation tells whether every speculated count was right
struct SpecScope {
  Ctx* ctx;
  Speculation* sp;
  Speculation* saved;
  SpecScope(Ctx* c, Speculation* s);
  ~SpecScope() { ctx->spec = saved; }
  bool verify();
  // replay only: the comparison of the actual counts with the speculated ones rides with `p` (resolved before it); a
  // mismatch is an InternalError there.  false: the replay already diverged on the host (read fewer / more counts)
  bool defer_verify(const PendingP& p);
};

// ----------------------------------------------------------------------------------------------
// table (HBM-resident MemoryTable)
// ----------------------------------------------------------------------------------------------
struct TableChunk {
  std::vector<DColP> cols;  // null entries for columns that were not uploaded
  int64_t rows = 0;
};
}  // namespace qgpu

namespace qgpu {
struct TableImpl {
  Ctx* ctx = nullptr;
  Schema schema;
  std::vector<TableChunk> chunks;  // appended batches, consolidated lazily
  // host batches appended but not uploaded yet (owned: released by flush_pending / the destructor); they follow `chunks`
  // in row order and share one upload-column set
  std::vector<ArrowArray> pending_host;
  std::vector<char> pending_want;
  void flush_pending();
  ~TableImpl();
  std::vector<DColP> cols;         // consolidated columns (one per schema field; null = not uploaded)
  bool consolidated = true;
  int64_t num_rows = 0;
  int64_t num_batches = 0;
  PendingP pending;  // result table of an asynchronous execute: see View::pending
  void resolve() {
    if (!pending) return;
    PendingP p = pending;
    pending.reset();
    p->resolve();
    num_rows = p->num_rows;
  }
  void consolidate();
};
}  // namespace qgpu

struct qgpu_ctx {
  qgpu::Ctx c;
};
struct qgpu_table {
  std::shared_ptr<qgpu::TableImpl> t;
};

namespace qgpu {

// ---- arrow_io.cu -------------------------------------------------------------------------------
Schema import_schema(const ArrowSchema* s);
void export_schema(const Schema& s, ArrowSchema* out);
TableChunk import_batch(Ctx* ctx, const Schema& schema, ArrowArray* batch, const int32_t* upload_columns,
                        int32_t n_upload, bool device_resident);
// ---- csv.cu ------------------------------------------------------------------------------------
TableChunk parse_csv_device(Ctx* ctx, const Schema& schema, const std::vector<char>& want, const unsigned char* text, int64_t len,
                            const qgpu_csv_options& opt);
DBufP read_file_to_device(Ctx* ctx, const char* path, int64_t* len_out);
// ---- ingest.cu ---------------------------------------------------------------------------------
int ingest_worker_threads(Ctx* ctx);
int comm_world_size(Ctx* ctx);  // comm.cu: 0 without a communicator
void validate_host_batch(const Schema& schema, const ArrowArray* batch, const std::vector<char>& want);
TableChunk import_host_batches(Ctx* ctx, const Schema& schema, const std::vector<ArrowArray>& batches, const std::vector<char>& want,
                               int host_narrow = -1);
// materialised columns -> host ArrowArray (struct); blocks until the copy is done
void export_batch(Ctx* ctx, const Schema& schema, const std::vector<DColP>& cols, int64_t num_rows,
                  ArrowArray* out);
void make_stream(const Schema& schema, std::vector<ArrowArray>&& batches, ArrowArrayStream* out);
std::string merge_metadata(const std::string& base, const std::string& key, const std::string& value);
bool metadata_get(const std::string& blob, const std::string& key, std::string* value);

}  // namespace qgpu
