// Staged ingest: MANY host RecordBatches -> ONE device chunk (replaces MemoryTable::insert, datasource/memory.rs:104-111,
// at the reference's native granularity: csv.rs:34-72 reads 1024-row batches, so TPC-H SF10 arrives as ~58 k batches).
//
// qgpu_table_append only validates and RETAINS a host batch; the upload happens once, when the table is first used
// (TableImpl::consolidate -> flush_pending) or on qgpu_table_flush:
//   * every uploaded column gets its final, contiguous device buffer up front (no per-batch allocation, no D2D
//     concatenation, no per-batch kernel, no per-batch synchronisation);
//   * the destination is cut into tasks of one pinned ring slot each; host worker threads GATHER the rows of a task out of
//     however many small batches it spans into their slot -- plain copy, Decimal128(p <= 18) narrowed to the int64 the
//     kernels read (16 -> 8 B over PCIe; min / max statistics and the fits-int64 proof come out of the same loop), Utf8
//     offsets rebased, validity / boolean bits re-packed at the destination bit offset -- and queue ONE cudaMemcpyAsync per
//     slot on the copy stream (two slots per worker: the DMA of one overlaps the fill of the other);
//   * large page-locked sources that need no transformation skip the ring: one direct DMA per buffer.
// One synchronisation at the end of the flush.
#include <atomic>
#include <algorithm>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <mutex>
#include <thread>

#include "kernels.h"
#include "plan.h"

namespace qgpu {

namespace {

struct Part {  // one column of one host batch
  const uint8_t* validity = nullptr;
  const char* values = nullptr;      // fixed-width values / boolean bits
  const int32_t* offsets = nullptr;  // utf8
  const char* data = nullptr;        // utf8 bytes
  int64_t off = 0;                   // first element: child offset + struct offset
  int64_t n = 0;
  int64_t nulls = 0;
  int64_t row0 = 0;   // first destination row
  int64_t byte0 = 0;  // utf8: first destination byte
  int64_t o0 = 0;     // utf8: source offset of the first value
  int64_t sbytes = 0; // utf8: bytes of this part
};

enum TaskKind { TK_FIXED, TK_NARROW, TK_OFFSETS, TK_STRDATA, TK_VALIDITY, TK_BOOLBITS };

struct ColPlan {
  int field = 0;
  std::vector<Part> parts;  // non-empty parts only, ascending row0
  int w = 0;                // source value width (fixed-width columns)
  bool narrow = false;
  int64_t rows = 0, str_bytes = 0, nulls = 0;
  DColP col;
  std::mutex mu;  // narrowing results
  int64_t mn = INT64_MAX, mx = INT64_MIN;
  bool fits = true;
};

struct Task {
  ColPlan* cp;
  TaskKind kind;
  int64_t u0, u1;  // rows (FIXED / NARROW / OFFSETS: elements), bytes (STRDATA), 32-bit words (VALIDITY / BOOLBITS)
};

inline bool get_bit(const uint8_t* b, int64_t i) { return (b[i >> 3] >> (i & 7)) & 1; }

// dst bits [dbit, dbit+n) = src bits [sbit, sbit+n); dst was zeroed
void copy_bits_host(uint8_t* dst, int64_t dbit, const uint8_t* src, int64_t sbit, int64_t n) {
  while (n > 0 && (dbit & 7)) {  // head: up to the next destination byte boundary
    if (get_bit(src, sbit)) dst[dbit >> 3] |= (uint8_t)(1u << (dbit & 7));
    ++dbit, ++sbit, --n;
  }
  if (n <= 0) return;
  const int64_t nb = n >> 3;
  uint8_t* d = dst + (dbit >> 3);
  const uint8_t* s = src + (sbit >> 3);
  const int sh = (int)(sbit & 7);
  if (sh == 0) {
    memcpy(d, s, (size_t)nb);
  } else {
    for (int64_t i = 0; i < nb; ++i) d[i] = (uint8_t)((s[i] >> sh) | (s[i + 1] << (8 - sh)));
  }
  dbit += nb * 8, sbit += nb * 8, n -= nb * 8;
  while (n > 0) {
    if (get_bit(src, sbit)) dst[dbit >> 3] |= (uint8_t)(1u << (dbit & 7));
    ++dbit, ++sbit, --n;
  }
}
void set_bits_host(uint8_t* dst, int64_t dbit, int64_t n) {
  while (n > 0 && (dbit & 7)) {
    dst[dbit >> 3] |= (uint8_t)(1u << (dbit & 7));
    ++dbit, --n;
  }
  if (n <= 0) return;
  memset(dst + (dbit >> 3), 0xff, (size_t)(n >> 3));
  dbit += (n >> 3) * 8;
  n &= 7;
  while (n > 0) {
    dst[dbit >> 3] |= (uint8_t)(1u << (dbit & 7));
    ++dbit, --n;
  }
}
int64_t count_zero_bits_host(const uint8_t* bits, int64_t off, int64_t n) {
  int64_t ones = 0, i = 0;
  while (i < n && ((off + i) & 7)) ones += get_bit(bits, off + i), ++i;
  for (; i + 64 <= n; i += 64) {
    uint64_t v;
    memcpy(&v, bits + ((off + i) >> 3), 8);
    ones += __builtin_popcountll(v);
  }
  for (; i < n; ++i) ones += get_bit(bits, off + i);
  return n - ones;
}

// index of the part holding destination row r (parts ascending by row0, all non-empty)
size_t part_of_row(const std::vector<Part>& ps, int64_t r) {
  size_t lo = 0, hi = ps.size();
  while (hi - lo > 1) {
    size_t mid = (lo + hi) >> 1;
    if (ps[mid].row0 <= r) lo = mid;
    else hi = mid;
  }
  return lo;
}

// fills `slot` with the task's destination bytes; returns (device destination, byte count)
std::pair<char*, size_t> fill_task(const Task& t, char* slot) {
  ColPlan& cp = *t.cp;
  const std::vector<Part>& ps = cp.parts;
  switch (t.kind) {
    case TK_FIXED: {
      const int w = cp.w;
      for (size_t i = part_of_row(ps, t.u0); i < ps.size() && ps[i].row0 < t.u1; ++i) {
        const Part& p = ps[i];
        const int64_t a = std::max(t.u0, p.row0), b = std::min(t.u1, p.row0 + p.n);
        if (b <= a) continue;
        if (p.values) memcpy(slot + (a - t.u0) * w, p.values + (p.off + a - p.row0) * w, (size_t)(b - a) * w);
        else memset(slot + (a - t.u0) * w, 0, (size_t)(b - a) * w);
      }
      return {(char*)cp.col->data->ptr + t.u0 * w, (size_t)(t.u1 - t.u0) * w};
    }
    case TK_NARROW: {
      int64_t mn = INT64_MAX, mx = INT64_MIN;
      uint64_t bad = 0;
      int64_t* out = (int64_t*)slot;
      for (size_t i = part_of_row(ps, t.u0); i < ps.size() && ps[i].row0 < t.u1; ++i) {
        const Part& p = ps[i];
        const int64_t a = std::max(t.u0, p.row0), b = std::min(t.u1, p.row0 + p.n);
        if (b <= a) continue;
        int64_t* d = out + (a - t.u0);
        const int64_t m = b - a;
        if (!p.values) {
          memset(d, 0, (size_t)m * 8);
          continue;
        }
        const uint64_t* s = (const uint64_t*)p.values + 2 * (p.off + a - p.row0);
        if (p.nulls == 0 || !p.validity) {
          for (int64_t k = 0; k < m; ++k) {
            const int64_t lo = (int64_t)s[2 * k];
            bad |= s[2 * k + 1] ^ (uint64_t)(lo >> 63);
            mn = lo < mn ? lo : mn;
            mx = lo > mx ? lo : mx;
            d[k] = lo;
          }
        } else {
          const int64_t vb = p.off + a - p.row0;
          for (int64_t k = 0; k < m; ++k) {
            int64_t lo = (int64_t)s[2 * k];
            if (get_bit(p.validity, vb + k)) {
              bad |= s[2 * k + 1] ^ (uint64_t)(lo >> 63);
              mn = lo < mn ? lo : mn;
              mx = lo > mx ? lo : mx;
            } else {
              lo = 0;  // like k_narrow: NULL slots hold 0
            }
            d[k] = lo;
          }
        }
      }
      {
        std::lock_guard<std::mutex> lk(cp.mu);
        if (bad) cp.fits = false;
        cp.mn = std::min(cp.mn, mn);
        cp.mx = std::max(cp.mx, mx);
      }
      return {(char*)cp.col->data->ptr + t.u0 * 8, (size_t)(t.u1 - t.u0) * 8};
    }
    case TK_OFFSETS: {
      int32_t* out = (int32_t*)slot;
      int64_t e1 = std::min(t.u1, cp.rows);  // element cp.rows is the grand total
      if (t.u0 < e1)
        for (size_t i = part_of_row(ps, t.u0); i < ps.size() && ps[i].row0 < e1; ++i) {
          const Part& p = ps[i];
          const int64_t a = std::max(t.u0, p.row0), b = std::min(e1, p.row0 + p.n);
          if (b <= a) continue;
          int32_t* d = out + (a - t.u0);
          const int64_t add = p.byte0 - p.o0;
          if (p.offsets) {
            const int32_t* s = p.offsets + p.off + (a - p.row0);
            for (int64_t k = 0; k < b - a; ++k) d[k] = (int32_t)(s[k] + add);
          } else {
            for (int64_t k = 0; k < b - a; ++k) d[k] = (int32_t)p.byte0;
          }
        }
      if (t.u1 > cp.rows) out[cp.rows - t.u0] = (int32_t)cp.str_bytes;
      return {(char*)cp.col->offsets->ptr + t.u0 * 4, (size_t)(t.u1 - t.u0) * 4};
    }
    case TK_STRDATA: {
      // parts ascending by byte0 as well; find the first part whose byte range reaches u0
      size_t lo = 0, hi = ps.size();
      while (hi - lo > 1) {
        size_t mid = (lo + hi) >> 1;
        if (ps[mid].byte0 <= t.u0) lo = mid;
        else hi = mid;
      }
      for (size_t i = lo; i < ps.size() && ps[i].byte0 < t.u1; ++i) {
        const Part& p = ps[i];
        const int64_t a = std::max(t.u0, p.byte0), b = std::min(t.u1, p.byte0 + p.sbytes);
        if (b <= a) continue;
        memcpy(slot + (a - t.u0), p.data + p.o0 + (a - p.byte0), (size_t)(b - a));
      }
      return {(char*)cp.col->data->ptr + t.u0, (size_t)(t.u1 - t.u0)};
    }
    case TK_VALIDITY:
    case TK_BOOLBITS: {
      const int64_t r0 = t.u0 * 32, r1 = std::min(t.u1 * 32, cp.rows);
      memset(slot, 0, (size_t)(t.u1 - t.u0) * 4);
      for (size_t i = part_of_row(ps, r0); i < ps.size() && ps[i].row0 < r1; ++i) {
        const Part& p = ps[i];
        const int64_t a = std::max(r0, p.row0), b = std::min(r1, p.row0 + p.n);
        if (b <= a) continue;
        if (t.kind == TK_VALIDITY) {
          if (!p.validity || p.nulls == 0) set_bits_host((uint8_t*)slot, a - r0, b - a);
          else copy_bits_host((uint8_t*)slot, a - r0, p.validity, p.off + a - p.row0, b - a);
        } else if (p.values) {
          copy_bits_host((uint8_t*)slot, a - r0, (const uint8_t*)p.values, p.off + a - p.row0, b - a);
        }
      }
      DBuf* dst = t.kind == TK_VALIDITY ? cp.col->validity.get() : cp.col->data.get();
      return {(char*)dst->ptr + t.u0 * 4, (size_t)(t.u1 - t.u0) * 4};
    }
  }
  return {nullptr, 0};
}

bool is_pinned(const void* p) {
  if (!p) return false;
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, p) == cudaSuccess) return attr.type == cudaMemoryTypeHost;
  cudaGetLastError();
  return false;
}

int env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  return (s && *s) ? atoi(s) : dflt;
}

Phys phys_of(const DType& t) {
  switch (t.id) {
    case QGPU_T_BOOL: return PH_BIT;
    case QGPU_T_INT8: return PH_I8;
    case QGPU_T_INT16: return PH_I16;
    case QGPU_T_INT32: case QGPU_T_DATE32: case QGPU_T_TIME32: return PH_I32;
    case QGPU_T_INT64: case QGPU_T_DATE64: case QGPU_T_TIME64: return PH_I64;
    case QGPU_T_UINT8: return PH_U8;
    case QGPU_T_UINT16: return PH_U16;
    case QGPU_T_UINT32: return PH_U32;
    case QGPU_T_UINT64: return PH_U64;
    case QGPU_T_FLOAT32: return PH_F32;
    case QGPU_T_FLOAT64: return PH_F64;
    case QGPU_T_DECIMAL128: return PH_I128;
    case QGPU_T_UTF8: return PH_STR;
    default: return PH_NULL;
  }
}

}  // namespace

// Host worker threads of one upload: qgpu_set_option / QGPU_INGEST_THREADS, else min(16, host threads / ranks on this
// node) -- eight processes uploading at once must share the cores (16 workers each on a 2 x 64-thread host took 446 ms
// per rank where 204 ms were measured without any host work); the node-local rank count is LOCAL_WORLD_SIZE (torchrun,
// mpirun) or the communicator's world size
int ingest_worker_threads(Ctx* ctx) {
  int threads = ctx->ingest_threads > 0 ? ctx->ingest_threads : env_int("QGPU_INGEST_THREADS", 0);
  if (threads > 0) return std::min(threads, 16);
  int ranks = env_int("LOCAL_WORLD_SIZE", 0);
  if (ranks <= 0 && ctx->comm) ranks = comm_world_size(ctx);
  if (ranks <= 0) ranks = 1;
  const int hw = (int)std::max(1u, std::thread::hardware_concurrency());
  return std::max(2, std::min(16, hw / ranks));
}

// Cheap structural checks at append time: a failing batch must be reported by qgpu_table_append itself and leave the
// table untouched (the expensive part, the upload, is deferred).
void validate_host_batch(const Schema& schema, const ArrowArray* batch, const std::vector<char>& want) {
  if (!batch) throw_internal("null batch");
  if (batch->n_children != (int64_t)schema.fields.size())
    throw_arrow("RecordBatch has " + std::to_string(batch->n_children) + " columns but the table schema has " +
                std::to_string(schema.fields.size()));
  for (size_t f = 0; f < schema.fields.size(); ++f) {
    if (!want[f]) continue;
    const ArrowArray* a = batch->children[f];
    const DType& ty = schema.fields[f].type;
    if (!a) throw_arrow("missing child array");
    if (ty.id == QGPU_T_NULL) continue;
    // Arrow: child.length >= struct.offset + struct.length
    if (a->length < batch->offset + batch->length) throw_arrow("child array shorter than the record batch");
    const int64_t need = ty.id == QGPU_T_UTF8 ? 3 : 2;
    if (a->n_buffers < need) throw_arrow("column '" + schema.fields[f].name + "' has too few buffers for its type");
    if (batch->length > 0) {
      if (ty.id == QGPU_T_UTF8) {
        if (!a->buffers[1]) throw_arrow("Utf8 column '" + schema.fields[f].name + "' has no offsets buffer");
      } else if (!a->buffers[1] && a->null_count != a->length) {
        throw_arrow("column '" + schema.fields[f].name + "' has no values buffer");
      }
    }
  }
}

static void run_tasks(Ctx* ctx, std::vector<Task>& tasks, size_t total_bytes) {
  if (tasks.empty()) return;
  // the ring: the context's 2 x 32 MiB pinned staging area cut into 2 slots per worker
  const size_t ring_bytes = ctx->stage_bytes * Ctx::kStageSlots;
  int threads = ingest_worker_threads(ctx);
  if (total_bytes < ((size_t)4 << 20)) threads = 1;  // small tables: the calling thread alone
  threads = (int)std::min<size_t>((size_t)threads, tasks.size());
  const size_t slot_bytes = ctx->ingest_slot_bytes;  // tasks were cut for this size; 16 workers x 2 slots x 2 MiB = the ring
  threads = (int)std::min<size_t>((size_t)threads, ring_bytes / (2 * slot_bytes));
  while ((int)ctx->ingest_ev.size() < threads * 2) {
    cudaEvent_t e;
    CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ctx->ingest_ev.push_back(e);
  }
  // Ctx::h2d's ring uses the same memory: wait for its last copies
  for (int i = 0; i < Ctx::kStageSlots; ++i) CUDA_CHECK(cudaEventSynchronize(ctx->stage_ev[i]));
  std::atomic<size_t> next(0);
  std::mutex err_mu;
  std::string err;
  int err_code = 0;
  char* ring = (char*)ctx->stage[0];
  char* ring2 = (char*)ctx->stage[1];
  const size_t per_half = ctx->stage_bytes / slot_bytes;
  auto slot_ptr = [&](int s) { return (size_t)s < per_half ? ring + (size_t)s * slot_bytes : ring2 + ((size_t)s - per_half) * slot_bytes; };
  auto worker = [&](int w) {
    try {
      CUDA_CHECK(cudaSetDevice(ctx->device));
      int cur = 0;
      bool used[2] = {false, false};
      for (;;) {
        const size_t i = next.fetch_add(1);
        if (i >= tasks.size()) break;
        const int s = w * 2 + cur;
        if (used[cur]) CUDA_CHECK(cudaEventSynchronize(ctx->ingest_ev[s]));  // the slot's previous DMA has left it
        char* slot = slot_ptr(s);
        auto dst = fill_task(tasks[i], slot);
        if (dst.second) {
          CUDA_CHECK(cudaMemcpyAsync(dst.first, slot, dst.second, cudaMemcpyHostToDevice, ctx->copy_stream));
          CUDA_CHECK(cudaEventRecord(ctx->ingest_ev[s], ctx->copy_stream));
          used[cur] = true;
        }
        cur ^= 1;
        {
          std::lock_guard<std::mutex> lk(err_mu);
          if (err_code) break;
        }
      }
    } catch (QError& e) {
      std::lock_guard<std::mutex> lk(err_mu);
      if (!err_code) err_code = e.code, err = e.what();
    } catch (std::exception& e) {
      std::lock_guard<std::mutex> lk(err_mu);
      if (!err_code) err_code = QGPU_ERR_INTERNAL, err = e.what();
    }
  };
  if (threads == 1) {
    worker(0);
  } else {
    std::vector<std::thread> pool;
    for (int w = 1; w < threads; ++w) pool.emplace_back(worker, w);
    worker(0);
    for (auto& th : pool) th.join();
  }
  if (err_code) {
    cudaStreamSynchronize(ctx->copy_stream);
    throw QError(err_code, err);
  }
}

static double now_ms() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

TableChunk import_host_batches(Ctx* ctx, const Schema& schema, const std::vector<ArrowArray>& batches, const std::vector<char>& want,
                               int host_narrow /* -1: context default */) {
  static const bool trace = env_int("QGPU_INGEST_TRACE", 0) != 0;
  const double t_start = now_ms();
  TableChunk ch;
  ch.cols.resize(schema.fields.size());
  int64_t rows = 0;
  for (auto& b : batches) rows += b.length;
  ch.rows = rows;
  if (host_narrow < 0) host_narrow = ctx->ingest_host_narrow;
  if (host_narrow < 0) host_narrow = env_int("QGPU_INGEST_HOST_NARROW", -1);
  // -1 = automatic, decided per column below: a source the workers must touch anyway (pageable, or many small batches) is
  // narrowed on the way; a large page-locked source is narrowed by the host only when this rank has enough worker threads
  // to keep PCIe busy (>= 8) -- otherwise it is DMA'd as 16-byte values and narrowed by one kernel.  Measured, Q1 SF10
  // lineitem per GPU: 16 workers 62 ms vs 89 ms direct; 4 workers per rank (8 ranks on a 32-thread host) 357 ms vs 218 ms.
  const bool many_workers = ingest_worker_threads(ctx) >= 8;
  const size_t slot_bytes = ctx->ingest_slot_bytes;
  const int64_t n_words = (rows + 31) >> 5;
  const size_t kDirectMin = (size_t)256 << 10;  // page-locked source buffers at least this large are DMA'd directly

  std::vector<std::unique_ptr<ColPlan>> plans;
  for (size_t f = 0; f < schema.fields.size(); ++f) {
    if (!want[f]) continue;
    const DType& ty = schema.fields[f].type;
    auto col = std::make_shared<DCol>();
    col->type = ty;
    col->length = rows;
    col->phys = phys_of(ty);
    ch.cols[f] = col;
    if (ty.id == QGPU_T_NULL) {
      col->null_count = rows;
      continue;
    }
    auto cp = std::make_unique<ColPlan>();
    cp->field = (int)f;
    cp->col = col;
    cp->rows = rows;
    cp->w = arrow_width(ty);
    cp->narrow = host_narrow != 0 && col->phys == PH_I128 && ty.precision <= 18 && rows > 0;
    plans.push_back(std::move(cp));
  }
  // one Part per (column, non-empty batch): with tens of thousands of batches this walk over separately allocated
  // ArrowArray structs is pointer chasing, so the columns are walked concurrently
  auto plan_column = [&](ColPlan& cp) {
    const size_t f = (size_t)cp.field;
    const bool is_str = cp.col->phys == PH_STR;
    int64_t row = 0, byte = 0;
    cp.parts.reserve(batches.size());
    for (auto& b : batches) {
      const ArrowArray* a = b.children[f];
      Part p;
      p.n = b.length;
      p.off = a->offset + b.offset;
      p.row0 = row;
      row += p.n;
      if (p.n == 0) continue;
      p.validity = (const uint8_t*)a->buffers[0];
      p.nulls = p.validity ? a->null_count : 0;
      if (p.nulls < 0) p.nulls = count_zero_bits_host(p.validity, p.off, p.n);
      if (is_str) {
        p.offsets = (const int32_t*)a->buffers[1];
        p.data = (const char*)a->buffers[2];
        p.o0 = p.offsets[p.off];
        p.sbytes = (int64_t)p.offsets[p.off + p.n] - p.o0;
        if (p.sbytes < 0) throw_arrow("Utf8 offsets are not ascending");
        if (p.sbytes > 0 && !p.data) throw_arrow("Utf8 column has no data buffer");
        p.byte0 = byte;
        byte += p.sbytes;
      } else {
        p.values = (const char*)a->buffers[1];
      }
      cp.nulls += p.nulls;
      cp.parts.push_back(p);
    }
    cp.str_bytes = byte;
    if (byte > 2147483647LL) throw_arrow("Utf8 column exceeds 2 GiB of string data; LargeUtf8 is not supported");
    cp.col->null_count = cp.nulls;
  };
  if (batches.size() >= 2048 && plans.size() > 1) {
    std::mutex err_mu;
    std::string err;
    int err_code = 0;
    std::atomic<size_t> next(0);
    auto w = [&] {
      for (;;) {
        const size_t i = next.fetch_add(1);
        if (i >= plans.size()) break;
        try {
          plan_column(*plans[i]);
        } catch (QError& e) {
          std::lock_guard<std::mutex> lk(err_mu);
          if (!err_code) err_code = e.code, err = e.what();
        }
      }
    };
    std::vector<std::thread> pool;
    const size_t nt = std::min<size_t>(plans.size(), 8);
    for (size_t i = 1; i < nt; ++i) pool.emplace_back(w);
    w();
    for (auto& th : pool) th.join();
    if (err_code) throw QError(err_code, err);
  } else {
    for (auto& cpp : plans) plan_column(*cpp);
  }

  if (host_narrow < 0 && !many_workers)
    for (auto& cpp : plans) {
      ColPlan& cp = *cpp;
      if (!cp.narrow || cp.parts.empty()) continue;
      bool all_big = true;
      for (auto& p : cp.parts) all_big = all_big && (size_t)p.n * cp.w >= kDirectMin;
      if (all_big && cp.parts[0].values && is_pinned(cp.parts[0].values)) cp.narrow = false;  // direct DMA + k_narrow
    }
  // destination buffers (stream-ordered allocations on the compute stream; the copy stream waits for them below)
  for (auto& cpp : plans) {
    ColPlan& cp = *cpp;
    DCol& col = *cp.col;
    if (col.phys == PH_BIT) col.data = ctx->alloc(std::max<size_t>((size_t)n_words * 4, 4));
    else if (col.phys == PH_STR) {
      col.offsets = rows > 0 ? ctx->alloc((size_t)(rows + 1) * 4) : ctx->alloc_zero(4);
      col.data = ctx->alloc(std::max<size_t>((size_t)cp.str_bytes, 4));
      col.str_bytes = cp.str_bytes;
    } else {
      col.data = ctx->alloc(std::max<size_t>((size_t)rows * (cp.narrow ? 8 : cp.w), 16));
    }
    if (cp.nulls > 0) col.validity = ctx->alloc(std::max<size_t>((size_t)n_words * 4, 4));
  }
  if (rows == 0) return ch;
  {
    cudaEvent_t ready;
    CUDA_CHECK(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventRecord(ready, ctx->stream));
    CUDA_CHECK(cudaStreamWaitEvent(ctx->copy_stream, ready, 0));
    CUDA_CHECK(cudaEventDestroy(ready));
  }

  // direct DMAs first (they keep PCIe busy while the workers fill their first slots), then the staged tasks
  std::vector<Task> tasks;
  size_t staged_bytes = 0;
  auto add_tasks = [&](ColPlan* cp, TaskKind kind, int64_t units, int64_t unit_bytes) {
    const int64_t per = std::max<int64_t>(1, (int64_t)slot_bytes / unit_bytes);
    for (int64_t u = 0; u < units; u += per) tasks.push_back({cp, kind, u, std::min(units, u + per)});
    staged_bytes += (size_t)(units * unit_bytes);
  };
  for (auto& cpp : plans) {
    ColPlan& cp = *cpp;
    DCol& col = *cp.col;
    if (cp.nulls > 0) add_tasks(&cp, TK_VALIDITY, n_words, 4);
    if (cp.parts.empty()) continue;
    const bool all_big = [&] {
      for (auto& p : cp.parts) {
        const size_t b = col.phys == PH_STR ? (size_t)p.sbytes : (size_t)p.n * cp.w;
        if (b < kDirectMin) return false;
      }
      return true;
    }();
    if (col.phys == PH_BIT) {
      add_tasks(&cp, TK_BOOLBITS, n_words, 4);
    } else if (col.phys == PH_STR) {
      const Part& p0 = cp.parts[0];
      if (cp.parts.size() == 1 && p0.o0 == 0 && (size_t)(rows + 1) * 4 >= kDirectMin && is_pinned(p0.offsets))
        CUDA_CHECK(cudaMemcpyAsync(col.offsets->ptr, p0.offsets + p0.off, (size_t)(rows + 1) * 4, cudaMemcpyHostToDevice, ctx->copy_stream));
      else
        add_tasks(&cp, TK_OFFSETS, rows + 1, 4);
      if (cp.str_bytes > 0) {
        if (all_big && is_pinned(p0.data)) {
          for (auto& p : cp.parts)
            CUDA_CHECK(cudaMemcpyAsync((char*)col.data->ptr + p.byte0, p.data + p.o0, (size_t)p.sbytes, cudaMemcpyHostToDevice, ctx->copy_stream));
        } else {
          add_tasks(&cp, TK_STRDATA, cp.str_bytes, 1);
        }
      }
    } else if (cp.narrow) {
      add_tasks(&cp, TK_NARROW, rows, 8);
    } else if (all_big && cp.parts[0].values && is_pinned(cp.parts[0].values)) {
      for (auto& p : cp.parts) {
        if (p.values)
          CUDA_CHECK(cudaMemcpyAsync((char*)col.data->ptr + p.row0 * cp.w, p.values + p.off * cp.w, (size_t)p.n * cp.w, cudaMemcpyHostToDevice, ctx->copy_stream));
        else
          CUDA_CHECK(cudaMemsetAsync((char*)col.data->ptr + p.row0 * cp.w, 0, (size_t)p.n * cp.w, ctx->copy_stream));
      }
    } else {
      add_tasks(&cp, TK_FIXED, rows, cp.w);
    }
  }
  const double t_plan = now_ms();
  run_tasks(ctx, tasks, staged_bytes);
  const double t_run = now_ms();
  CUDA_CHECK(cudaStreamSynchronize(ctx->copy_stream));
  if (trace)
    fprintf(stderr, "[qgpu ingest] %zu batches, %lld rows, %zu tasks, %.1f MB staged: plan %.2f ms, fill+queue %.2f ms, drain %.2f ms\n",
            batches.size(), (long long)rows, tasks.size(), staged_bytes / 1e6, t_plan - t_start, t_run - t_plan, now_ms() - t_run);  // the host batches may be released after this; data visible to every stream

  // results of the host narrowing; a value outside int64 (malformed for p <= 18) sends that column back as 16 B values
  for (auto& cpp : plans) {
    ColPlan& cp = *cpp;
    DCol& col = *cp.col;
    if (cp.narrow) {
      if (cp.fits) {
        col.phys = PH_D64;
        if (cp.mn <= cp.mx) {
          col.has_stats = true;
          col.vmin = cp.mn;
          col.vmax = cp.mx;
        }
      } else {
        std::vector<char> only(schema.fields.size(), 0);
        only[cp.field] = 1;
        TableChunk wide = import_host_batches(ctx, schema, batches, only, 0);
        ch.cols[cp.field] = wide.cols[cp.field];
      }
    } else if (col.phys == PH_I128 && col.type.precision <= 18) {
      DColP nar = try_narrow_decimal(ctx, col);  // host narrowing off: one pass on the device over the whole column
      if (nar) ch.cols[cp.field] = nar;
    }
  }
  return ch;
}

}  // namespace qgpu
