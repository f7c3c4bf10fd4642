// Plan execution + table consolidation + the extern "C" surface declared in include/qgpu.h.
#include <cstring>
#include <thread>

#include "comm.h"
#include "fused_jit.h"
#include "launch.h"
#include "plan.h"

using namespace qgpu;

namespace qgpu {

// ------------------------------------------------------------------------------------------------
// table consolidation: appended batches -> one contiguous column each
// ------------------------------------------------------------------------------------------------
__global__ void k_widen_chunk(const int64_t* __restrict__ src, ulonglong2* __restrict__ dst, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    int64_t x = src[i];
    dst[i] = make_ulonglong2((uint64_t)x, x < 0 ? ~0ull : 0ull);
  }
}

// The producer's release callbacks, on the calling thread (~1 us per 1024-row batch; pyarrow's callbacks contend for
// its allocator / the GIL when called from several threads: 8 threads took 115 ms instead of 55 ms for SF10's lineitem)
static void release_host_batches(std::vector<ArrowArray>& v) {
  for (auto& b : v)
    if (b.release) b.release(&b);
}

TableImpl::~TableImpl() { release_host_batches(pending_host); }

// uploads every retained host batch as ONE chunk (ingest.cu)
void TableImpl::flush_pending() {
  if (pending_host.empty()) return;
  TableChunk ch = import_host_batches(ctx, schema, pending_host, pending_want);
  if (consolidated && num_batches > (int64_t)pending_host.size()) {
    // re-open: keep the consolidated columns as the first chunk
    TableChunk first;
    first.cols = cols;
    first.rows = num_rows - ch.rows;
    chunks.push_back(first);
  }
  chunks.push_back(ch);
  consolidated = false;
  release_host_batches(pending_host);
  pending_host.clear();
}

void TableImpl::consolidate() {
  resolve();  // result table of an asynchronous execute: its row count is needed from here on
  flush_pending();
  if (consolidated) return;
  const size_t nf = schema.fields.size();
  cols.assign(nf, nullptr);
  for (size_t f = 0; f < nf; ++f) {
    std::vector<DColP> parts;
    bool missing = false;
    for (auto& ch : chunks) {
      if (!ch.cols[f]) missing = true;
      else parts.push_back(ch.cols[f]);
    }
    if (parts.empty()) continue;
    if (missing) throw_internal("column '" + schema.fields[f].name + "' was uploaded for some batches only");
    if (parts.size() == 1) {
      cols[f] = parts[0];
      continue;
    }
    auto out = std::make_shared<DCol>();
    out->type = schema.fields[f].type;
    out->length = num_rows;
    Phys phys = parts[0]->phys;
    bool any_valid_buf = false, all_null_phys = true;
    for (auto& p : parts) {
      if (p->phys == PH_I128) phys = PH_I128;  // one wide chunk forces the wide layout
      if (p->validity) any_valid_buf = true;
      if (p->phys != PH_NULL) all_null_phys = false;
      out->null_count += p->null_count;
      if (p->str_bytes_is_bound && p->phys == PH_STR && p->offsets) {  // the concatenation below needs exact sizes
        p->str_bytes = (int64_t)ctx->read_scalar((const int32_t*)p->offsets->ptr + p->length);
        p->str_bytes_is_bound = false;
      }
      out->str_bytes += p->str_bytes;
    }
    if (all_null_phys) {
      out->phys = PH_NULL;
      out->null_count = num_rows;
      cols[f] = out;
      continue;
    }
    if (phys == PH_NULL) phys = parts.back()->phys;
    for (auto& p : parts)
      if (p->phys != PH_NULL && p->phys != phys && !(p->phys == PH_D64 && phys == PH_I128) && !(p->phys == PH_I128 && phys == PH_D64))
        throw_internal("inconsistent chunk layouts");
    for (auto& p : parts)
      if (p->phys == PH_I128) phys = PH_I128;
    out->phys = phys;
    const int w = phys_width(phys);
    const int64_t n_words = (num_rows + 31) >> 5;
    if (phys == PH_BIT) out->data = ctx->alloc_zero((size_t)n_words * 4 + 4);
    else if (phys == PH_STR) {
      if (out->str_bytes > 2147483647LL) throw_arrow("Utf8 column exceeds 2 GiB of string data; LargeUtf8 is not supported");
      out->data = ctx->alloc(std::max<size_t>((size_t)out->str_bytes, 4));
      out->offsets = ctx->alloc_zero((size_t)(num_rows + 1) * 4);
    } else {
      out->data = ctx->alloc(std::max<size_t>((size_t)num_rows * w, 16));
    }
    if (any_valid_buf || out->null_count > 0) out->validity = ctx->alloc_zero((size_t)n_words * 4 + 4);
    int64_t row = 0, byte = 0;
    bool stats_ok = true;
    i128 mn = 0, mx = 0;
    bool first_stats = true;
    for (auto& p : parts) {
      const int64_t n = p->length;
      if (n == 0) continue;
      if (out->validity) {
        if (p->phys == PH_NULL) fill_bits(ctx, (uint32_t*)out->validity->ptr, row, n, false);
        else if (p->validity) copy_bits(ctx, (uint32_t*)out->validity->ptr, row, (const uint32_t*)p->validity->ptr, 0, n);
        else fill_bits(ctx, (uint32_t*)out->validity->ptr, row, n, true);
      }
      if (p->phys == PH_NULL) {
        if (phys == PH_STR) {
          // offsets of NULL rows all equal the running byte offset
          std::vector<int32_t> o((size_t)n + 1, (int32_t)byte);
          ctx->h2d((int32_t*)out->offsets->ptr + row, o.data(), o.size() * 4);
          ctx->sync();
        } else if (phys != PH_BIT) {
          CUDA_CHECK(cudaMemsetAsync((char*)out->data->ptr + row * w, 0, (size_t)n * w, ctx->stream));
        }
      } else if (phys == PH_BIT) {
        copy_bits(ctx, (uint32_t*)out->data->ptr, row, (const uint32_t*)p->data->ptr, 0, n);
      } else if (phys == PH_STR) {
        rebase_offsets(ctx, (int32_t*)out->offsets->ptr + row, (const int32_t*)p->offsets->ptr, n + 1, byte);
        if (p->str_bytes > 0)
          CUDA_CHECK(cudaMemcpyAsync((char*)out->data->ptr + byte, p->data->ptr, (size_t)p->str_bytes, cudaMemcpyDeviceToDevice,
                                     ctx->stream));
        byte += p->str_bytes;
      } else if (p->phys == PH_D64 && phys == PH_I128) {
        LAUNCH(ctx, k_widen_chunk, grid_for(ctx, n, 256), 256, 0, (const int64_t*)p->data->ptr,
               (ulonglong2*)out->data->ptr + row, n);
      } else {
        CUDA_CHECK(cudaMemcpyAsync((char*)out->data->ptr + row * w, p->data->ptr, (size_t)n * w, cudaMemcpyDeviceToDevice,
                                   ctx->stream));
      }
      if (p->has_stats) {
        if (first_stats) {
          mn = p->vmin;
          mx = p->vmax;
          first_stats = false;
        } else {
          mn = std::min(mn, p->vmin);
          mx = std::max(mx, p->vmax);
        }
      } else if (p->null_count != p->length) {
        stats_ok = false;
      }
      row += n;
    }
    if (stats_ok && !first_stats) {
      out->has_stats = true;
      out->vmin = mn;
      out->vmax = mx;
    }
    if (out->null_count == 0) out->validity.reset();
    cols[f] = out;
  }
  ctx->sync();
  chunks.clear();
  consolidated = true;
}

// ------------------------------------------------------------------------------------------------
// execution
// ------------------------------------------------------------------------------------------------
View scan_view(PlanNode& n) {
  TableImpl& t = *n.table;
  t.consolidate();
  View v;
  v.num_rows = t.num_rows;
  v.num_batches = t.num_batches;
  if (t.num_batches == 0) {
    // CREATE TABLE without INSERT: no buffers exist; give every column an empty all-NULL placeholder so
    // that outer joins can still emit NULL-padded rows for this side
    for (size_t i = 0; i < t.cols.size(); ++i)
      if (!t.cols[i]) {
        auto c = std::make_shared<DCol>();
        c->type = t.schema.fields[i].type;
        c->phys = PH_NULL;
        t.cols[i] = c;
      }
  }
  if (n.has_projection) {
    for (int ci : n.projection) {
      if (ci < 0 || ci >= (int)t.schema.fields.size()) throw_arrow("Schema error: projection index out of bounds");
      v.schema.fields.push_back(t.schema.fields[ci]);
      v.cols.push_back({t.cols[ci], nullptr});
    }
    v.schema.metadata = t.schema.metadata;
  } else {
    v.schema = t.schema;
    for (auto& c : t.cols) v.cols.push_back({c, nullptr});
  }
  return v;
}

// a child's result with its metadata resolved (row count / NULL counts known on the host): what every operator but a
// pure column Projection needs
View PlanNode::child_view(int i) {
  View v = children[i]->execute();
  v.resolve();
  return v;
}

View PlanNode::execute() {
  switch (kind) {
    case PK_SCAN: {
      View v = scan_view(*this);
      strategy = "scan";
      if (predicate) {
        auto c = compile_expr(*predicate, v.schema);
        IdxP sel = eval_filter(ctx, *c, v);
        v = apply_selection_view(ctx, v, sel);
        strategy = "scan+filter(selection-vector)";
      }
      return v;
    }
    case PK_FILTER: {
      View in = child_view(0);
      auto c = compile_expr(*predicate, in.schema);
      IdxP sel = eval_filter(ctx, *c, in);
      strategy = "filter(selection-vector)";
      return apply_selection_view(ctx, in, sel);
    }
    case PK_PROJECTION: {
      View in = children[0]->execute();
      if (in.pending) {
        // result metadata still in flight (dense fused aggregate): a Projection of plain columns only re-orders column
        // handles and passes the pending metadata on; anything computed needs the row count now
        bool pure = exprs.size() == schema.fields.size();
        for (size_t i = 0; pure && i < exprs.size(); ++i) {
          auto c = compile_expr(*exprs[i], in.schema);
          pure = c->is_column_ref && c->result_type == schema.fields[i].type;
        }
        if (!pure) in.resolve();
      }
      View out;
      out.pending = in.pending;
      out.schema = schema;
      out.num_rows = in.num_rows;
      out.num_batches = in.num_batches;
      if (exprs.size() != schema.fields.size()) throw_arrow("number of columns must match number of fields in schema");
      for (size_t i = 0; i < exprs.size(); ++i) {
        auto c = compile_expr(*exprs[i], in.schema);
        if (in.num_batches > 0 && c->result_type != schema.fields[i].type)
          throw_arrow("column types must match schema types, expected " + schema.fields[i].type.str() + " but found " +
                      c->result_type.str() + " at column index " + std::to_string(i));
        if (c->is_column_ref) out.cols.push_back(in.cols[c->column_ref]);
        else out.cols.push_back({eval_to_column(ctx, *c, in), nullptr});
      }
      strategy = "projection";
      return out;
    }
    case PK_AGGREGATE: {
      if (merged_override) {  // sharded execution: the merged states were finalised by shard_execute_merged
        View v = *merged_override;
        merged_override.reset();
        return v;
      }
      View fused;
      if (try_fused_scan_aggregate(*this, &fused)) return fused;
      if (try_fused_join_aggregate(*this, &fused)) return fused;
      View in = child_view(0);
      std::vector<std::shared_ptr<Compiled>> keys;
      for (auto& e : group_exprs) keys.push_back(compile_expr(*e, in.schema));
      std::vector<AggSpec> specs;
      for (auto& a : aggs) {
        AggSpec s;
        s.op = a.op;
        s.arg = compile_expr(*a.expr, in.schema);
        s.return_type = a.return_type;
        s.expr_type = a.expr_type;
        specs.push_back(s);
      }
      strategy = group_exprs.empty() ? "generic-no-grouping-aggregate" : "generic-hash-aggregate";
      return run_aggregate(ctx, in, keys, specs, schema, defer);
    }
    case PK_SORT: {
      View in = child_view(0);
      return run_sort(*this, in);
    }
    case PK_LIMIT: {
      View in = child_view(0);
      return run_limit(*this, in);
    }
    case PK_HASH_JOIN: {
      if (order_free) {
        View uj;
        if (fused_unordered_join(*this, &uj)) return uj;
      }
      View l = child_view(0);
      View r = child_view(1);
      std::vector<std::shared_ptr<Compiled>> lo, ro;
      for (auto& e : left_on) lo.push_back(compile_expr(*e, l.schema));
      for (auto& e : right_on) ro.push_back(compile_expr(*e, r.schema));
      JoinFilterSpec fs;
      if (has_join_filter) {
        fs.schema = join_filter_schema;
        fs.column_index = join_filter_index;
        fs.column_side = join_filter_side;
        fs.expr = compile_expr(*join_filter_expr, join_filter_schema);
      }
      strategy = "generic-hash-join(late-materialisation)";
      return run_hash_join(ctx, l, r, join_type, lo, ro, has_join_filter ? &fs : nullptr, schema);
    }
    case PK_NL_JOIN: {
      View l = child_view(0);
      View r = child_view(1);
      JoinFilterSpec fs;
      if (has_join_filter) {
        fs.schema = join_filter_schema;
        fs.column_index = join_filter_index;
        fs.column_side = join_filter_side;
        fs.expr = compile_expr(*join_filter_expr, join_filter_schema);
      }
      strategy = "nested-loop-join(cross pairs + filter, late-materialisation)";
      return run_nested_loop_join(ctx, l, r, join_type, has_join_filter ? &fs : nullptr, schema);
    }
    case PK_BROADCAST: return run_broadcast(*this);
    case PK_FINAL_AGG: return run_final_aggregate(*this);
    case PK_CROSS_JOIN: {
      View l = child_view(0);
      View r = child_view(1);
      strategy = "cross-join(index vectors, late-materialisation)";
      return run_cross_join(ctx, l, r, schema);
    }
  }
  throw_internal("unknown plan node");
}

static std::vector<DColP> materialize_view(Ctx* ctx, const View& v) {
  std::vector<DColP> cols;
  for (size_t i = 0; i < v.cols.size(); ++i) {
    if (!v.cols[i].base)
      throw_internal("column '" + v.schema.fields[i].name + "' must be returned but was not uploaded to the GPU table");
    cols.push_back(materialize_arrow(ctx, v.cols[i], v.num_rows));
  }
  return cols;
}

}  // namespace qgpu

// ================================================================================================
// extern "C"
// ================================================================================================
static thread_local std::string g_last_error = "";

template <typename F>
static int guard(Ctx* ctx, F&& f) {
  try {
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) throw QError(QGPU_ERR_CUDA, std::string("CUDA error: ") + cudaGetErrorString(e));
    ++ctx->exec_epoch;
    f();
    return QGPU_OK;
  } catch (QError& e) {
    ctx->last_error = e.what();
    cudaGetLastError();
    return e.code;
  } catch (std::bad_alloc&) {
    ctx->last_error = "out of host memory";
    return QGPU_ERR_OOM;
  } catch (std::exception& e) {
    ctx->last_error = std::string("InternalError: ") + e.what();
    return QGPU_ERR_INTERNAL;
  }
}

extern "C" {

int qgpu_init(const int* devices, int n, qgpu_ctx** out) {
  if (!out) return QGPU_ERR_INTERNAL;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    g_last_error = std::string("CUDA error: no CUDA device available (") + cudaGetErrorString(e) +
                   "); libqgpu has no CPU fallback";
    cudaGetLastError();
    return QGPU_ERR_CUDA;
  }
  int dev = (devices && n > 0) ? devices[0] : 0;
  if (n > 1) {
    g_last_error = "InternalError: one qgpu_ctx drives one GPU (one process per GPU); create one context per device";
    return QGPU_ERR_INTERNAL;
  }
  if (dev < 0 || dev >= count) {
    g_last_error = "InternalError: invalid CUDA device ordinal " + std::to_string(dev);
    return QGPU_ERR_INTERNAL;
  }
  qgpu_ctx* h = new qgpu_ctx();
  Ctx* c = &h->c;
  c->device = dev;
  try {
    CUDA_CHECK(cudaSetDevice(dev));
    CUDA_CHECK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CUDA_CHECK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CUDA_CHECK(cudaStreamCreateWithFlags(&c->epi_stream, cudaStreamNonBlocking));
    CUDA_CHECK(cudaEventCreateWithFlags(&c->epi_ready, cudaEventDisableTiming));
    // One stream-ordered pool PER CONTEXT: contexts of one process (ranks emulated on one GPU, several qurious sessions)
    // must not share a pool -- the driver may make an allocation on one context's stream wait for a free that is still
    // queued on another's, and a peer-exchange kernel waiting for that other context would never see it arrive.
    cudaMemPoolProps props;
    memset(&props, 0, sizeof(props));
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    CUDA_CHECK(cudaMemPoolCreate(&c->pool, &props));
    uint64_t thr = UINT64_MAX;
    CUDA_CHECK(cudaMemPoolSetAttribute(c->pool, cudaMemPoolAttrReleaseThreshold, &thr));
    // pinned slots for asynchronous result metadata: allocated up front (a page-locked allocation while a peer-exchange
    // kernel of another context spins may synchronise with it)
    for (int i = 0; i < 8; ++i) {
      MetaSlot* m = new MetaSlot();
      CUDA_CHECK(cudaHostAlloc((void**)&m->host, META_WORDS * 8, cudaHostAllocDefault));
      CUDA_CHECK(cudaEventCreateWithFlags(&m->ev, cudaEventDisableTiming));
      c->meta_slots.push_back(m);
    }
    cudaDeviceProp prop;
    CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
    c->sm_count = prop.multiProcessorCount;
    for (int i = 0; i < Ctx::kStageSlots; ++i) {
      CUDA_CHECK(cudaHostAlloc(&c->stage[i], c->stage_bytes, cudaHostAllocDefault));
      CUDA_CHECK(cudaEventCreateWithFlags(&c->stage_ev[i], cudaEventDisableTiming));
    }
    CUDA_CHECK(cudaHostAlloc(&c->pinned_scratch, c->pinned_scratch_bytes, cudaHostAllocDefault));
    if (const char* e = getenv("QGPU_POOL_RESERVE_MB")) c->pool_floor_bytes = (size_t)std::max(0, atoi(e)) << 20;
    c->reserve_pool();
  } catch (QError& err) {
    g_last_error = err.what();
    delete h;
    return err.code;
  }
  *out = h;
  return QGPU_OK;
}

void qgpu_shutdown(qgpu_ctx* ctx) {
  if (!ctx) return;
  Ctx* c = &ctx->c;
  cudaSetDevice(c->device);
  c->release_big_blocks();
  cudaStreamSynchronize(c->stream);
  cudaStreamSynchronize(c->copy_stream);
  if (c->epi_stream) cudaStreamSynchronize(c->epi_stream);
  for (auto& e : c->prof_events) {
    cudaEventDestroy(e.first);
    cudaEventDestroy(e.second);
  }
  for (int i = 0; i < Ctx::kStageSlots; ++i) {
    if (c->stage[i]) cudaFreeHost(c->stage[i]);
    if (c->stage_ev[i]) cudaEventDestroy(c->stage_ev[i]);
  }
  for (cudaEvent_t e : c->ingest_ev) cudaEventDestroy(e);
  if (c->pinned_scratch) cudaFreeHost(c->pinned_scratch);
  for (MetaSlot* m : c->meta_slots) {
    cudaFreeHost(m->host);
    cudaEventDestroy(m->ev);
    delete m;
  }
  c->comm.reset();
  cudaStreamDestroy(c->stream);
  cudaStreamDestroy(c->copy_stream);
  if (c->epi_stream) cudaStreamDestroy(c->epi_stream);
  if (c->epi_ready) cudaEventDestroy(c->epi_ready);
  if (c->pool) cudaMemPoolDestroy(c->pool);
  delete ctx;
}

const char* qgpu_last_error(const qgpu_ctx* ctx) { return ctx ? ctx->c.last_error.c_str() : g_last_error.c_str(); }

int qgpu_set_compat(qgpu_ctx* ctx, const char* name, int value) {
  if (!ctx || !name) return QGPU_ERR_INTERNAL;
  return guard(&ctx->c, [&] {
    std::string s(name);
    if (s == "avg_precision") ctx->c.compat_avg_precision = value != 0;
    else if (s == "empty_decimal_sum") ctx->c.compat_empty_decimal_sum = value != 0;
    else throw_internal("unknown compat switch '" + s + "'");
  });
}

int64_t qgpu_jit_compile(const uint64_t* signature4, uint32_t pack, char* log_buf, int64_t cap) {
  if (!signature4) return -1;
  try {
    std::string log;
    const std::string cubin = jit_compile_cubin(signature4, pack, &log);
    if (log_buf && cap > 0) {
      const size_t n = std::min<size_t>((size_t)cap - 1, log.size());
      memcpy(log_buf, log.data(), n);
      log_buf[n] = 0;
    }
    return (int64_t)cubin.size();
  } catch (std::exception& e) {
    g_last_error = e.what();
    return -1;
  }
}

int64_t qgpu_counter(const qgpu_ctx* ctx, const char* name) {
  if (!ctx || !name) return -1;
  const std::string s(name);
  if (s == "alloc_bytes") return ctx->c.alloc_bytes_total;
  if (s == "gather_bytes") return ctx->c.gather_bytes_total;
  if (s == "kernel_launches") return ctx->c.launches;
  return -1;
}

int64_t qgpu_kernel_launches(const qgpu_ctx* ctx) { return ctx ? ctx->c.launches : 0; }

int qgpu_release_cached_memory(qgpu_ctx* ctx) {
  if (!ctx) return QGPU_ERR_INTERNAL;
  return guard(&ctx->c, [&] {
    ctx->c.release_big_blocks();
    ctx->c.sync();
    // back to the driver (other allocators of the process can use it) -- down to the floor the context keeps reserved
    CUDA_CHECK(cudaMemPoolTrimTo(ctx->c.pool, ctx->c.pool_floor_bytes));
  });
}

void* qgpu_ctx_stream(const qgpu_ctx* ctx) { return ctx ? (void*)ctx->c.stream : nullptr; }

int qgpu_profile_enable(qgpu_ctx* ctx, int on) {
  if (!ctx) return QGPU_ERR_INTERNAL;
  return guard(&ctx->c, [&] {
    ctx->c.profiling = on != 0;
    ctx->c.prof_min_blocks = on >= 2 ? on : 0;
    if (!on) ctx->c.prof_report();
  });
}

int64_t qgpu_profile_report(qgpu_ctx* ctx, char* buf, int64_t cap) {
  if (!ctx) return -1;
  int64_t need = -1;
  guard(&ctx->c, [&] {
    std::string r = ctx->c.prof_report();
    need = (int64_t)r.size() + 1;
    if (buf && cap > 0) {
      size_t n = std::min<size_t>(r.size(), (size_t)cap - 1);
      memcpy(buf, r.data(), n);
      buf[n] = 0;
    }
  });
  return need;
}

// ---- tables ------------------------------------------------------------------------------------
int qgpu_table_create(qgpu_ctx* ctx, const struct ArrowSchema* schema, qgpu_table** out) {
  if (!ctx || !out) return QGPU_ERR_INTERNAL;
  return guard(&ctx->c, [&] {
    auto t = std::make_shared<TableImpl>();
    t->ctx = &ctx->c;
    t->schema = import_schema(schema);
    t->cols.assign(t->schema.fields.size(), nullptr);
    *out = new qgpu_table{t};
  });
}

static std::vector<char> want_columns(const Schema& schema, const int32_t* cols, int32_t n) {
  std::vector<char> want(schema.fields.size(), cols ? 0 : 1);
  if (cols)
    for (int32_t i = 0; i < n; ++i) {
      if (cols[i] < 0 || cols[i] >= (int32_t)schema.fields.size()) throw_internal("upload column index out of range");
      want[cols[i]] = 1;
    }
  return want;
}

// A HOST batch is validated and retained; the upload of all retained batches happens at once, when the table is first
// used (ingest.cu).  Takes ownership of *batch (moved into the table; released after the upload).
static void retain_host_batch(TableImpl& ti, struct ArrowArray* batch, const std::vector<char>& want) {
  validate_host_batch(ti.schema, batch, want);
  if (!ti.pending_host.empty() && ti.pending_want != want) ti.flush_pending();
  ti.pending_want = want;
  ti.pending_host.push_back(*batch);
  batch->release = nullptr;  // moved
  ti.num_rows += ti.pending_host.back().length;
  ti.num_batches += 1;
}

static int table_append(qgpu_table* t, struct ArrowArray* batch, const int32_t* cols, int32_t n, bool dev) {
  if (!t || !batch) return QGPU_ERR_INTERNAL;
  TableImpl& ti = *t->t;
  int rc = guard(ti.ctx, [&] {
    ti.resolve();
    if (!dev) {
      retain_host_batch(ti, batch, want_columns(ti.schema, cols, n));
      return;
    }
    ti.flush_pending();  // row order: earlier host batches first
    // import first: a failing batch (column count, short child, CUDA error) must leave the table untouched
    TableChunk ch = import_batch(ti.ctx, ti.schema, batch, cols, n, dev);
    if (ti.consolidated && ti.num_batches > 0) {
      // re-open: keep the consolidated columns as the first chunk
      TableChunk first;
      first.cols = ti.cols;
      first.rows = ti.num_rows;
      ti.chunks.push_back(first);
    }
    ti.chunks.push_back(ch);
    ti.num_rows += ch.rows;
    ti.num_batches += 1;
    ti.consolidated = false;
  });
  if (batch->release) batch->release(batch);  // takes ownership, also on failure
  return rc;
}

int qgpu_table_append(qgpu_table* t, struct ArrowArray* batch, const int32_t* upload_columns, int32_t n) {
  return table_append(t, batch, upload_columns, n, false);
}
int qgpu_table_append_device(qgpu_table* t, struct ArrowArray* batch) { return table_append(t, batch, nullptr, 0, true); }

int qgpu_table_append_stream(qgpu_table* t, struct ArrowArrayStream* stream, const int32_t* upload_columns, int32_t n, int64_t* out_batches) {
  if (!t || !stream || !stream->get_next) return QGPU_ERR_INTERNAL;
  TableImpl& ti = *t->t;
  int64_t got = 0;
  int rc = guard(ti.ctx, [&] {
    ti.resolve();
    const std::vector<char> want = want_columns(ti.schema, upload_columns, n);
    for (;;) {
      ArrowArray b;
      memset(&b, 0, sizeof(b));
      const int e = stream->get_next(stream, &b);
      if (e != 0) {
        const char* m = stream->get_last_error ? stream->get_last_error(stream) : nullptr;
        throw_arrow(std::string("ArrowArrayStream::get_next failed: ") + (m ? m : "(no message)"));
      }
      if (!b.release) break;  // end of stream
      try {
        retain_host_batch(ti, &b, want);
      } catch (...) {
        if (b.release) b.release(&b);
        throw;
      }
      ++got;
    }
  });
  if (out_batches) *out_batches = got;
  ti.ctx->trace("append_stream: pulled");
  if (stream->release) stream->release(stream);  // takes ownership, also on failure
  return rc;
}

static int table_append_csv(qgpu_table* t, const void* host_text, int64_t len, const char* path, const qgpu_csv_options* options,
                            const int32_t* cols, int32_t n, int64_t* out_rows) {
  if (!t || (!host_text && !path) || len < 0) return QGPU_ERR_INTERNAL;
  TableImpl& ti = *t->t;
  return guard(ti.ctx, [&] {
    ti.resolve();
    qgpu_csv_options opt;
    memset(&opt, 0, sizeof(opt));
    if (options) opt = *options;
    const std::vector<char> want = want_columns(ti.schema, cols, n);
    ti.flush_pending();  // row order: earlier host batches first
    DBufP text;
    int64_t bytes = len;
    if (path) {
      text = read_file_to_device(ti.ctx, path, &bytes);
    } else {
      text = ti.ctx->alloc(std::max<size_t>((size_t)len, 16));
      if (len > 0) ti.ctx->h2d(text->ptr, host_text, (size_t)len);
    }
    TableChunk ch = parse_csv_device(ti.ctx, ti.schema, want, (const unsigned char*)text->ptr, bytes, opt);
    if (ti.consolidated && ti.num_batches > 0) {  // re-open: keep the consolidated columns as the first chunk
      TableChunk first;
      first.cols = ti.cols;
      first.rows = ti.num_rows;
      ti.chunks.push_back(first);
    }
    ti.chunks.push_back(ch);
    ti.num_rows += ch.rows;
    ti.num_batches += std::max<int64_t>(1, (ch.rows + 1023) / 1024);  // the reference's reader yields 1024-row batches (csv.rs:63-66)
    ti.consolidated = false;
    if (out_rows) *out_rows = ch.rows;
  });
}

int qgpu_table_append_csv(qgpu_table* t, const void* text, int64_t len, const qgpu_csv_options* options, const int32_t* upload_columns,
                          int32_t n, int64_t* out_rows) {
  return table_append_csv(t, text, len, nullptr, options, upload_columns, n, out_rows);
}
int qgpu_table_append_csv_file(qgpu_table* t, const char* path, const qgpu_csv_options* options, const int32_t* upload_columns, int32_t n,
                               int64_t* out_rows) {
  return table_append_csv(t, nullptr, 0, path, options, upload_columns, n, out_rows);
}

int qgpu_table_flush(qgpu_table* t) {
  if (!t) return QGPU_ERR_INTERNAL;
  TableImpl& ti = *t->t;
  return guard(ti.ctx, [&] { ti.consolidate(); });
}

int qgpu_set_option(qgpu_ctx* ctx, const char* name, int64_t value) {
  if (!ctx || !name) return QGPU_ERR_INTERNAL;
  return guard(&ctx->c, [&] {
    std::string s(name);
    if (s == "ingest_threads") {
      if (value < 0 || value > 16) throw_internal("ingest_threads must be 0 (automatic) .. 16");
      ctx->c.ingest_threads = (int)value;
    } else if (s == "pool_reserve_mb") {
      // grow the stream-ordered pool now (it keeps what it once reserved): later executions then never wait for the driver
      // to map more physical memory in the middle of a step
      if (value < 0 || value > (1 << 20)) throw_internal("pool_reserve_mb out of range");
      ctx->c.pool_floor_bytes = (size_t)value << 20;
      ctx->c.reserve_pool();
    } else if (s == "ingest_host_narrow") {
      ctx->c.ingest_host_narrow = value < 0 ? -1 : (value != 0);
    } else {
      throw_internal("unknown option '" + s + "'");
    }
  });
}

int64_t qgpu_table_num_rows(const qgpu_table* t) {
  if (!t) return -1;
  if (t->t->pending) {  // result of an asynchronous execute: wait for its metadata (errors of the producing kernels surface here)
    if (guard(t->t->ctx, [&] { t->t->resolve(); }) != QGPU_OK) return -1;
  }
  return t->t->num_rows;
}

int qgpu_table_wait(qgpu_table* t) {
  if (!t) return QGPU_ERR_INTERNAL;
  return guard(t->t->ctx, [&] { t->t->resolve(); });
}
int64_t qgpu_table_num_batches(const qgpu_table* t) { return t ? t->t->num_batches : -1; }

int64_t qgpu_table_column_bytes(const qgpu_table* t, int32_t col) {
  if (!t) return -1;
  TableImpl& ti = *t->t;
  int64_t r = -1;
  guard(ti.ctx, [&] {
    ti.consolidate();
    if (col < 0 || col >= (int32_t)ti.cols.size()) throw_internal("column index out of range");
    if (!ti.cols[col]) {
      r = 0;
      return;
    }
    const DCol& c = *ti.cols[col];
    int64_t b = 0;
    if (c.phys == PH_STR && c.dict_state == 1) b = c.length;  // dictionary codes are what the kernels stream
    else if (c.phys == PH_STR) b = (c.length + 1) * 4 + c.str_bytes;
    else if (c.phys == PH_BIT) b = (c.length + 7) / 8;
    else b = c.length * phys_width(c.phys);
    if (c.validity) b += (c.length + 7) / 8;
    r = b;
  });
  return r;
}

int qgpu_table_schema(const qgpu_table* t, struct ArrowSchema* out) {
  if (!t || !out) return QGPU_ERR_INTERNAL;
  return guard(t->t->ctx, [&] { export_schema(t->t->schema, out); });
}

int qgpu_table_export(qgpu_table* t, struct ArrowArray* out_array, struct ArrowSchema* out_schema) {
  if (!t || !out_array || !out_schema) return QGPU_ERR_INTERNAL;
  TableImpl& ti = *t->t;
  return guard(ti.ctx, [&] {
    ti.consolidate();
    View v;
    v.schema = ti.schema;
    v.num_rows = ti.num_rows;
    for (auto& c : ti.cols) v.cols.push_back({c, nullptr});
    std::vector<DColP> cols = materialize_view(ti.ctx, v);
    export_batch(ti.ctx, ti.schema, cols, ti.num_rows, out_array);
    export_schema(ti.schema, out_schema);
  });
}

int qgpu_table_hash_partition(qgpu_table* t, int32_t key_col, int32_t n_parts, qgpu_table** out, int64_t* offsets) {
  if (!t || !out || !offsets) return QGPU_ERR_INTERNAL;
  TableImpl& ti = *t->t;
  return guard(ti.ctx, [&] {
    std::vector<int64_t> offs;
    auto r = hash_partition_table(ti, key_col, n_parts, offs);
    for (size_t i = 0; i < offs.size(); ++i) offsets[i] = offs[i];
    *out = new qgpu_table{r};
  });
}

int qgpu_table_column_device_buffer(qgpu_table* t, int32_t col, void** ptr, int64_t* bytes, int32_t* value_width) {
  if (!t || !ptr || !bytes) return QGPU_ERR_INTERNAL;
  TableImpl& ti = *t->t;
  return guard(ti.ctx, [&] {
    ti.consolidate();
    if (col < 0 || col >= (int32_t)ti.cols.size() || !ti.cols[col]) throw_internal("column is not resident");
    const DCol& c = *ti.cols[col];
    const int w = phys_width(c.phys);
    if (w == 0) throw_internal("column has no fixed-width value buffer");
    if (c.phys == PH_D64) throw_internal("column is resident as a narrowed decimal: it has no Arrow-layout value buffer");
    if (c.null_count != 0) throw_internal("column has NULLs: the value buffer alone does not describe it");
    *ptr = c.data->ptr;
    *bytes = c.length * w;
    if (value_width) *value_width = w;
  });
}

void qgpu_table_free(qgpu_table* t) {
  if (!t) return;
  {
    qgpu::Ctx* c = t->t->ctx;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    cudaSetDevice(c->device);
    c->trace("(before table_free)");
    t->t.reset();
    c->trace("table_free");
  }
  delete t;
}

// ---- expressions ---------------------------------------------------------------------------------
int qgpu_expr_parse(qgpu_ctx* ctx, const uint8_t* ir, size_t len, qgpu_expr** out) {
  if (!ctx || !ir || !out) return QGPU_ERR_INTERNAL;
  return guard(&ctx->c, [&] {
    auto root = parse_ir(ir, len);
    auto* e = new qgpu_expr();
    e->ctx = &ctx->c;
    e->root = std::shared_ptr<ExprNode>(root.release());
    *out = e;
  });
}
void qgpu_expr_free(qgpu_expr* e) { delete e; }

// ---- plan nodes ------------------------------------------------------------------------------------
static std::shared_ptr<PlanNode> new_node(qgpu_ctx* ctx, int kind) {
  auto n = std::make_shared<PlanNode>();
  n->ctx = &ctx->c;
  n->kind = kind;
  return n;
}

int qgpu_plan_scan(qgpu_ctx* ctx, qgpu_table* datasource, const int32_t* projection, int32_t n_projection,
                   const qgpu_expr* filter, qgpu_plan** out) {
  if (!ctx || !datasource || !out) return QGPU_ERR_INTERNAL;
  return guard(&ctx->c, [&] {
    auto n = new_node(ctx, PK_SCAN);
    n->table = datasource->t;
    if (projection) {
      n->has_projection = true;
      n->projection.assign(projection, projection + n_projection);
      for (int ci : n->projection) {
        if (ci < 0 || ci >= (int)n->table->schema.fields.size()) throw_arrow("Schema error: projection index out of bounds");
        n->schema.fields.push_back(n->table->schema.fields[ci]);
      }
      n->schema.metadata = n->table->schema.metadata;
    } else {
      n->schema = n->table->schema;
    }
    if (filter) n->predicate = filter->root;
    *out = new qgpu_plan{n};
  });
}

int qgpu_plan_filter(qgpu_ctx* ctx, qgpu_plan* input, const qgpu_expr* predicate, qgpu_plan** out) {
  if (!ctx || !input || !predicate || !out) return QGPU_ERR_INTERNAL;
  return guard(&ctx->c, [&] {
    auto n = new_node(ctx, PK_FILTER);
    n->children.push_back(input->node);
    n->schema = input->node->schema;  // filter.rs:24-26
    n->predicate = predicate->root;
    *out = new qgpu_plan{n};
  });
}

int qgpu_plan_sort(qgpu_ctx* ctx, qgpu_plan* input, const qgpu_expr* const* exprs, const int32_t* descending, const int32_t* nulls_first,
                   int32_t n_exprs, int64_t limit, qgpu_plan** out) {
  if (!ctx || !input || !out || (n_exprs > 0 && (!exprs || !descending || !nulls_first))) return QGPU_ERR_INTERNAL;
  return guard(&ctx->c, [&] {
    auto n = new_node(ctx, PK_SORT);
    n->children.push_back(input->node);
    n->schema = input->node->schema;  // sort.rs:44-46
    for (int i = 0; i < n_exprs; ++i) {
      n->exprs.push_back(exprs[i]->root);
      n->sort_desc.push_back(descending[i] ? 1 : 0);
      n->sort_nulls_first.push_back(nulls_first[i] ? 1 : 0);
    }
    n->sort_limit = limit;
    *out = new qgpu_plan{n};
  });
}

int qgpu_plan_limit(qgpu_ctx* ctx, qgpu_plan* input, int64_t fetch, int64_t skip, qgpu_plan** out) {
  if (!ctx || !input || !out || skip < 0) return QGPU_ERR_INTERNAL;
  return guard(&ctx->c, [&] {
    auto n = new_node(ctx, PK_LIMIT);
    n->children.push_back(input->node);
    n->schema = input->node->schema;  // limit.rs:23-25
    n->limit_fetch = fetch;
    n->limit_skip = skip;
    *out = new qgpu_plan{n};
  });
}

int qgpu_plan_projection(qgpu_ctx* ctx, const struct ArrowSchema* schema, qgpu_plan* input, const qgpu_expr* const* exprs,
                         int32_t n_exprs, qgpu_plan** out) {
  if (!ctx || !schema || !input || !out || (n_exprs > 0 && !exprs)) return QGPU_ERR_INTERNAL;
  return guard(&ctx->c, [&] {
    auto n = new_node(ctx, PK_PROJECTION);
    n->children.push_back(input->node);
    n->schema = import_schema(schema);
    for (int i = 0; i < n_exprs; ++i) n->exprs.push_back(exprs[i]->root);
    *out = new qgpu_plan{n};
  });
}

int qgpu_plan_aggregate(qgpu_ctx* ctx, const struct ArrowSchema* schema, qgpu_plan* input, const qgpu_expr* const* group_exprs,
                        int32_t n_group, const qgpu_agg_desc* aggs, int32_t n_aggs, qgpu_plan** out) {
  if (!ctx || !schema || !input || !out || (n_group > 0 && !group_exprs) || (n_aggs > 0 && !aggs)) return QGPU_ERR_INTERNAL;
  return guard(&ctx->c, [&] {
    auto n = new_node(ctx, PK_AGGREGATE);
    n->children.push_back(input->node);
    n->schema = import_schema(schema);
    for (int i = 0; i < n_group; ++i) n->group_exprs.push_back(group_exprs[i]->root);
    for (int i = 0; i < n_aggs; ++i) {
      AggDesc d;
      d.op = aggs[i].op;
      if (d.op < 0 || d.op > QGPU_AGG_COUNT) throw_internal("unknown aggregate operator");
      if (!aggs[i].expr) throw_internal("aggregate expression missing");
      d.expr = aggs[i].expr->root;
      d.return_type = mk_type(aggs[i].return_type.id, aggs[i].return_type.precision, aggs[i].return_type.scale);
      d.expr_type = mk_type(aggs[i].expr_type.id, aggs[i].expr_type.precision, aggs[i].expr_type.scale);
      n->aggs.push_back(d);
    }
    *out = new qgpu_plan{n};
  });
}

int qgpu_plan_hash_join(qgpu_ctx* ctx, qgpu_plan* left, qgpu_plan* right, int32_t join_type, const qgpu_expr* const* left_on,
                        const qgpu_expr* const* right_on, int32_t n_on, const qgpu_join_filter* filter, qgpu_plan** out) {
  if (!ctx || !left || !right || !out) return QGPU_ERR_INTERNAL;
  return guard(&ctx->c, [&] {
    if (n_on <= 0 || !left_on || !right_on) throw_internal("On constraints in HashJoinExec should be non-empty");
    if (join_type < 0 || join_type > QGPU_JOIN_LEFT_ANTI) throw_internal("unknown join type");
    auto n = new_node(ctx, PK_HASH_JOIN);
    n->children.push_back(left->node);
    n->children.push_back(right->node);
    n->join_type = join_type;
    n->schema = build_join_schema(left->node->schema, right->node->schema, join_type);
    for (int i = 0; i < n_on; ++i) {
      n->left_on.push_back(left_on[i]->root);
      n->right_on.push_back(right_on[i]->root);
    }
    if (filter) {
      n->has_join_filter = true;
      n->join_filter_expr = filter->expr->root;
      n->join_filter_schema = import_schema(filter->schema);
      n->join_filter_index.assign(filter->column_index, filter->column_index + filter->n_columns);
      n->join_filter_side.assign(filter->column_side, filter->column_side + filter->n_columns);
      if ((int)n->join_filter_schema.fields.size() != filter->n_columns) throw_internal("join filter schema/column_indices mismatch");
    }
    *out = new qgpu_plan{n};
  });
}

int qgpu_plan_nested_loop_join(qgpu_ctx* ctx, qgpu_plan* left, qgpu_plan* right, int32_t join_type, const qgpu_join_filter* filter,
                               qgpu_plan** out) {
  if (!ctx || !left || !right || !out) return QGPU_ERR_INTERNAL;
  return guard(&ctx->c, [&] {
    if (join_type < 0 || join_type > QGPU_JOIN_LEFT_ANTI) throw_internal("unknown join type");
    auto n = new_node(ctx, PK_NL_JOIN);
    n->children.push_back(left->node);
    n->children.push_back(right->node);
    n->join_type = join_type;
    n->schema = build_join_schema(left->node->schema, right->node->schema, join_type);
    if (filter) {
      n->has_join_filter = true;
      n->join_filter_expr = filter->expr->root;
      n->join_filter_schema = import_schema(filter->schema);
      n->join_filter_index.assign(filter->column_index, filter->column_index + filter->n_columns);
      n->join_filter_side.assign(filter->column_side, filter->column_side + filter->n_columns);
      if ((int)n->join_filter_schema.fields.size() != filter->n_columns) throw_internal("join filter schema/column_indices mismatch");
    }
    *out = new qgpu_plan{n};
  });
}

int qgpu_plan_cross_join(qgpu_ctx* ctx, qgpu_plan* left, qgpu_plan* right, qgpu_plan** out) {
  if (!ctx || !left || !right || !out) return QGPU_ERR_INTERNAL;
  return guard(&ctx->c, [&] {
    auto n = new_node(ctx, PK_CROSS_JOIN);
    n->children.push_back(left->node);
    n->children.push_back(right->node);
    n->join_type = QGPU_JOIN_INNER;
    n->schema = build_join_schema(left->node->schema, right->node->schema, QGPU_JOIN_INNER);  // fields and qualifiers concatenated
    *out = new qgpu_plan{n};
  });
}

int qgpu_plan_broadcast(qgpu_ctx* ctx, qgpu_plan* child, int32_t order_free, qgpu_plan** out) {
  if (!ctx || !child || !out) return QGPU_ERR_INTERNAL;
  return guard(&ctx->c, [&] {
    auto n = new_node(ctx, PK_BROADCAST);
    n->children.push_back(child->node);
    n->schema = child->node->schema;
    n->order_free = order_free != 0;
    *out = new qgpu_plan{n};
  });
}

int qgpu_plan_broadcast_pruned(qgpu_ctx* ctx, qgpu_plan* child, int32_t key_column, qgpu_table* probe_table, int32_t probe_key_column,
                               qgpu_plan** out) {
  if (!ctx || !child || !probe_table || !out) return QGPU_ERR_INTERNAL;
  return guard(&ctx->c, [&] {
    if (key_column < 0 || key_column >= (int32_t)child->node->schema.fields.size()) throw_internal("Broadcast: key column out of range");
    if (probe_key_column < 0 || probe_key_column >= (int32_t)probe_table->t->schema.fields.size()) throw_internal("Broadcast: probe key column out of range");
    auto n = new_node(ctx, PK_BROADCAST);
    n->children.push_back(child->node);
    n->schema = child->node->schema;
    n->order_free = true;
    n->prune_table = probe_table->t;
    n->prune_key_col = key_column;
    n->prune_probe_col = probe_key_column;
    *out = new qgpu_plan{n};
  });
}

int qgpu_plan_final_aggregate(qgpu_ctx* ctx, qgpu_plan* child, const int32_t* key_columns, int32_t n_keys, const int32_t* value_columns,
                              const int32_t* merge_ops, int32_t n_values, qgpu_plan** out) {
  if (!ctx || !child || !out || !key_columns || n_keys <= 0 || n_values < 0 || (n_values > 0 && (!value_columns || !merge_ops)))
    return QGPU_ERR_INTERNAL;
  return guard(&ctx->c, [&] {
    auto n = new_node(ctx, PK_FINAL_AGG);
    n->children.push_back(child->node);
    n->schema = child->node->schema;
    n->exchange_keys.assign(key_columns, key_columns + n_keys);
    n->exchange_cols.assign(value_columns, value_columns + n_values);
    n->exchange_ops.assign(merge_ops, merge_ops + n_values);
    for (int op : n->exchange_ops)
      if (op < 0 || op > 2) throw_internal("FinalAggregate: merge operator must be 0 (SUM), 1 (MIN) or 2 (MAX)");
    *out = new qgpu_plan{n};
  });
}

int qgpu_plan_schema(const qgpu_plan* p, struct ArrowSchema* out) {
  if (!p || !out) return QGPU_ERR_INTERNAL;
  return guard(p->node->ctx, [&] { export_schema(p->node->schema, out); });
}

namespace {
struct Timer {
  Ctx* ctx;
  cudaEvent_t a, b;
  int64_t l0;
  explicit Timer(Ctx* c) : ctx(c) {
    CUDA_CHECK(cudaEventCreate(&a));
    CUDA_CHECK(cudaEventCreate(&b));
    l0 = c->launches;
    CUDA_CHECK(cudaEventRecord(a, c->stream));
  }
  void stop(PlanNode& n) {
    CUDA_CHECK(cudaEventRecord(b, ctx->stream));
    CUDA_CHECK(cudaEventSynchronize(b));
    float ms = 0;
    CUDA_CHECK(cudaEventElapsedTime(&ms, a, b));
    n.last_ms = ms;
    n.last_launches = ctx->launches - l0;
  }
  ~Timer() {
    cudaEventDestroy(a);
    cudaEventDestroy(b);
  }
};
}  // namespace

int qgpu_plan_execute(qgpu_plan* p, struct ArrowArrayStream* out) {
  if (!p || !out) return QGPU_ERR_INTERNAL;
  PlanNode& n = *p->node;
  return guard(n.ctx, [&] {
    Timer tm(n.ctx);
    View v = n.execute();
    v.resolve();
    tm.stop(n);
    std::vector<ArrowArray> batches;
    if (v.num_batches > 0) {
      std::vector<DColP> cols = materialize_view(n.ctx, v);
      batches.resize(1);
      export_batch(n.ctx, v.schema, cols, v.num_rows, &batches[0]);
    }
    make_stream(n.schema, std::move(batches), out);
  });
}

int qgpu_plan_state_bytes(qgpu_plan* p, int32_t max_groups, int64_t* bytes) {
  if (!p || !bytes) return QGPU_ERR_INTERNAL;
  return guard(p->node->ctx, [&] { *bytes = shard_state_bytes(*p->node, max_groups); });
}

int qgpu_plan_partial_state(qgpu_plan* p, int64_t row_offset, int32_t max_groups, void* device_buf, int64_t cap_bytes) {
  if (!p || !device_buf) return QGPU_ERR_INTERNAL;
  PlanNode& n = *p->node;
  return guard(n.ctx, [&] {
    Timer tm(n.ctx);
    shard_partial_state(n, row_offset, max_groups, device_buf, cap_bytes);
    tm.stop(n);
  });
}

int qgpu_plan_execute_merged(qgpu_plan* p, const void* gathered, int32_t n_states, int32_t max_groups, struct ArrowArrayStream* out) {
  if (!p || !gathered || !out) return QGPU_ERR_INTERNAL;
  PlanNode& n = *p->node;
  return guard(n.ctx, [&] {
    View v = shard_execute_merged(n, gathered, n_states, max_groups);
    v.resolve();
    std::vector<ArrowArray> batches;
    if (v.num_batches > 0) {
      std::vector<DColP> cols = materialize_view(n.ctx, v);
      batches.resize(1);
      export_batch(n.ctx, v.schema, cols, v.num_rows, &batches[0]);
    }
    make_stream(n.schema, std::move(batches), out);
  });
}

int qgpu_plan_set_order_free(qgpu_plan* p, int32_t on) {
  if (!p) return QGPU_ERR_INTERNAL;
  // the flag is inherited by the operators below that preserve order (Projection / Filter) down to the first join
  PlanNode* n = p->node.get();
  while (n && (n->kind == PK_PROJECTION || n->kind == PK_FILTER)) n = n->children[0].get();
  if (n) n->order_free = on != 0;
  return QGPU_OK;
}

int qgpu_plan_execute_merged_device(qgpu_plan* p, const void* gathered, int32_t n_states, int32_t max_groups, qgpu_table** out) {
  if (!p || !gathered || !out) return QGPU_ERR_INTERNAL;
  PlanNode& n = *p->node;
  return guard(n.ctx, [&] {
    View v = shard_execute_merged(n, gathered, n_states, max_groups);
    v.resolve();
    auto t = std::make_shared<TableImpl>();
    t->ctx = n.ctx;
    t->schema = n.schema;
    t->num_rows = v.num_rows;
    t->num_batches = v.num_batches;
    for (size_t i = 0; i < v.cols.size(); ++i) {
      if (!v.cols[i].base) t->cols.push_back(nullptr);
      else t->cols.push_back(materialize(n.ctx, v.cols[i], v.num_rows));
    }
    *out = new qgpu_table{t};
  });
}

int qgpu_plan_exchange_keystats(qgpu_plan* p, int64_t* stats, int32_t* n_keys) {
  if (!p || !stats || !n_keys) return QGPU_ERR_INTERNAL;
  PlanNode& n = *p->node;
  return guard(n.ctx, [&] { radix_exchange_keystats(n, stats, n_keys); });
}

int qgpu_plan_exchange_sketch(qgpu_plan* p, const int64_t* global_stats, void** device_buf, int64_t* bytes, int32_t* eligible) {
  if (!p || !global_stats || !device_buf || !bytes || !eligible) return QGPU_ERR_INTERNAL;
  PlanNode& n = *p->node;
  return guard(n.ctx, [&] { *eligible = radix_exchange_sketch(n, global_stats, device_buf, bytes) == 0 ? 1 : 0; });
}

int qgpu_plan_exchange_prepare(qgpu_plan* p, const void* gathered_host, int32_t world, int32_t rank, void* handles_out,
                               int32_t* n_handles, int32_t* eligible) {
  if (!p || !gathered_host || !handles_out || !n_handles || !eligible) return QGPU_ERR_INTERNAL;
  PlanNode& n = *p->node;
  return guard(n.ctx, [&] { *eligible = radix_exchange_prepare(n, gathered_host, world, rank, handles_out, n_handles) == 0 ? 1 : 0; });
}

int qgpu_plan_exchange_scatter(qgpu_plan* p, const void* all_handles_host) {
  if (!p || !all_handles_host) return QGPU_ERR_INTERNAL;
  PlanNode& n = *p->node;
  return guard(n.ctx, [&] { radix_exchange_scatter(n, all_handles_host); });
}

int qgpu_plan_exchange_finish(qgpu_plan* p, int32_t* overflow) {
  if (!p || !overflow) return QGPU_ERR_INTERNAL;
  PlanNode& n = *p->node;
  return guard(n.ctx, [&] { *overflow = radix_exchange_finish(n); });
}

// the result View of a plan as an HBM-resident table; `async`: the metadata may still be pending
static std::shared_ptr<TableImpl> view_to_table(PlanNode& n, View& v, bool async) {
  if (!async) v.resolve();
  auto t = std::make_shared<TableImpl>();
  t->ctx = n.ctx;
  t->schema = n.schema;
  t->num_rows = v.num_rows;
  t->num_batches = v.num_batches;
  for (size_t i = 0; i < v.cols.size(); ++i) {
    if (!v.cols[i].base) t->cols.push_back(nullptr);
    else t->cols.push_back(materialize(n.ctx, v.cols[i], v.num_rows));  // pending results carry no index vectors: handles only
  }
  t->pending = v.pending;
  return t;
}

int qgpu_plan_execute_device(qgpu_plan* p, qgpu_table** out, int64_t* out_batches) {
  if (!p || !out) return QGPU_ERR_INTERNAL;
  PlanNode& n = *p->node;
  return guard(n.ctx, [&] {
    n.ctx->trace("(between calls)");
    Timer tm(n.ctx);
    View v = n.execute();
    n.ctx->trace("execute_device: plan");
    auto t = view_to_table(n, v, false);
    n.ctx->trace("execute_device: materialize result");
    tm.stop(n);
    if (out_batches) *out_batches = v.num_batches;
    *out = new qgpu_table{t};
  });
}

int qgpu_plan_execute_device_async(qgpu_plan* p, qgpu_table** out) {
  if (!p || !out) return QGPU_ERR_INTERNAL;
  PlanNode& n = *p->node;
  return guard(n.ctx, [&] {
    const int64_t l0 = n.ctx->launches;
    struct AsyncOk {
      Ctx* c;
      explicit AsyncOk(Ctx* x) : c(x) { c->async_ok = true; }
      ~AsyncOk() { c->async_ok = false; }
    } scope(n.ctx);
    View v = n.execute();
    auto t = view_to_table(n, v, true);
    n.last_launches = n.ctx->launches - l0;
    *out = new qgpu_table{t};
  });
}

// ---- multi-GPU: communicator + sharded execution below the ABI ---------------------------------------------------------------
int qgpu_comm_unique_id(void* out, int64_t cap) {
  if (!out || cap < 128) return QGPU_ERR_INTERNAL;
  try {
    comm_unique_id(out);
    return QGPU_OK;
  } catch (QError& e) {
    g_last_error = e.what();
    return e.code;
  }
}

int qgpu_comm_init(qgpu_ctx* ctx, const void* unique_id, int32_t rank, int32_t world) {
  if (!ctx || !unique_id) return QGPU_ERR_INTERNAL;
  return guard(&ctx->c, [&] { comm_init(&ctx->c, unique_id, rank, world); });
}

int qgpu_comm_init_local(qgpu_ctx** ctxs, int32_t n) {
  if (!ctxs || n < 1) return QGPU_ERR_INTERNAL;
  return guard(&ctxs[0]->c, [&] {
    std::vector<Ctx*> cs;
    for (int i = 0; i < n; ++i) cs.push_back(&ctxs[i]->c);
    comm_init_local(cs.data(), n);
  });
}

int qgpu_comm_destroy(qgpu_ctx* ctx) {
  if (!ctx) return QGPU_ERR_INTERNAL;
  return guard(&ctx->c, [&] { comm_destroy(&ctx->c); });
}

int qgpu_comm_world(const qgpu_ctx* ctx, int32_t* rank, int32_t* world) {
  if (!ctx) return QGPU_ERR_INTERNAL;
  if (rank) *rank = ctx->c.comm ? ctx->c.comm->rank : 0;
  if (world) *world = ctx->c.comm ? ctx->c.comm->world : 1;
  return QGPU_OK;
}

int qgpu_comm_all_gather(qgpu_ctx* ctx, const void* send_device, void* recv_device, int64_t bytes_per_rank) {
  if (!ctx || !send_device || !recv_device || bytes_per_rank < 0) return QGPU_ERR_INTERNAL;
  return guard(&ctx->c, [&] { comm_all_gather(&ctx->c, send_device, recv_device, (size_t)bytes_per_rank); });
}

int qgpu_comm_all_to_all(qgpu_ctx* ctx, const void* send_device, const int64_t* send_offsets, const int64_t* send_bytes, void* recv_device,
                         const int64_t* recv_offsets, const int64_t* recv_bytes) {
  if (!ctx || !send_offsets || !send_bytes || !recv_offsets || !recv_bytes) return QGPU_ERR_INTERNAL;
  return guard(&ctx->c, [&] { comm_all_to_all(&ctx->c, send_device, send_offsets, send_bytes, recv_device, recv_offsets, recv_bytes); });
}

int qgpu_comm_barrier(qgpu_ctx* ctx) {
  if (!ctx) return QGPU_ERR_INTERNAL;
  return guard(&ctx->c, [&] { comm_barrier(&ctx->c); });
}

int qgpu_plan_execute_sharded(qgpu_plan* p, int64_t row_offset, int32_t max_groups, struct ArrowArrayStream* out) {
  if (!p || !out) return QGPU_ERR_INTERNAL;
  PlanNode& n = *p->node;
  return guard(n.ctx, [&] {
    Timer tm(n.ctx);
    View v = shard_execute_fused(n, row_offset, max_groups);
    v.resolve();
    tm.stop(n);
    std::vector<ArrowArray> batches;
    if (v.num_batches > 0) {
      std::vector<DColP> cols = materialize_view(n.ctx, v);
      batches.resize(1);
      export_batch(n.ctx, v.schema, cols, v.num_rows, &batches[0]);
    }
    make_stream(n.schema, std::move(batches), out);
  });
}

int qgpu_plan_execute_sharded_device(qgpu_plan* p, int64_t row_offset, int32_t max_groups, int32_t async, qgpu_table** out) {
  if (!p || !out) return QGPU_ERR_INTERNAL;
  PlanNode& n = *p->node;
  return guard(n.ctx, [&] {
    const int64_t l0 = n.ctx->launches;
    View v = shard_execute_fused(n, row_offset, max_groups);
    auto t = view_to_table(n, v, async != 0);
    n.last_launches = n.ctx->launches - l0;
    *out = new qgpu_table{t};
  });
}

int qgpu_plan_last_stats(const qgpu_plan* p, double* device_ms, int64_t* launches) {
  if (!p) return QGPU_ERR_INTERNAL;
  if (device_ms) *device_ms = p->node->last_ms;
  if (launches) *launches = p->node->last_launches;
  return QGPU_OK;
}

static std::string describe_strategy(const PlanNode& n) {
  std::string s = n.strategy;
  if (!n.children.empty() && n.strategy.find("fused") == std::string::npos) {
    s += " <- ";
    if (n.children.size() > 1) s += "[";
    for (size_t i = 0; i < n.children.size(); ++i) s += (i ? ", " : "") + describe_strategy(*n.children[i]);
    if (n.children.size() > 1) s += "]";
  }
  return s;
}

const char* qgpu_plan_strategy(const qgpu_plan* p) {
  if (!p) return "";
  p->node->strategy_desc = describe_strategy(*p->node);
  return p->node->strategy_desc.c_str();
}

void qgpu_plan_free(qgpu_plan* p) {
  if (!p) return;
  {
    Ctx* c = p->node->ctx;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    cudaSetDevice(c->device);
    p->node.reset();
  }
  delete p;
}

}  // extern "C"
