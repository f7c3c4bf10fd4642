// Arrow C Data Interface import (host or device buffers -> HBM columns) and export (HBM -> host).
// Staging policy (north_star (1)): pinned host sources are DMA'd directly with cudaMemcpyAsync;
// pageable sources go through the context's pinned ring on a side stream (Ctx::h2d).
#include <cstdlib>
#include <cstring>

#include "kernels.h"
#include "qgpu_internal.h"

namespace qgpu {

// ------------------------------------------------------------------------------------------------
// schema
// ------------------------------------------------------------------------------------------------
static size_t metadata_len(const char* md) {
  if (!md) return 0;
  const char* p = md;
  int32_t n;
  memcpy(&n, p, 4);
  p += 4;
  for (int32_t i = 0; i < n; ++i) {
    int32_t l;
    memcpy(&l, p, 4);
    p += 4 + l;
    memcpy(&l, p, 4);
    p += 4 + l;
  }
  return (size_t)(p - md);
}

bool metadata_get(const std::string& blob, const std::string& key, std::string* value) {
  if (blob.size() < 4) return false;
  const char* p = blob.data();
  int32_t n;
  memcpy(&n, p, 4);
  p += 4;
  for (int32_t i = 0; i < n; ++i) {
    int32_t kl, vl;
    memcpy(&kl, p, 4);
    std::string k(p + 4, (size_t)kl);
    p += 4 + kl;
    memcpy(&vl, p, 4);
    if (k == key) {
      value->assign(p + 4, (size_t)vl);
      return true;
    }
    p += 4 + vl;
  }
  return false;
}

std::string merge_metadata(const std::string& base, const std::string& key, const std::string& value) {
  // re-encode all pairs, replacing/adding `key`
  std::vector<std::pair<std::string, std::string>> kv;
  if (base.size() >= 4) {
    const char* p = base.data();
    int32_t n;
    memcpy(&n, p, 4);
    p += 4;
    for (int32_t i = 0; i < n; ++i) {
      int32_t kl, vl;
      memcpy(&kl, p, 4);
      std::string k(p + 4, (size_t)kl);
      p += 4 + kl;
      memcpy(&vl, p, 4);
      std::string v(p + 4, (size_t)vl);
      p += 4 + vl;
      if (k != key) kv.push_back({k, v});
    }
  }
  kv.push_back({key, value});
  std::string out;
  int32_t n = (int32_t)kv.size();
  out.append((const char*)&n, 4);
  for (auto& e : kv) {
    int32_t l = (int32_t)e.first.size();
    out.append((const char*)&l, 4);
    out.append(e.first);
    l = (int32_t)e.second.size();
    out.append((const char*)&l, 4);
    out.append(e.second);
  }
  return out;
}

static DType parse_format(const char* f) {
  std::string s(f ? f : "");
  if (s == "n") return mk_type(QGPU_T_NULL);
  if (s == "b") return mk_type(QGPU_T_BOOL);
  if (s == "c") return mk_type(QGPU_T_INT8);
  if (s == "C") return mk_type(QGPU_T_UINT8);
  if (s == "s") return mk_type(QGPU_T_INT16);
  if (s == "S") return mk_type(QGPU_T_UINT16);
  if (s == "i") return mk_type(QGPU_T_INT32);
  if (s == "I") return mk_type(QGPU_T_UINT32);
  if (s == "l") return mk_type(QGPU_T_INT64);
  if (s == "L") return mk_type(QGPU_T_UINT64);
  if (s == "f") return mk_type(QGPU_T_FLOAT32);
  if (s == "g") return mk_type(QGPU_T_FLOAT64);
  if (s == "u") return mk_type(QGPU_T_UTF8);
  if (s == "tdD") return mk_type(QGPU_T_DATE32);
  if (s == "tdm") return mk_type(QGPU_T_DATE64);
  if (s == "tts") return mk_type(QGPU_T_TIME32, 0, 0);
  if (s == "ttm") return mk_type(QGPU_T_TIME32, 0, 1);
  if (s == "ttu") return mk_type(QGPU_T_TIME64, 0, 2);
  if (s == "ttn") return mk_type(QGPU_T_TIME64, 0, 3);
  if (s.rfind("d:", 0) == 0) {
    int p = 0, sc = 0, bits = 128;
    int k = sscanf(s.c_str(), "d:%d,%d,%d", &p, &sc, &bits);
    if (k >= 2 && bits == 128 && p >= 1 && p <= 38) return mk_type(QGPU_T_DECIMAL128, p, sc);
  }
  throw_internal("Unsupported Arrow data type (format '" + s + "') on the GPU path");
}

static std::string format_of(const DType& t) {
  switch (t.id) {
    case QGPU_T_NULL: return "n";
    case QGPU_T_BOOL: return "b";
    case QGPU_T_INT8: return "c";
    case QGPU_T_UINT8: return "C";
    case QGPU_T_INT16: return "s";
    case QGPU_T_UINT16: return "S";
    case QGPU_T_INT32: return "i";
    case QGPU_T_UINT32: return "I";
    case QGPU_T_INT64: return "l";
    case QGPU_T_UINT64: return "L";
    case QGPU_T_FLOAT32: return "f";
    case QGPU_T_FLOAT64: return "g";
    case QGPU_T_UTF8: return "u";
    case QGPU_T_DATE32: return "tdD";
    case QGPU_T_DATE64: return "tdm";
    case QGPU_T_TIME32: return t.scale == 1 ? "ttm" : "tts";
    case QGPU_T_TIME64: return t.scale == 3 ? "ttn" : "ttu";
    case QGPU_T_DECIMAL128: return "d:" + std::to_string(t.precision) + "," + std::to_string(t.scale);
  }
  return "n";
}

Schema import_schema(const ArrowSchema* s) {
  if (!s || !s->format || std::string(s->format) != "+s") throw_internal("expected a struct ArrowSchema (RecordBatch schema)");
  Schema out;
  if (s->metadata) out.metadata.assign(s->metadata, metadata_len(s->metadata));
  for (int64_t i = 0; i < s->n_children; ++i) {
    const ArrowSchema* c = s->children[i];
    Field f;
    f.name = c->name ? c->name : "";
    f.type = parse_format(c->format);
    f.nullable = (c->flags & ARROW_FLAG_NULLABLE) != 0;
    if (c->metadata) f.metadata.assign(c->metadata, metadata_len(c->metadata));
    out.fields.push_back(f);
  }
  return out;
}

namespace {
struct SchemaPriv {
  std::string format, name, metadata;
  std::vector<ArrowSchema> children;
  std::vector<ArrowSchema*> child_ptrs;
};
void release_schema(ArrowSchema* s) {
  if (!s || !s->release) return;
  SchemaPriv* p = (SchemaPriv*)s->private_data;
  for (auto& c : p->children)
    if (c.release) c.release(&c);
  delete p;
  s->release = nullptr;
}
void fill_schema(ArrowSchema* out, const std::string& fmt, const std::string& name, const std::string& md, int64_t flags) {
  SchemaPriv* p = new SchemaPriv();
  p->format = fmt;
  p->name = name;
  p->metadata = md;
  memset(out, 0, sizeof(ArrowSchema));
  out->format = p->format.c_str();
  out->name = p->name.c_str();
  out->metadata = p->metadata.empty() ? nullptr : p->metadata.data();
  out->flags = flags;
  out->release = release_schema;
  out->private_data = p;
}
}  // namespace

void export_schema(const Schema& s, ArrowSchema* out) {
  fill_schema(out, "+s", "", s.metadata, 0);
  SchemaPriv* p = (SchemaPriv*)out->private_data;
  p->children.resize(s.fields.size());
  p->child_ptrs.resize(s.fields.size());
  for (size_t i = 0; i < s.fields.size(); ++i) {
    const Field& f = s.fields[i];
    fill_schema(&p->children[i], format_of(f.type), f.name, f.metadata, f.nullable ? ARROW_FLAG_NULLABLE : 0);
    p->child_ptrs[i] = &p->children[i];
  }
  out->n_children = (int64_t)s.fields.size();
  out->children = p->child_ptrs.data();
}

// ------------------------------------------------------------------------------------------------
// import
// ------------------------------------------------------------------------------------------------
static Phys canonical_phys(const DType& t) {
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

static void copy_in(Ctx* ctx, void* dst, const void* src, size_t bytes, bool device_resident) {
  if (bytes == 0) return;
  if (device_resident) CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  else ctx->h2d(dst, src, bytes);
}

// upload n bits starting at bit `bit_off` of `src` into a fresh word-aligned device bitmap
static DBufP upload_bits(Ctx* ctx, const uint8_t* src, int64_t bit_off, int64_t n, bool device_resident) {
  const int64_t n_words = (n + 31) >> 5;
  DBufP b = ctx->alloc_zero(std::max<size_t>((size_t)n_words * 4, 4));
  if (n == 0) return b;
  if ((bit_off & 7) == 0) {
    copy_in(ctx, b->ptr, src + (bit_off >> 3), (size_t)((n + 7) >> 3), device_resident);
  } else if (device_resident) {
    // unaligned device bitmap: stage the covering bytes then shift on the device
    int64_t nbytes = ((bit_off + n + 7) >> 3) - (bit_off >> 3);
    DBufP tmp = ctx->alloc_zero((size_t)nbytes + 8);
    CUDA_CHECK(cudaMemcpyAsync(tmp->ptr, src + (bit_off >> 3), (size_t)nbytes, cudaMemcpyDeviceToDevice, ctx->stream));
    copy_bits(ctx, (uint32_t*)b->ptr, 0, (const uint32_t*)tmp->ptr, bit_off & 7, n);
  } else {
    std::vector<uint8_t> tmp((size_t)((n + 7) >> 3), 0);
    for (int64_t i = 0; i < n; ++i) {
      int64_t s = bit_off + i;
      if ((src[s >> 3] >> (s & 7)) & 1) tmp[(size_t)(i >> 3)] |= (uint8_t)(1u << (i & 7));
    }
    ctx->h2d(b->ptr, tmp.data(), tmp.size());
    ctx->sync();  // tmp is freed on return
  }
  return b;
}

static int64_t host_count_zero_bits(const uint8_t* bits, int64_t bit_off, int64_t n) {
  int64_t ones = 0;
  for (int64_t i = 0; i < n; ++i) {
    int64_t s = bit_off + i;
    ones += (bits[s >> 3] >> (s & 7)) & 1;
  }
  return n - ones;
}

static DColP import_column(Ctx* ctx, const Field& f, const ArrowArray* a, int64_t parent_off, int64_t n, bool dev) {
  auto col = std::make_shared<DCol>();
  col->type = f.type;
  col->length = n;
  col->phys = canonical_phys(f.type);
  const int64_t off = a->offset + parent_off;
  if (f.type.id == QGPU_T_NULL) {
    col->null_count = n;
    return col;
  }
  // Arrow: child.length >= struct.offset + struct.length
  if (a->length < parent_off + n) throw_arrow("child array shorter than the record batch");
  // validity
  const uint8_t* vbits = (a->n_buffers > 0) ? (const uint8_t*)a->buffers[0] : nullptr;
  int64_t nulls = a->null_count;
  if (!vbits) nulls = 0;
  if (vbits && nulls != 0) {
    if (nulls < 0 && !dev) nulls = host_count_zero_bits(vbits, off, n);
    if (nulls != 0) {
      col->validity = upload_bits(ctx, vbits, off, n, dev);
      if (nulls < 0) nulls = n - count_set_bits(ctx, (const uint32_t*)col->validity->ptr, n);
      if (nulls == 0) col->validity.reset();
    }
  }
  col->null_count = nulls;
  const int w = arrow_width(f.type);
  if (col->phys == PH_BIT) {
    col->data = upload_bits(ctx, (const uint8_t*)a->buffers[1], off, n, dev);
  } else if (col->phys == PH_STR) {
    const int32_t* offs = (const int32_t*)a->buffers[1];
    int32_t o0 = 0, o1 = 0;
    col->offsets = ctx->alloc_zero((size_t)(n + 1) * 4);
    if (n > 0 || offs) {
      if (dev) {
        if (n > 0) {
          o0 = ctx->read_scalar(offs + off);
          o1 = ctx->read_scalar(offs + off + n);
        }
      } else if (offs) {
        o0 = offs[off];
        o1 = offs[off + n];
      }
      if (offs) {
        copy_in(ctx, col->offsets->ptr, offs + off, (size_t)(n + 1) * 4, dev);
        if (o0 != 0) rebase_offsets(ctx, (int32_t*)col->offsets->ptr, (const int32_t*)col->offsets->ptr, n + 1, -(int64_t)o0);
      }
    }
    col->str_bytes = (int64_t)o1 - o0;
    col->data = ctx->alloc(std::max<size_t>((size_t)col->str_bytes, 4));
    if (col->str_bytes > 0) copy_in(ctx, col->data->ptr, (const char*)a->buffers[2] + o0, (size_t)col->str_bytes, dev);
  } else {
    col->data = ctx->alloc(std::max<size_t>((size_t)n * w, 16));
    if (n > 0) copy_in(ctx, col->data->ptr, (const char*)a->buffers[1] + off * w, (size_t)n * w, dev);
  }
  // Decimal128 with precision <= 18 always fits int64 by declared type; verify and narrow (8 B/value)
  if (col->phys == PH_I128 && f.type.precision <= 18) {
    DColP nar = try_narrow_decimal(ctx, *col);
    if (nar) return nar;
  }
  return col;
}

TableChunk import_batch(Ctx* ctx, const Schema& schema, ArrowArray* batch, const int32_t* upload_columns, int32_t n_upload,
                        bool device_resident) {
  if (!batch) throw_internal("null batch");
  if (batch->n_children != (int64_t)schema.fields.size())
    throw_arrow("RecordBatch has " + std::to_string(batch->n_children) + " columns but the table schema has " +
                std::to_string(schema.fields.size()));
  TableChunk ch;
  ch.rows = batch->length;
  ch.cols.resize(schema.fields.size());
  std::vector<char> want(schema.fields.size(), upload_columns ? 0 : 1);
  if (upload_columns)
    for (int32_t i = 0; i < n_upload; ++i) {
      if (upload_columns[i] < 0 || upload_columns[i] >= (int32_t)schema.fields.size()) throw_internal("upload column index out of range");
      want[upload_columns[i]] = 1;
    }
  for (size_t i = 0; i < schema.fields.size(); ++i) {
    if (!want[i]) continue;
    ch.cols[i] = import_column(ctx, schema.fields[i], batch->children[i], batch->offset, batch->length, device_resident);
  }
  ctx->sync();  // host buffers may be released by the caller after this returns
  return ch;
}

// ------------------------------------------------------------------------------------------------
// export
// ------------------------------------------------------------------------------------------------
namespace {
struct ArrayPriv {
  std::vector<void*> owned;
  std::vector<const void*> buffers;
  std::vector<ArrowArray> children;
  std::vector<ArrowArray*> child_ptrs;
};
void release_array(ArrowArray* a) {
  if (!a || !a->release) return;
  ArrayPriv* p = (ArrayPriv*)a->private_data;
  for (auto& c : p->children)
    if (c.release) c.release(&c);
  for (void* b : p->owned) free(b);
  delete p;
  a->release = nullptr;
}
void* host_alloc(ArrayPriv* p, size_t bytes) {
  void* b = nullptr;
  if (posix_memalign(&b, 64, ((bytes + 63) / 64) * 64 + 64) != 0) throw QError(QGPU_ERR_OOM, "host allocation failed");
  memset(b, 0, ((bytes + 63) / 64) * 64 + 64);
  p->owned.push_back(b);
  return b;
}
}  // namespace

namespace {
struct PendingCopy {
  void* dst;
  const void* src;
  size_t bytes;
};
}  // namespace

// device->host copies are only RECORDED here; export_batch issues them (small results go through the pinned
// scratch so that every copy is truly asynchronous and the whole batch costs one synchronisation)
static void export_column(Ctx* ctx, const DCol& c, int64_t n, ArrowArray* out, std::vector<PendingCopy>& copies) {
  ArrayPriv* p = new ArrayPriv();
  memset(out, 0, sizeof(ArrowArray));
  out->private_data = p;
  out->release = release_array;
  out->length = n;
  out->null_count = c.phys == PH_NULL ? n : c.null_count;
  out->offset = 0;
  if (c.type.id == QGPU_T_NULL) {
    out->n_buffers = 0;
    out->null_count = n;
    return;
  }
  const size_t vbytes = (size_t)((n + 7) >> 3);
  void* vb = nullptr;
  if (c.phys == PH_NULL) {
    vb = host_alloc(p, vbytes);  // all-zero bitmap: every slot NULL
  } else if (c.validity && c.null_count > 0) {
    vb = host_alloc(p, vbytes + 4);
    copies.push_back({vb, c.validity->ptr, ((vbytes + 3) / 4) * 4});
  }
  p->buffers.push_back(vb);
  const int w = arrow_width(c.type);
  if (c.type.id == QGPU_T_BOOL) {
    void* d = host_alloc(p, vbytes + 4);
    if (c.phys != PH_NULL && n > 0)
      copies.push_back({d, c.data->ptr, ((vbytes + 3) / 4) * 4});
    p->buffers.push_back(d);
  } else if (c.type.id == QGPU_T_UTF8) {
    void* o = host_alloc(p, (size_t)(n + 1) * 4);
    void* d = host_alloc(p, (size_t)std::max<int64_t>(c.str_bytes, 1));
    if (c.phys != PH_NULL) {
      copies.push_back({o, c.offsets->ptr, (size_t)(n + 1) * 4});
      if (c.str_bytes > 0) copies.push_back({d, c.data->ptr, (size_t)c.str_bytes});
    }
    p->buffers.push_back(o);
    p->buffers.push_back(d);
  } else {
    if (c.phys == PH_D64) throw_internal("export: narrowed decimal must be widened first");
    void* d = host_alloc(p, (size_t)n * w);
    if (c.phys != PH_NULL && n > 0) copies.push_back({d, c.data->ptr, (size_t)n * w});
    p->buffers.push_back(d);
  }
  out->n_buffers = (int64_t)p->buffers.size();
  out->buffers = p->buffers.data();
}

void export_batch(Ctx* ctx, const Schema& schema, const std::vector<DColP>& cols, int64_t num_rows, ArrowArray* out) {
  ArrayPriv* p = new ArrayPriv();
  memset(out, 0, sizeof(ArrowArray));
  out->private_data = p;
  out->release = release_array;
  out->length = num_rows;
  out->null_count = 0;
  p->buffers.push_back(nullptr);
  out->n_buffers = 1;
  out->buffers = p->buffers.data();
  p->children.resize(cols.size());
  p->child_ptrs.resize(cols.size());
  for (size_t i = 0; i < cols.size(); ++i) {
    memset(&p->children[i], 0, sizeof(ArrowArray));
    p->child_ptrs[i] = &p->children[i];
  }
  out->n_children = (int64_t)cols.size();
  out->children = p->child_ptrs.data();
  try {
    std::vector<PendingCopy> copies;
    for (size_t i = 0; i < cols.size(); ++i) export_column(ctx, *cols[i], num_rows, &p->children[i], copies);
    size_t total = 0;
    for (auto& c : copies) total += ((c.bytes + 15) / 16) * 16;
    if (total <= ctx->pinned_scratch_bytes) {
      size_t off = 0;
      for (auto& c : copies) {
        CUDA_CHECK(cudaMemcpyAsync((char*)ctx->pinned_scratch + off, c.src, c.bytes, cudaMemcpyDeviceToHost, ctx->stream));
        off += ((c.bytes + 15) / 16) * 16;
      }
      ctx->sync();
      off = 0;
      for (auto& c : copies) {
        memcpy(c.dst, (char*)ctx->pinned_scratch + off, c.bytes);
        off += ((c.bytes + 15) / 16) * 16;
      }
    } else {
      for (auto& c : copies) CUDA_CHECK(cudaMemcpyAsync(c.dst, c.src, c.bytes, cudaMemcpyDeviceToHost, ctx->stream));
      ctx->sync();
    }
  } catch (...) {
    release_array(out);
    throw;
  }
}

namespace {
struct StreamPriv {
  Schema schema;
  std::vector<ArrowArray> batches;
  size_t next = 0;
  std::string err;
};
int stream_get_schema(ArrowArrayStream* s, ArrowSchema* out) {
  StreamPriv* p = (StreamPriv*)s->private_data;
  try {
    export_schema(p->schema, out);
    return 0;
  } catch (std::exception& e) {
    p->err = e.what();
    return 5;
  }
}
int stream_get_next(ArrowArrayStream* s, ArrowArray* out) {
  StreamPriv* p = (StreamPriv*)s->private_data;
  if (p->next >= p->batches.size()) {
    memset(out, 0, sizeof(ArrowArray));  // release == NULL marks end of stream
    return 0;
  }
  *out = p->batches[p->next];
  p->batches[p->next].release = nullptr;  // moved
  p->next++;
  return 0;
}
const char* stream_last_error(ArrowArrayStream* s) { return ((StreamPriv*)s->private_data)->err.c_str(); }
void stream_release(ArrowArrayStream* s) {
  if (!s || !s->release) return;
  StreamPriv* p = (StreamPriv*)s->private_data;
  for (auto& b : p->batches)
    if (b.release) b.release(&b);
  delete p;
  s->release = nullptr;
}
}  // namespace

void make_stream(const Schema& schema, std::vector<ArrowArray>&& batches, ArrowArrayStream* out) {
  StreamPriv* p = new StreamPriv();
  p->schema = schema;
  p->batches = std::move(batches);
  out->get_schema = stream_get_schema;
  out->get_next = stream_get_next;
  out->get_last_error = stream_last_error;
  out->release = stream_release;
  out->private_data = p;
}

}  // namespace qgpu
