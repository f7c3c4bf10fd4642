// Context plumbing + the generic kernels: interpreter-driven evaluation, predicate -> selection
// vector (ballot/popc stream compaction), gathers ("take"), prefix scan, ingest helpers.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>

#include "kernels.h"
#include "launch.h"

namespace qgpu {

// ================================================================================================
// Ctx / DBuf
// ================================================================================================
// Large blocks (>= 256 MB: resident columns' temporaries, the radix aggregate's tuple slabs, 10^8-row result columns) are
// kept in a small exact-size free list per context instead of going back to the stream-ordered pool: re-executed plans
// ask for the same sizes again, and the pool otherwise fragments and re-maps tens of GB (steps of ~1 s instead of 55 ms
// were measured).  Everything is allocated, used and freed in the order of ctx->stream, so immediate re-use is safe.
static constexpr size_t kBigBlock = (size_t)256 << 20;
static constexpr size_t kBigCacheBytes = (size_t)96 << 30;
static inline size_t padded_size(size_t n) { return ((n + 255) / 256) * 256 + 256; }  // 16 B vector / bulk over-reads stay in bounds

DBuf::DBuf(Ctx* c, size_t n) : ctx(c), bytes(n) {
  const size_t alloc = padded_size(n);
  c->alloc_bytes_total += (int64_t)n;
  if (alloc >= kBigBlock) {
    for (size_t i = 0; i < c->big_free.size(); ++i)
      if (c->big_free[i].first == alloc) {
        ptr = c->big_free[i].second;
        c->big_free_bytes -= alloc;
        c->big_free.erase(c->big_free.begin() + (long)i);
        return;
      }
  }
  cudaError_t e = cudaMallocAsync(&ptr, alloc, c->pool, c->stream);
  if (e == cudaErrorMemoryAllocation && !c->big_free.empty()) {  // give the cached blocks back and retry
    cudaGetLastError();
    c->release_big_blocks();
    e = cudaMallocAsync(&ptr, alloc, c->pool, c->stream);
  }
  CUDA_CHECK(e);
}
DBuf::~DBuf() {
  if (!ptr || parent) return;  // a slab's sub-buffer owns nothing
  const size_t alloc = padded_size(bytes);
  if (alloc >= kBigBlock && !free_stream && ctx->big_free.size() < 48 && ctx->big_free_bytes + alloc <= kBigCacheBytes) {
    ctx->big_free.push_back({alloc, ptr});
    ctx->big_free_bytes += alloc;
    return;
  }
  cudaFreeAsync(ptr, free_stream ? free_stream : ctx->stream);
}
void Ctx::release_big_blocks() {
  for (auto& b : big_free) cudaFreeAsync(b.second, stream);
  big_free.clear();
  big_free_bytes = 0;
}

DBufP Ctx::alloc(size_t bytes) { return std::make_shared<DBuf>(this, bytes); }

void Ctx::reserve_pool() {
  if (!pool || pool_floor_bytes == 0) return;
  void* p = nullptr;
  if (cudaMallocAsync(&p, pool_floor_bytes, pool, stream) != cudaSuccess) {  // a small GPU: no reservation, not an error
    cudaGetLastError();
    return;
  }
  cudaFreeAsync(p, stream);
  cudaStreamSynchronize(stream);
}

Slab::Slab(Ctx* ctx, size_t total_bytes, bool zero) {
  buf = zero ? ctx->alloc_zero(total_bytes) : ctx->alloc(total_bytes);
}
DBufP Slab::take(size_t bytes) {
  const size_t n = need(bytes);
  if (off + n > buf->bytes + 256) throw_internal("slab overflow (internal error)");
  DBufP r = std::make_shared<DBuf>(buf, (char*)buf->ptr + off, bytes);
  off += n;
  return r;
}
DBufP Ctx::alloc_zero(size_t bytes) {
  DBufP b = alloc(bytes);
  CUDA_CHECK(cudaMemsetAsync(b->ptr, 0, ((bytes + 255) / 256) * 256 + 256, stream));
  return b;
}
void Ctx::sync() {
  CUDA_CHECK(cudaStreamSynchronize(stream));
  if (epi_stream) CUDA_CHECK(cudaStreamSynchronize(epi_stream));
}

void debug_sync_launch(Ctx* ctx, const char* name) {
  static int on = -1;
  if (on < 0) on = getenv("QGPU_SYNC_LAUNCH") ? 1 : 0;
  if (!on) return;
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) {
    fprintf(stderr, "[qgpu] kernel %s failed: %s\n", name, cudaGetErrorString(e));
    throw QError(QGPU_ERR_CUDA, std::string("CUDA error in ") + name + ": " + cudaGetErrorString(e));
  }
}

void Ctx::trace(const char* what) {
  if (trace_on < 0) trace_on = getenv("QGPU_TRACE") ? 1 : 0;
  if (!trace_on) return;
  cudaStreamSynchronize(stream);
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  const double now = ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
  if (what) fprintf(stderr, "[qgpu trace] %-40s %8.3f ms\n", what, now - trace_t0);
  trace_t0 = now;
}

void Ctx::h2d(void* dst, const void* src, size_t bytes) {
  if (bytes == 0) return;
  cudaPointerAttributes attr;
  bool pinned = false;
  if (cudaPointerGetAttributes(&attr, src) == cudaSuccess) pinned = (attr.type == cudaMemoryTypeHost);
  else cudaGetLastError();
  if (pinned || bytes <= 4096) {
    // pinned (or tiny) source: one async DMA on the compute stream
    CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream));
    if (!pinned) CUDA_CHECK(cudaStreamSynchronize(stream));
    return;
  }
  // pageable source: pipeline host memcpy -> pinned ring -> cudaMemcpyAsync on the side stream
  size_t off = 0;
  // the destination may have been allocated (stream-ordered) on `stream`: make copy_stream wait for it
  cudaEvent_t ready;
  CUDA_CHECK(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
  CUDA_CHECK(cudaEventRecord(ready, stream));
  CUDA_CHECK(cudaStreamWaitEvent(copy_stream, ready, 0));
  CUDA_CHECK(cudaEventDestroy(ready));
  while (off < bytes) {
    size_t n = std::min(stage_bytes, bytes - off);
    int s = stage_next;
    stage_next = (stage_next + 1) % kStageSlots;
    CUDA_CHECK(cudaEventSynchronize(stage_ev[s]));
    memcpy(stage[s], (const char*)src + off, n);
    CUDA_CHECK(cudaMemcpyAsync((char*)dst + off, stage[s], n, cudaMemcpyHostToDevice, copy_stream));
    CUDA_CHECK(cudaEventRecord(stage_ev[s], copy_stream));
    off += n;
  }
  // compute stream must observe the copies
  cudaEvent_t done;
  CUDA_CHECK(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
  CUDA_CHECK(cudaEventRecord(done, copy_stream));
  CUDA_CHECK(cudaStreamWaitEvent(stream, done, 0));
  CUDA_CHECK(cudaEventDestroy(done));
}

void Ctx::d2h_sync(void* dst, const void* src, size_t bytes) {
  if (bytes == 0) return;
  if (bytes <= pinned_scratch_bytes) {
    CUDA_CHECK(cudaMemcpyAsync(pinned_scratch, src, bytes, cudaMemcpyDeviceToHost, stream));
    CUDA_CHECK(cudaStreamSynchronize(stream));
    memcpy(dst, pinned_scratch, bytes);
  } else {
    CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, stream));
    CUDA_CHECK(cudaStreamSynchronize(stream));
  }
}

// ---- speculation on device-side counts (qgpu_internal.h: Speculation) ---------------------------------
SpecScope::SpecScope(Ctx* c, Speculation* s) : ctx(c), sp(s), saved(c->spec) {
  sp->replay = sp->have && !getenv("QGPU_NO_SPECULATION");
  sp->cursor = 0;
  if (sp->replay) {
    sp->check = ctx->alloc_zero(8 * (size_t)Speculation::kMax);
  } else {
    sp->learned.clear();
    sp->sites.clear();
    sp->have = false;
  }
  ctx->spec = sp;
}
bool SpecScope::verify() {
  if (!sp->replay) return true;
  bool ok = sp->cursor == sp->learned.size();
  if (ok && sp->cursor > 0) {
    std::vector<unsigned long long> got(sp->cursor);
    ctx->d2h_sync(got.data(), sp->check->ptr, 8 * sp->cursor);
    ok = memcmp(got.data(), sp->learned.data(), 8 * sp->cursor) == 0;
  }
  sp->check.reset();
  return ok;
}

bool SpecScope::defer_verify(const PendingP& p) {
  if (!sp->replay || !p) return verify();
  if (sp->cursor != sp->learned.size()) {
    sp->check.reset();
    return false;
  }
  if (sp->cursor > 0) {
    std::vector<unsigned long long> expect(sp->learned.begin(), sp->learned.begin() + (long)sp->cursor);
    DBufP keep = sp->check;
    p->also = make_pending(ctx, sp->check->ptr, (int)sp->cursor, [expect, keep](const unsigned long long* m, Pending&) {
      if (memcmp(m, expect.data(), expect.size() * 8) != 0)
        throw_internal("a speculated device-side count of an asynchronous execution did not match (were the plan's tables modified "
                       "while it ran?); re-run with qgpu_plan_execute_device");
    });
  }
  sp->check.reset();
  return true;
}

// ---- asynchronous result metadata (qgpu_internal.h: Pending) ------------------------------------------
PendingP make_pending(Ctx* ctx, const void* dev_meta, int words, std::function<void(const unsigned long long*, Pending&)> apply,
                      cudaStream_t stream) {
  if (!stream) stream = ctx->stream;
  if (words > META_WORDS) throw_internal("metadata block too large");
  MetaSlot* slot = nullptr;
  for (MetaSlot* s : ctx->meta_slots)
    if (!s->busy && cudaEventQuery(s->ev) == cudaSuccess) {  // an abandoned slot is reusable once its copy has landed
      slot = s;
      break;
    }
  cudaGetLastError();
  if (!slot) {
    slot = new MetaSlot();
    CUDA_CHECK(cudaHostAlloc((void**)&slot->host, META_WORDS * 8, cudaHostAllocDefault));
    CUDA_CHECK(cudaEventCreateWithFlags(&slot->ev, cudaEventDisableTiming));
    ctx->meta_slots.push_back(slot);
  }
  slot->busy = true;
  auto p = std::make_shared<Pending>();
  p->ctx = ctx;
  p->slot = slot;
  p->apply = std::move(apply);
  CUDA_CHECK(cudaMemcpyAsync(slot->host, dev_meta, (size_t)words * 8, cudaMemcpyDeviceToHost, stream));
  CUDA_CHECK(cudaEventRecord(slot->ev, stream));
  return p;
}
void Pending::resolve() {
  if (also) {
    PendingP a = also;
    also.reset();
    a->resolve();
  }
  if (!done) {
    done = true;
    cudaError_t e = cudaEventSynchronize(slot->ev);
    unsigned long long meta[META_WORDS];
    memcpy(meta, slot->host, sizeof(meta));
    slot->busy = false;
    slot = nullptr;
    if (e != cudaSuccess) {
      err_code = QGPU_ERR_CUDA;
      err_msg = std::string("CUDA error: ") + cudaGetErrorString(e);
    } else {
      try {
        if (apply) apply(meta, *this);
      } catch (QError& q) {
        err_code = q.code;
        err_msg = q.what();
      }
    }
    apply = nullptr;
  }
  if (err_code) throw QError(err_code, err_msg);
}
Pending::~Pending() {
  if (slot) slot->busy = false;  // the slot's event tells the next user when the abandoned copy has landed
}

// ---- per-kernel profiling (bench.py roofline leg) ---------------------------------------------------
int Ctx::prof_begin(const char* name) {
  if (prof_used == (int)prof_events.size()) {
    cudaEvent_t a, b;
    if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return -1;
    prof_events.push_back({a, b});
  }
  int slot = prof_used++;
  prof_names.push_back(name);
  cudaEventRecord(prof_events[slot].first, stream);
  return slot;
}
void Ctx::prof_end(int slot) { cudaEventRecord(prof_events[slot].second, stream); }
std::string Ctx::prof_report() {
  cudaStreamSynchronize(stream);
  struct Agg { int64_t n = 0; double ms = 0, mx = 0; };
  std::vector<std::pair<std::string, Agg>> aggs;
  for (int i = 0; i < prof_used; ++i) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, prof_events[i].first, prof_events[i].second) != cudaSuccess) continue;
    size_t j = 0;
    for (; j < aggs.size(); ++j)
      if (aggs[j].first == prof_names[i]) break;
    if (j == aggs.size()) aggs.push_back({prof_names[i], Agg()});
    aggs[j].second.n++;
    aggs[j].second.ms += ms;
    aggs[j].second.mx = std::max(aggs[j].second.mx, (double)ms);
  }
  std::string out;
  for (auto& a : aggs) {
    char buf[256];
    snprintf(buf, sizeof buf, "%s\t%lld\t%.6f\t%.6f\n", a.first.c_str(), (long long)a.second.n, a.second.ms, a.second.mx);
    out += buf;
  }
  prof_used = 0;
  prof_names.clear();
  return out;
}

// ================================================================================================
// scan
// ================================================================================================
#define SCAN_THREADS 512
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

__device__ __forceinline__ int64_t block_exclusive_scan(int64_t v, int64_t* total, int64_t* smem /*32*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int64_t x = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int64_t y = __shfl_up_sync(0xffffffffu, x, d);
    if (lane >= d) x += y;
  }
  if (lane == 31) smem[warp] = x;
  __syncthreads();
  if (warp == 0) {
    int nw = blockDim.x >> 5;
    int64_t s = lane < nw ? smem[lane] : 0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int64_t y = __shfl_up_sync(0xffffffffu, s, d);
      if (lane >= d) s += y;
    }
    smem[lane] = s;  // inclusive warp totals
  }
  __syncthreads();
  int64_t warp_off = warp == 0 ? 0 : smem[warp - 1];
  *total = smem[(blockDim.x >> 5) - 1];
  __syncthreads();
  return warp_off + x - v;
}

// per tile: local exclusive scan into out, tile total into sums[tile]
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(const int64_t* __restrict__ in, int64_t* __restrict__ out,
                                                              int64_t* __restrict__ sums, int64_t n) {
  __shared__ int64_t sm[32];
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int64_t v[SCAN_ITEMS];
  int64_t t = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    v[i] = (base + i < n) ? in[base + i] : 0;
    t += v[i];
  }
  int64_t total;
  int64_t off = block_exclusive_scan(t, &total, sm);
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    if (base + i < n) out[base + i] = off;
    off += v[i];
  }
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_add(int64_t* __restrict__ out, const int64_t* __restrict__ tile_off,
                                                            int64_t n) {
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  const int64_t add = tile_off[blockIdx.x];
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i)
    if (base + i < n) out[base + i] += add;
}

static void scan_rec(Ctx* ctx, const int64_t* in, int64_t* out, int64_t n, int64_t* total_dev) {
  int64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  if (tiles < 1) tiles = 1;
  DBufP sums = ctx->alloc((size_t)(tiles + 1) * 8);
  LAUNCH(ctx, k_scan_tiles, (int)tiles, SCAN_THREADS, 0, in, out, (int64_t*)sums->ptr, n);
  if (tiles == 1) {
    CUDA_CHECK(cudaMemcpyAsync(total_dev, sums->ptr, 8, cudaMemcpyDeviceToDevice, ctx->stream));
    return;
  }
  DBufP offs = ctx->alloc((size_t)tiles * 8);
  scan_rec(ctx, (const int64_t*)sums->ptr, (int64_t*)offs->ptr, tiles, total_dev);
  LAUNCH(ctx, k_scan_add, (int)tiles, SCAN_THREADS, 0, out, (const int64_t*)offs->ptr, n);
}

int64_t exclusive_scan_i64(Ctx* ctx, const int64_t* in, int64_t* out, int64_t n) {
  if (n <= 0) return 0;
  if (n > (int64_t)SCAN_TILE * 2147483647LL) throw_internal("scan too large");
  DBufP total = ctx->alloc(8);
  scan_rec(ctx, in, out, n, (int64_t*)total->ptr);
  return ctx->read_count((const int64_t*)total->ptr, "exclusive_scan_i64");
}

// ================================================================================================
// interpreter-driven evaluation
// ================================================================================================
struct OutCol {
  void* data;
  uint32_t* validity;
  uint8_t phys;
};

// Each warp handles 32 consecutive rows per iteration so that bit-packed outputs (booleans,
// validity) are produced with one ballot and written by one lane.
__global__ void __launch_bounds__(256) k_eval_store(const Program* __restrict__ Pp, OutCol out, int64_t n,
                                                    int* __restrict__ err, unsigned long long* __restrict__ null_count) {
  __shared__ Program P;
  for (int i = threadIdx.x; i < (int)(sizeof(Program) / 4); i += blockDim.x) ((uint32_t*)&P)[i] = ((const uint32_t*)Pp)[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t warp_id = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t n_words = (n + 31) >> 5;
  unsigned long long nulls = 0;
  for (int64_t w = warp_id; w < n_words; w += warps) {
    const int64_t row = (w << 5) + lane;
    Val v;
    v.lo = v.hi = 0;
    v.valid = 0;
    if (row < n) v = eval_row(P, row, err);
    const uint32_t vw = __ballot_sync(0xffffffffu, v.valid != 0);
    if (lane == 0) {
      out.validity[w] = vw;
      int live = (int)min((int64_t)32, n - (w << 5));
      nulls += live - __popc(vw);
    }
    if (out.phys == PH_BIT) {
      const uint32_t bw = __ballot_sync(0xffffffffu, v.valid && (v.lo & 1));
      if (lane == 0) ((uint32_t*)out.data)[w] = bw;
    } else if (row < n) {
      switch (out.phys) {
        case PH_I8: case PH_U8: ((uint8_t*)out.data)[row] = (uint8_t)v.lo; break;
        case PH_I16: case PH_U16: ((uint16_t*)out.data)[row] = (uint16_t)v.lo; break;
        case PH_I32: case PH_U32: ((uint32_t*)out.data)[row] = (uint32_t)v.lo; break;
        case PH_I64: case PH_U64: case PH_F64: ((uint64_t*)out.data)[row] = v.lo; break;
        case PH_F32: ((float*)out.data)[row] = (float)val_f64(v); break;
        case PH_I128: ((ulonglong2*)out.data)[row] = make_ulonglong2(v.lo, v.hi); break;
        default: break;
      }
    }
  }
  if (lane == 0 && nulls) atomicAdd(null_count, nulls);
}

// String-valued expressions (CASE over Utf8 branches -- case.rs:30-47 zips string arrays --, Utf8 literals): a row's value is
// a (pointer, length) pair into a source column or the literal blob.  Pass 1 stores the lengths (+ validity), a scan turns
// them into Arrow offsets, pass 2 evaluates again and copies the bytes.
__global__ void __launch_bounds__(256) k_eval_strlen(const Program* __restrict__ Pp, int64_t* __restrict__ lens,
                                                     uint32_t* __restrict__ validity, int64_t n, int* __restrict__ err,
                                                     unsigned long long* __restrict__ null_count) {
  __shared__ Program P;
  for (int i = threadIdx.x; i < (int)(sizeof(Program) / 4); i += blockDim.x) ((uint32_t*)&P)[i] = ((const uint32_t*)Pp)[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t warp_id = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t n_words = (n + 31) >> 5;
  unsigned long long nulls = 0;
  for (int64_t w = warp_id; w < n_words; w += warps) {
    const int64_t row = (w << 5) + lane;
    Val v;
    v.lo = v.hi = 0;
    v.valid = 0;
    if (row < n) {
      v = eval_row(P, row, err);
      lens[row] = v.valid ? (int64_t)v.hi : 0;
    }
    const uint32_t vw = __ballot_sync(0xffffffffu, v.valid != 0);
    if (lane == 0) {
      validity[w] = vw;
      nulls += (int)min((int64_t)32, n - (w << 5)) - __popc(vw);
    }
  }
  if (lane == 0 && nulls) atomicAdd(null_count, nulls);
}
__global__ void __launch_bounds__(256) k_eval_strcopy(const Program* __restrict__ Pp, const int64_t* __restrict__ offs64,
                                                      int32_t* __restrict__ offsets, char* __restrict__ data, int64_t n,
                                                      int64_t total, int* __restrict__ err) {
  __shared__ Program P;
  for (int i = threadIdx.x; i < (int)(sizeof(Program) / 4); i += blockDim.x) ((uint32_t*)&P)[i] = ((const uint32_t*)Pp)[i];
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < n; row += stride) {
    const Val v = eval_row(P, row, err);
    const int64_t o = offs64[row];
    offsets[row] = (int32_t)o;
    if (v.valid) {
      const char* src = (const char*)v.lo;
      for (int64_t i = 0; i < (int64_t)v.hi; ++i) data[o + i] = src[i];
    }
    if (row == n - 1) offsets[n] = (int32_t)total;
  }
}

// predicate -> keep bitmap words + per-word popcounts (int64 so the scan can be reused)
__global__ void __launch_bounds__(256) k_eval_mask(const Program* __restrict__ Pp, uint32_t* __restrict__ keep,
                                                   int64_t* __restrict__ counts, int64_t n, int* __restrict__ err) {
  __shared__ Program P;
  for (int i = threadIdx.x; i < (int)(sizeof(Program) / 4); i += blockDim.x) ((uint32_t*)&P)[i] = ((const uint32_t*)Pp)[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t warp_id = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t n_words = (n + 31) >> 5;
  for (int64_t w = warp_id; w < n_words; w += warps) {
    const int64_t row = (w << 5) + lane;
    bool k = false;
    if (row < n) {
      Val v = eval_row(P, row, err);
      k = v.valid && (v.lo & 1);
    }
    const uint32_t bw = __ballot_sync(0xffffffffu, k);
    if (lane == 0) {
      keep[w] = bw;
      counts[w] = __popc(bw);
    }
  }
}

__global__ void __launch_bounds__(256) k_select(const uint32_t* __restrict__ keep, const int64_t* __restrict__ offs,
                                                int64_t* __restrict__ sel, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < n; row += stride) {
    const uint32_t wv = keep[row >> 5];
    const int b = (int)(row & 31);
    if ((wv >> b) & 1u) sel[offs[row >> 5] + __popc(wv & ((1u << b) - 1u))] = row;
  }
}

static DBufP upload_program(Ctx* ctx, const Program& P) {
  DBufP b = ctx->alloc(sizeof(Program));
  // Program is POD; stage through pinned scratch when it fits, else a sync copy
  CUDA_CHECK(cudaMemcpyAsync(b->ptr, &P, sizeof(Program), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));  // P lives on the caller's stack
  return b;
}

static Phys result_phys(const DType& t) {
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
    case QGPU_T_NULL: return PH_NULL;
    default: return PH_STR;
  }
}

static void check_eval_err(Ctx* ctx, const int* err_dev) {
  int e = ctx->read_count(err_dev, "check_eval_err");
  if (e) throw_eval_error(e);
}

DColP eval_to_column(Ctx* ctx, Compiled& c, const View& v) {
  const int64_t n = v.num_rows;
  if (c.deferred_err && n > 0) throw_eval_error(c.deferred_err);
  auto col = std::make_shared<DCol>();
  col->type = c.result_type;
  col->length = n;
  col->phys = result_phys(c.result_type);
  if (col->phys == PH_NULL) {
    col->null_count = n;
    return col;
  }
  if (col->phys == PH_STR) {
    // string-valued computed expressions (CASE over Utf8 branches, Utf8 literals); bare string columns never reach here
    // (they are aliased).  Two evaluation passes around a scan of the lengths.
    const int64_t words = (n + 31) >> 5;
    col->offsets = ctx->alloc_zero((size_t)(n + 1) * 4);
    if (n == 0) {
      col->data = ctx->alloc(4);
      return col;
    }
    col->validity = ctx->alloc(std::max<size_t>((size_t)words * 4, 4));
    Program P = bind_program(ctx, c, v);
    DBufP dp = upload_program(ctx, P);
    DBufP flags = ctx->alloc_zero(16);
    DBufP lens = ctx->alloc((size_t)n * 8), offs64 = ctx->alloc((size_t)n * 8);
    LAUNCH(ctx, k_eval_strlen, grid_for(ctx, n, 256), 256, 0, (const Program*)dp->ptr, (int64_t*)lens->ptr,
           (uint32_t*)col->validity->ptr, n, (int*)flags->ptr, (unsigned long long*)((char*)flags->ptr + 8));
    const int64_t total = exclusive_scan_i64(ctx, (const int64_t*)lens->ptr, (int64_t*)offs64->ptr, n);
    struct { int err; int pad; unsigned long long nulls; } h;
    ctx->d2h_sync(&h, flags->ptr, 16);
    if (h.err) throw_eval_error(h.err);
    if (total > 2147483647LL) throw_arrow("Utf8 column exceeds 2 GiB of string data; LargeUtf8 is not supported");
    col->str_bytes = total;
    col->data = ctx->alloc(std::max<size_t>((size_t)total, 4));
    LAUNCH(ctx, k_eval_strcopy, grid_for(ctx, n, 256), 256, 0, (const Program*)dp->ptr, (const int64_t*)offs64->ptr,
           (int32_t*)col->offsets->ptr, (char*)col->data->ptr, n, total, (int*)flags->ptr);
    col->null_count = (int64_t)h.nulls;
    if (col->null_count == 0) col->validity.reset();
    return col;
  }
  const int64_t n_words = (n + 31) >> 5;
  size_t data_bytes = col->phys == PH_BIT ? (size_t)n_words * 4 : (size_t)n * phys_width(col->phys);
  col->data = ctx->alloc(std::max<size_t>(data_bytes, 4));
  col->validity = ctx->alloc(std::max<size_t>((size_t)n_words * 4, 4));
  if (n == 0) {
    col->validity.reset();
    return col;
  }
  Program P = bind_program(ctx, c, v);
  DBufP dp = upload_program(ctx, P);
  DBufP flags = ctx->alloc_zero(16);
  OutCol out{col->data->ptr, (uint32_t*)col->validity->ptr, (uint8_t)col->phys};
  LAUNCH(ctx, k_eval_store, grid_for(ctx, n, 256), 256, 0, (const Program*)dp->ptr, out, n, (int*)flags->ptr,
         (unsigned long long*)((char*)flags->ptr + 8));
  struct { int err; int pad; unsigned long long nulls; } h;
  ctx->d2h_sync(&h, flags->ptr, 16);
  if (h.err) throw_eval_error(h.err);
  col->null_count = (int64_t)h.nulls;
  if (col->null_count == 0) col->validity.reset();
  return col;
}

IdxP eval_filter(Ctx* ctx, Compiled& c, const View& v) {
  const int64_t n = v.num_rows;
  if (c.result_type.id != QGPU_T_BOOL) throw_internal("filter predicate must be Boolean, got " + c.result_type.str());
  if (c.deferred_err && n > 0) throw_eval_error(c.deferred_err);
  auto sel = std::make_shared<IdxVec>();
  if (n == 0) return sel;
  const int64_t n_words = (n + 31) >> 5;
  Program P = bind_program(ctx, c, v);
  DBufP dp = upload_program(ctx, P);
  DBufP keep = ctx->alloc((size_t)n_words * 4);
  DBufP counts = ctx->alloc((size_t)n_words * 8);
  DBufP offs = ctx->alloc((size_t)n_words * 8);
  DBufP err = ctx->alloc_zero(4);
  LAUNCH(ctx, k_eval_mask, grid_for(ctx, n, 256), 256, 0, (const Program*)dp->ptr, (uint32_t*)keep->ptr,
         (int64_t*)counts->ptr, n, (int*)err->ptr);
  int64_t total = exclusive_scan_i64(ctx, (const int64_t*)counts->ptr, (int64_t*)offs->ptr, n_words);
  check_eval_err(ctx, (const int*)err->ptr);
  sel->length = total;
  sel->buf = ctx->alloc(std::max<size_t>((size_t)total * 8, 8));
  if (total > 0)
    LAUNCH(ctx, k_select, grid_for(ctx, n, 256), 256, 0, (const uint32_t*)keep->ptr, (const int64_t*)offs->ptr,
           (int64_t*)sel->buf->ptr, n);
  return sel;
}

// ================================================================================================
// gathers
// ================================================================================================
__global__ void k_compose(const int64_t* __restrict__ inner, const int64_t* __restrict__ outer, int64_t* __restrict__ out,
                          int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    int64_t o = outer[i];
    out[i] = o < 0 ? -1 : inner[o];
  }
}

IdxP compose_idx(Ctx* ctx, const IdxP& inner, const IdxP& outer) {
  if (!inner) return outer;
  if (!outer) return inner;
  auto r = std::make_shared<IdxVec>();
  r->length = outer->length;
  r->may_have_null = inner->may_have_null || outer->may_have_null;
  r->buf = ctx->alloc(std::max<size_t>((size_t)outer->length * 8, 8));
  if (outer->length > 0)
    LAUNCH(ctx, k_compose, grid_for(ctx, outer->length, 256), 256, 0, inner->ptr(), outer->ptr(), (int64_t*)r->buf->ptr,
           outer->length);
  return r;
}

LazyCol apply_selection(Ctx* ctx, const LazyCol& col, const IdxP& sel, std::vector<std::pair<IdxP, IdxP>>* cache) {
  LazyCol out;
  out.base = col.base;
  if (!col.idx) {
    out.idx = sel;
    return out;
  }
  if (cache)
    for (auto& kv : *cache)
      if (kv.first == col.idx) {
        out.idx = kv.second;
        return out;
      }
  out.idx = compose_idx(ctx, col.idx, sel);
  if (cache) cache->push_back({col.idx, out.idx});
  return out;
}

View apply_selection_view(Ctx* ctx, const View& v, const IdxP& sel) {
  View out;
  out.schema = v.schema;
  out.num_rows = sel->length;
  out.num_batches = v.num_batches;
  std::vector<std::pair<IdxP, IdxP>> cache;
  for (const LazyCol& c : v.cols) out.cols.push_back(apply_selection(ctx, c, sel, &cache));
  return out;
}

template <typename T>
__global__ void k_take(const T* __restrict__ src, const int64_t* __restrict__ idx, T* __restrict__ dst, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    int64_t r = idx[i];
    if (r >= 0) dst[i] = src[r];
    else memset(&dst[i], 0, sizeof(T));
  }
}

// bit gathers: src_bits may be null (all ones); out word per warp via ballot
__global__ void k_take_bits(const uint32_t* __restrict__ src_bits, const int64_t* __restrict__ idx, uint32_t* __restrict__ dst,
                            int64_t n, bool null_idx_value, unsigned long long* __restrict__ zero_count) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t warp_id = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t n_words = (n + 31) >> 5;
  unsigned long long zeros = 0;
  for (int64_t w = warp_id; w < n_words; w += warps) {
    const int64_t i = (w << 5) + lane;
    bool b = false;
    if (i < n) {
      int64_t r = idx ? idx[i] : i;
      if (r < 0) b = null_idx_value;
      else b = src_bits ? ((src_bits[r >> 5] >> (r & 31)) & 1u) : true;
    }
    uint32_t bw = __ballot_sync(0xffffffffu, b);
    if (lane == 0) {
      dst[w] = bw;
      int live = (int)min((int64_t)32, n - (w << 5));
      zeros += live - __popc(bw);
    }
  }
  if (lane == 0 && zeros && zero_count) atomicAdd(zero_count, zeros);
}

__global__ void k_str_lens(const int32_t* __restrict__ offs, const int64_t* __restrict__ idx, int64_t* __restrict__ lens,
                           int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    int64_t r = idx[i];
    lens[i] = r < 0 ? 0 : (int64_t)(offs[r + 1] - offs[r]);
  }
}

__global__ void k_str_copy(const char* __restrict__ src, const int32_t* __restrict__ src_offs, const int64_t* __restrict__ idx,
                           const int64_t* __restrict__ new_offs, int32_t* __restrict__ out_offs, char* __restrict__ dst,
                           int64_t n, int64_t total) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += stride) {
    if (i == n) {
      out_offs[n] = (int32_t)total;
      continue;
    }
    int64_t o = new_offs[i];
    out_offs[i] = (int32_t)o;
    int64_t r = idx[i];
    if (r < 0) continue;
    int32_t s0 = src_offs[r], s1 = src_offs[r + 1];
    for (int32_t k = s0; k < s1; ++k) dst[o + (k - s0)] = src[k];
  }
}

__global__ void k_str_maxlen(const int32_t* __restrict__ offs, int64_t n, int* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int m = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) m = max(m, offs[i + 1] - offs[i]);
  for (int d = 16; d > 0; d >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
  if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(out, m);
}

// Gather of <= 1024 strings (group keys of a small aggregate result) in ONE single-block launch and without a host
// round trip: lengths -> block scan -> offsets + bytes.  The data buffer is sized by n x (longest value of the source).
__global__ void __launch_bounds__(1024) k_str_take_small(const char* __restrict__ src, const int32_t* __restrict__ src_offs,
                                                         const int64_t* __restrict__ idx, int32_t* __restrict__ out_offs,
                                                         char* __restrict__ dst, int n) {
  __shared__ int64_t sm[32];
  const int i = threadIdx.x;
  int64_t r = -1;
  int32_t s0 = 0, len = 0;
  if (i < n) {
    r = idx[i];
    if (r >= 0) {
      s0 = src_offs[r];
      len = src_offs[r + 1] - s0;
    }
  }
  int64_t total;
  const int64_t o = block_exclusive_scan((int64_t)len, &total, sm);
  if (i < n) {
    out_offs[i] = (int32_t)o;
    for (int32_t k = 0; k < len; ++k) dst[o + k] = src[s0 + k];
  }
  if (i == 0) out_offs[n] = (int32_t)total;
}

static int32_t max_str_len_of(Ctx* ctx, const DCol& base) {
  if (base.max_str_len < 0 && base.dict_state == 1) {  // dictionary-coded column: the dictionary holds every value
    size_t m = 0;
    for (const std::string& v : base.dict_values) m = std::max(m, v.size());
    base.max_str_len = (int32_t)m;
  }
  if (base.max_str_len < 0) {
    DBufP m = ctx->alloc_zero(8);
    if (base.length > 0)
      LAUNCH(ctx, k_str_maxlen, grid_for(ctx, base.length, 256), 256, 0, (const int32_t*)base.offsets->ptr, base.length, (int*)m->ptr);
    base.max_str_len = ctx->read_scalar((const int*)m->ptr);
  }
  return base.max_str_len;
}

DColP take_column(Ctx* ctx, const DCol& base, const int64_t* idx, int64_t n, bool idx_may_have_null) {
  auto col = std::make_shared<DCol>();
  col->type = base.type;
  col->phys = base.phys;
  col->length = n;
  if (base.phys == PH_NULL || base.length == 0) {
    col->phys = PH_NULL;
    col->null_count = n;
    return col;
  }
  const int g = grid_for(ctx, n, 256);
  const int64_t n_words = (n + 31) >> 5;
  ctx->gather_bytes_total += base.phys == PH_STR ? (n + 1) * 4 + (base.length > 0 ? (int64_t)((double)base.str_bytes / (double)base.length * (double)n) : 0)
                                                 : (base.phys == PH_BIT ? n_words * 4 : n * phys_width(base.phys));
  // validity: a NULL index or a NULL source slot => NULL (skipped when neither can occur)
  if (base.null_count != 0 || idx_may_have_null) {
    col->validity = ctx->alloc(std::max<size_t>((size_t)n_words * 4, 4));
    DBufP zc = ctx->alloc_zero(8);
    if (n > 0)
      LAUNCH(ctx, k_take_bits, g, 256, 0, base.validity ? (const uint32_t*)base.validity->ptr : nullptr, idx,
             (uint32_t*)col->validity->ptr, n, false, (unsigned long long*)zc->ptr);
    col->null_count = (int64_t)ctx->read_count((const unsigned long long*)zc->ptr, "take_column.nulls");
    if (col->null_count == 0) col->validity.reset();
  }
  if (n == 0) {
    col->data = ctx->alloc(16);
    if (base.phys == PH_STR) col->offsets = ctx->alloc_zero(16);
    return col;
  }
  switch (base.phys) {
    case PH_BIT:
      col->data = ctx->alloc((size_t)n_words * 4);
      LAUNCH(ctx, k_take_bits, g, 256, 0, (const uint32_t*)base.data->ptr, idx, (uint32_t*)col->data->ptr, n, false,
             (unsigned long long*)nullptr);
      break;
    case PH_I8: case PH_U8:
      col->data = ctx->alloc((size_t)n);
      LAUNCH(ctx, k_take<uint8_t>, g, 256, 0, (const uint8_t*)base.data->ptr, idx, (uint8_t*)col->data->ptr, n);
      break;
    case PH_I16: case PH_U16:
      col->data = ctx->alloc((size_t)n * 2);
      LAUNCH(ctx, k_take<uint16_t>, g, 256, 0, (const uint16_t*)base.data->ptr, idx, (uint16_t*)col->data->ptr, n);
      break;
    case PH_I32: case PH_U32: case PH_F32:
      col->data = ctx->alloc((size_t)n * 4);
      LAUNCH(ctx, k_take<uint32_t>, g, 256, 0, (const uint32_t*)base.data->ptr, idx, (uint32_t*)col->data->ptr, n);
      break;
    case PH_I64: case PH_U64: case PH_F64: case PH_D64:
      col->data = ctx->alloc((size_t)n * 8);
      LAUNCH(ctx, k_take<uint64_t>, g, 256, 0, (const uint64_t*)base.data->ptr, idx, (uint64_t*)col->data->ptr, n);
      break;
    case PH_I128:
      col->data = ctx->alloc((size_t)n * 16);
      LAUNCH(ctx, k_take<ulonglong2>, g, 256, 0, (const ulonglong2*)base.data->ptr, idx, (ulonglong2*)col->data->ptr, n);
      break;
    case PH_STR: {
      if (n <= 1024) {
        const int64_t cap = (int64_t)max_str_len_of(ctx, base) * n;
        if (cap <= (1 << 20)) {
          col->offsets = ctx->alloc((size_t)(n + 1) * 4);
          col->data = ctx->alloc(std::max<size_t>((size_t)cap, 4));
          col->str_bytes = cap;
          col->str_bytes_is_bound = true;
          LAUNCH(ctx, k_str_take_small, 1, 1024, 0, (const char*)base.data->ptr, (const int32_t*)base.offsets->ptr, idx,
                 (int32_t*)col->offsets->ptr, (char*)col->data->ptr, (int)n);
          break;
        }
      }
      DBufP lens = ctx->alloc((size_t)n * 8);
      DBufP offs64 = ctx->alloc((size_t)n * 8);
      LAUNCH(ctx, k_str_lens, g, 256, 0, (const int32_t*)base.offsets->ptr, idx, (int64_t*)lens->ptr, n);
      int64_t total = exclusive_scan_i64(ctx, (const int64_t*)lens->ptr, (int64_t*)offs64->ptr, n);
      if (total > 2147483647LL) throw_arrow("Utf8 column exceeds 2 GiB of string data; LargeUtf8 is not supported");
      col->offsets = ctx->alloc((size_t)(n + 1) * 4);
      col->data = ctx->alloc(std::max<size_t>((size_t)total, 4));
      col->str_bytes = total;
      LAUNCH(ctx, k_str_copy, grid_for(ctx, n + 1, 256), 256, 0, (const char*)base.data->ptr,
             (const int32_t*)base.offsets->ptr, idx, (const int64_t*)offs64->ptr, (int32_t*)col->offsets->ptr,
             (char*)col->data->ptr, n, total);
      break;
    }
    default: throw_internal("take: unsupported physical type");
  }
  return col;
}

__global__ void k_widen_d64(const int64_t* __restrict__ src, ulonglong2* __restrict__ dst, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    int64_t x = src[i];
    dst[i] = make_ulonglong2((uint64_t)x, x < 0 ? ~0ull : 0ull);
  }
}

DColP materialize(Ctx* ctx, const LazyCol& col, int64_t n) {
  if (!col.base) throw_internal("column was not uploaded to the GPU table");
  if (!col.idx) return col.base;
  return take_column(ctx, *col.base, col.idx->ptr(), n, col.idx->may_have_null);
}

DColP materialize_arrow(Ctx* ctx, const LazyCol& col, int64_t n) {
  DColP c = materialize(ctx, col, n);
  if (c->phys != PH_D64) return c;
  auto w = std::make_shared<DCol>(*c);
  w->phys = PH_I128;
  w->data = ctx->alloc(std::max<size_t>((size_t)n * 16, 16));
  if (n > 0)
    LAUNCH(ctx, k_widen_d64, grid_for(ctx, n, 256), 256, 0, (const int64_t*)c->data->ptr, (ulonglong2*)w->data->ptr, n);
  return w;
}

// ================================================================================================
// ingest helpers
// ================================================================================================
// Narrow 16 B decimals to int64 when hi == sign-extension(lo); also produces min/max of valid values.
__global__ void __launch_bounds__(256) k_narrow(const ulonglong2* __restrict__ src, const uint32_t* __restrict__ validity,
                                                int64_t* __restrict__ dst, int64_t n, int* __restrict__ fail,
                                                long long* __restrict__ mn, long long* __restrict__ mx) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  long long lmin = INT64_MAX, lmax = INT64_MIN;
  bool bad = false;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    ulonglong2 v = src[i];
    bool valid = validity ? ((validity[i >> 5] >> (i & 31)) & 1u) : true;
    int64_t lo = (int64_t)v.x;
    if (valid) {
      if (v.y != (lo < 0 ? ~0ull : 0ull)) bad = true;
      lmin = lo < lmin ? lo : lmin;
      lmax = lo > lmax ? lo : lmax;
    } else {
      lo = 0;
    }
    dst[i] = lo;
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) *fail = 1;
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    long long a = __shfl_xor_sync(0xffffffffu, lmin, d), b = __shfl_xor_sync(0xffffffffu, lmax, d);
    lmin = a < lmin ? a : lmin;
    lmax = b > lmax ? b : lmax;
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(mn, lmin);
    atomicMax(mx, lmax);
  }
}

DColP try_narrow_decimal(Ctx* ctx, const DCol& wide) {
  const int64_t n = wide.length;
  auto col = std::make_shared<DCol>(wide);
  col->phys = PH_D64;
  col->data = ctx->alloc(std::max<size_t>((size_t)n * 8, 8));
  struct { int fail; int pad; long long mn; long long mx; } h = {0, 0, INT64_MAX, INT64_MIN};
  DBufP flags = ctx->alloc(24);
  ctx->h2d(flags->ptr, &h, 24);
  if (n > 0)
    LAUNCH(ctx, k_narrow, grid_for(ctx, n, 256 * 4), 256, 0, (const ulonglong2*)wide.data->ptr,
           wide.validity ? (const uint32_t*)wide.validity->ptr : nullptr, (int64_t*)col->data->ptr, n, (int*)flags->ptr,
           (long long*)((char*)flags->ptr + 8), (long long*)((char*)flags->ptr + 16));
  ctx->d2h_sync(&h, flags->ptr, 24);
  if (h.fail) return nullptr;
  if (h.mn <= h.mx) {
    col->has_stats = true;
    col->vmin = h.mn;
    col->vmax = h.mx;
  }
  return col;
}

template <typename T>
__global__ void __launch_bounds__(256) k_minmax(const T* __restrict__ src, const uint32_t* __restrict__ validity, int64_t n,
                                                long long* __restrict__ mn, long long* __restrict__ mx) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  long long lmin = INT64_MAX, lmax = INT64_MIN;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    bool valid = validity ? ((validity[i >> 5] >> (i & 31)) & 1u) : true;
    if (valid) {
      long long x = (long long)src[i];
      lmin = x < lmin ? x : lmin;
      lmax = x > lmax ? x : lmax;
    }
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    long long a = __shfl_xor_sync(0xffffffffu, lmin, d), b = __shfl_xor_sync(0xffffffffu, lmax, d);
    lmin = a < lmin ? a : lmin;
    lmax = b > lmax ? b : lmax;
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(mn, lmin);
    atomicMax(mx, lmax);
  }
}

__global__ void k_init_minmax(long long* __restrict__ mn_mx) {
  if (threadIdx.x == 0) {
    mn_mx[0] = INT64_MAX;
    mn_mx[1] = INT64_MIN;
  }
}
// min / max of the non-NULL values into dev_mn_mx[0..1] (INT64_MAX / INT64_MIN when there are none), no host round trip;
// false: the column's physical type has no integer statistics
bool stats_to_device(Ctx* ctx, const DCol& col, long long* dev_mn_mx) {
  LAUNCH(ctx, k_init_minmax, 1, 32, 0, dev_mn_mx);
  if (col.length == 0 || col.null_count == col.length || !col.data) return col.phys != PH_STR;
  const uint32_t* val = col.validity ? (const uint32_t*)col.validity->ptr : nullptr;
  const int g = grid_for(ctx, col.length, 256 * 4);
  long long* mn = dev_mn_mx;
  long long* mx = dev_mn_mx + 1;
  switch (col.phys) {
    case PH_I8: LAUNCH(ctx, k_minmax<int8_t>, g, 256, 0, (const int8_t*)col.data->ptr, val, col.length, mn, mx); break;
    case PH_I16: LAUNCH(ctx, k_minmax<int16_t>, g, 256, 0, (const int16_t*)col.data->ptr, val, col.length, mn, mx); break;
    case PH_I32: LAUNCH(ctx, k_minmax<int32_t>, g, 256, 0, (const int32_t*)col.data->ptr, val, col.length, mn, mx); break;
    case PH_I64: case PH_D64:
      LAUNCH(ctx, k_minmax<int64_t>, g, 256, 0, (const int64_t*)col.data->ptr, val, col.length, mn, mx);
      break;
    case PH_U8: LAUNCH(ctx, k_minmax<uint8_t>, g, 256, 0, (const uint8_t*)col.data->ptr, val, col.length, mn, mx); break;
    case PH_U16: LAUNCH(ctx, k_minmax<uint16_t>, g, 256, 0, (const uint16_t*)col.data->ptr, val, col.length, mn, mx); break;
    case PH_U32: LAUNCH(ctx, k_minmax<uint32_t>, g, 256, 0, (const uint32_t*)col.data->ptr, val, col.length, mn, mx); break;
    default: return false;
  }
  return true;
}

void ensure_stats(Ctx* ctx, DCol& col) {
  if (col.has_stats || col.length == 0 || col.null_count == col.length) return;
  struct { long long mn; long long mx; } h = {INT64_MAX, INT64_MIN};
  DBufP flags = ctx->alloc(16);
  ctx->h2d(flags->ptr, &h, 16);
  const uint32_t* val = col.validity ? (const uint32_t*)col.validity->ptr : nullptr;
  const int g = grid_for(ctx, col.length, 256 * 4);
  long long* mn = (long long*)flags->ptr;
  long long* mx = mn + 1;
  switch (col.phys) {
    case PH_I8: LAUNCH(ctx, k_minmax<int8_t>, g, 256, 0, (const int8_t*)col.data->ptr, val, col.length, mn, mx); break;
    case PH_I16: LAUNCH(ctx, k_minmax<int16_t>, g, 256, 0, (const int16_t*)col.data->ptr, val, col.length, mn, mx); break;
    case PH_I32: LAUNCH(ctx, k_minmax<int32_t>, g, 256, 0, (const int32_t*)col.data->ptr, val, col.length, mn, mx); break;
    case PH_I64: case PH_D64:
      LAUNCH(ctx, k_minmax<int64_t>, g, 256, 0, (const int64_t*)col.data->ptr, val, col.length, mn, mx);
      break;
    case PH_U8: LAUNCH(ctx, k_minmax<uint8_t>, g, 256, 0, (const uint8_t*)col.data->ptr, val, col.length, mn, mx); break;
    case PH_U16: LAUNCH(ctx, k_minmax<uint16_t>, g, 256, 0, (const uint16_t*)col.data->ptr, val, col.length, mn, mx); break;
    case PH_U32: LAUNCH(ctx, k_minmax<uint32_t>, g, 256, 0, (const uint32_t*)col.data->ptr, val, col.length, mn, mx); break;
    default: return;  // u64 / i128 / floats / strings: no stats (callers fall back to the generic path)
  }
  ctx->d2h_sync(&h, flags->ptr, 16);
  if (h.mn <= h.mx) {
    col.has_stats = true;
    col.vmin = h.mn;
    col.vmax = h.mx;
  }
}

// dst bits [dst_bit, dst_bit+n) = src bits [src_bit, src_bit+n); launches are stream-ordered so the
// read-modify-write of boundary words never races with a neighbouring chunk's copy.
__global__ void k_copy_bits(uint32_t* __restrict__ dst, int64_t dst_bit, const uint32_t* __restrict__ src, int64_t src_bit,
                            int64_t n, int fill /* -1: copy, 0/1: fill */) {
  const int64_t w0 = dst_bit >> 5, w1 = (dst_bit + n - 1) >> 5;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t w = w0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w <= w1; w += stride) {
    int64_t b0 = max(dst_bit, w << 5), b1 = min(dst_bit + n, (w + 1) << 5);  // dst bit range in this word
    uint32_t mask = 0, bits = 0;
    for (int64_t b = b0; b < b1; ++b) {
      uint32_t m = 1u << (b & 31);
      mask |= m;
      bool v;
      if (fill >= 0) v = fill != 0;
      else {
        int64_t s = src_bit + (b - dst_bit);
        v = (src[s >> 5] >> (s & 31)) & 1u;
      }
      if (v) bits |= m;
    }
    dst[w] = (dst[w] & ~mask) | bits;
  }
}

void copy_bits(Ctx* ctx, uint32_t* dst, int64_t dst_bit, const uint32_t* src, int64_t src_bit, int64_t n) {
  if (n <= 0) return;
  LAUNCH(ctx, k_copy_bits, grid_for(ctx, (n >> 5) + 2, 256), 256, 0, dst, dst_bit, src, src_bit, n, -1);
}
void fill_bits(Ctx* ctx, uint32_t* dst, int64_t dst_bit, int64_t n, bool value) {
  if (n <= 0) return;
  LAUNCH(ctx, k_copy_bits, grid_for(ctx, (n >> 5) + 2, 256), 256, 0, dst, dst_bit, (const uint32_t*)nullptr, (int64_t)0, n,
         value ? 1 : 0);
}

__global__ void k_rebase(int32_t* __restrict__ dst, const int32_t* __restrict__ src, int64_t n, int64_t add) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = (int32_t)(src[i] + add);
}
void rebase_offsets(Ctx* ctx, int32_t* dst, const int32_t* src, int64_t n_plus_1, int64_t add) {
  if (n_plus_1 <= 0) return;
  LAUNCH(ctx, k_rebase, grid_for(ctx, n_plus_1, 256), 256, 0, dst, src, n_plus_1, add);
}

__global__ void k_popc(const uint32_t* __restrict__ bits, int64_t n, unsigned long long* __restrict__ out) {
  const int64_t n_words = (n + 31) >> 5;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  unsigned long long c = 0;
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += stride) {
    uint32_t v = bits[w];
    if (w == n_words - 1 && (n & 31)) v &= (1u << (n & 31)) - 1u;
    c += __popc(v);
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}
int64_t count_set_bits(Ctx* ctx, const uint32_t* bits, int64_t n) {
  if (n <= 0) return 0;
  DBufP c = ctx->alloc_zero(8);
  LAUNCH(ctx, k_popc, grid_for(ctx, (n >> 5) + 1, 256), 256, 0, bits, n, (unsigned long long*)c->ptr);
  return (int64_t)ctx->read_scalar((const unsigned long long*)c->ptr);
}

__global__ void k_iota(int64_t* __restrict__ out, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = i;
}
IdxP iota_idx(Ctx* ctx, int64_t n) {
  auto r = std::make_shared<IdxVec>();
  r->length = n;
  r->buf = ctx->alloc(std::max<size_t>((size_t)n * 8, 8));
  if (n > 0) LAUNCH(ctx, k_iota, grid_for(ctx, n, 256), 256, 0, (int64_t*)r->buf->ptr, n);
  return r;
}

}  // namespace qgpu
