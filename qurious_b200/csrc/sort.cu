// Sort (+ top-N) and Limit -- SURVEY 8f "next" row #1: the operators right after the hot path in TPC-H Q1 / Q3.
//   Sort::execute     qurious/src/physical/plan/sort.rs:48-82   concat, evaluate the sort expressions, lexsort_to_indices
//                     with per-column SortOptions {descending, nulls_first} and an implicit final key = the row index
//                     (stable for equal keys), optional limit (top-N), take of every column; exactly one output batch
//   Limit::execute    qurious/src/physical/plan/limit.rs:27-58  skip + fetch over the concatenated input
//
// Device design: every sort column is normalised into order-preserving big-endian key bytes (sign bit flipped for signed
// integers / decimals, IEEE total order for floats like arrow-rs, strings padded to the column's longest value followed by
// their length, one leading byte per column placing NULLs first or last, value bytes inverted for DESC).  The row
// permutation is then sorted by a STABLE least-significant-digit radix sort over those bytes (4-bit digits: per-thread
// private digit counters in shared memory over contiguous runs keep every pass stable without warp ranking); byte planes
// on which all rows agree are skipped.  Starting from the identity permutation, stability IS the reference's row-index
// tie-break.  The sorted (and truncated) permutation becomes a selection vector: payload columns are gathered lazily.
#include <algorithm>
#include <cstring>

#include "launch.h"
#include "plan.h"

namespace qgpu {

constexpr int S_MAXC = 16;
constexpr int S_NT = 256;      // threads per block of the radix passes
constexpr int S_ITEMS = 8;     // contiguous elements per thread
constexpr int S_TILE = S_NT * S_ITEMS;

enum { SC_SIGNED = 0, SC_UNSIGNED = 1, SC_F64 = 2, SC_F32 = 3, SC_I128 = 4, SC_STR = 5, SC_BIT = 6 };

struct SortCol {
  int cls, width;  // value class, value bytes inside the key (strings: max_len + 4)
  int key_off;     // first byte of this column's segment (the NULL byte)
  int desc, nulls_first;
  int src_width;   // bytes per source value (fixed-width classes)
  const void* data;
  const int32_t* offsets;
  const uint32_t* validity;
};
struct SortArgs {
  int n_cols, key_bytes;
  SortCol c[S_MAXC];
};

// keys are stored as byte planes: plane b holds byte b of every row's key
__global__ void __launch_bounds__(256) k_sort_keys(const __grid_constant__ SortArgs a, int64_t n, unsigned char* __restrict__ planes) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < n; row += stride) {
    for (int ci = 0; ci < a.n_cols; ++ci) {
      const SortCol& c = a.c[ci];
      const bool valid = !c.validity || ((c.validity[row >> 5] >> (row & 31)) & 1u);
      unsigned char* out = planes + (size_t)c.key_off * n + row;
      out[0] = (unsigned char)(c.nulls_first ? (valid ? 1 : 0) : (valid ? 0 : 1));
      out += n;
      const unsigned char inv = c.desc ? 0xff : 0x00;
      if (!valid) {
        for (int b = 0; b < c.width; ++b) out[(size_t)b * n] = 0;
        continue;
      }
      if (c.cls == SC_STR) {
        const int32_t s = c.offsets[row], e = c.offsets[row + 1];
        const int len = e - s, max_len = c.width - 4;
        const unsigned char* src = (const unsigned char*)c.data + s;
        for (int b = 0; b < max_len; ++b) out[(size_t)b * n] = (unsigned char)((b < len ? src[b] : 0) ^ inv);
        for (int b = 0; b < 4; ++b) out[(size_t)(max_len + b) * n] = (unsigned char)(((unsigned)len >> (8 * (3 - b))) ^ inv);
        continue;
      }
      unsigned long long hi = 0, lo = 0;  // the value as up to 16 big-endian bytes: hi (only SC_I128) then lo
      switch (c.cls) {
        case SC_BIT:
          lo = (((const uint32_t*)c.data)[row >> 5] >> (row & 31)) & 1u;
          break;
        case SC_I128: {
          const ulonglong2 v = ((const ulonglong2*)c.data)[row];
          lo = v.x;
          hi = v.y ^ 0x8000000000000000ull;
          break;
        }
        case SC_F64: {
          const unsigned long long b = ((const unsigned long long*)c.data)[row];
          lo = (b >> 63) ? ~b : (b | 0x8000000000000000ull);
          break;
        }
        case SC_F32: {
          const unsigned int b = ((const unsigned int*)c.data)[row];
          lo = (b >> 31) ? (unsigned int)~b : (b | 0x80000000u);
          break;
        }
        default: {
          switch (c.src_width) {
            case 1: lo = ((const unsigned char*)c.data)[row]; break;
            case 2: lo = ((const unsigned short*)c.data)[row]; break;
            case 4: lo = ((const unsigned int*)c.data)[row]; break;
            default: lo = ((const unsigned long long*)c.data)[row]; break;
          }
          if (c.cls == SC_SIGNED) lo ^= 1ull << (8 * c.src_width - 1);
          break;
        }
      }
      int b = 0;
      if (c.cls == SC_I128)
        for (; b < 8; ++b) out[(size_t)b * n] = (unsigned char)((hi >> (8 * (7 - b))) ^ inv);
      const int lw = c.cls == SC_I128 ? 8 : c.width;
      for (int k = 0; k < lw; ++k, ++b) out[(size_t)b * n] = (unsigned char)((lo >> (8 * (lw - 1 - k))) ^ inv);
    }
  }
}

// which planes differ from row 0 anywhere?  (constant planes need no pass)
__global__ void __launch_bounds__(256) k_sort_plane_diff(const unsigned char* __restrict__ planes, int64_t n, int key_bytes,
                                                         unsigned int* __restrict__ plane_diff) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int b = 0; b < key_bytes; ++b) {
    const unsigned char* pl = planes + (size_t)b * n;
    const unsigned char first = pl[0];
    bool differs = false;
    for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < n; row += stride) differs |= pl[row] != first;
    if (differs) plane_diff[b] = 1;
  }
}

__global__ void __launch_bounds__(256) k_str_maxlen(const int32_t* __restrict__ offsets, const uint32_t* __restrict__ validity, int64_t n,
                                                    int* __restrict__ out) {
  int m = 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < n; row += stride)
    if (!validity || ((validity[row >> 5] >> (row & 31)) & 1u)) m = max(m, offsets[row + 1] - offsets[row]);
  if (m) atomicMax(out, m);
}

__device__ __forceinline__ int sort_digit(const unsigned char* __restrict__ plane, const long long* __restrict__ perm, int64_t i, int shift) {
  return (plane[perm[i]] >> shift) & 15;
}

// per-block digit totals, laid out digit-major: hist[d * n_blocks + block]
__global__ void __launch_bounds__(S_NT) k_rs_hist(const unsigned char* __restrict__ plane, const long long* __restrict__ perm, int64_t n,
                                                  int shift, long long* __restrict__ hist, int n_blocks) {
  __shared__ unsigned int h[16];
  if (threadIdx.x < 16) h[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * S_TILE + (int64_t)threadIdx.x * S_ITEMS;
  for (int k = 0; k < S_ITEMS; ++k)
    if (base + k < n) atomicAdd(&h[sort_digit(plane, perm, base + k, shift)], 1u);
  __syncthreads();
  if (threadIdx.x < 16) hist[(int64_t)threadIdx.x * n_blocks + blockIdx.x] = h[threadIdx.x];
}

// stable scatter: thread t owns the contiguous run [t * S_ITEMS, (t + 1) * S_ITEMS) of its block's tile
__global__ void __launch_bounds__(S_NT) k_rs_scatter(const unsigned char* __restrict__ plane, const long long* __restrict__ perm_in,
                                                     long long* __restrict__ perm_out, int64_t n, int shift,
                                                     const long long* __restrict__ offs, int n_blocks) {
  __shared__ unsigned short cnt[16][S_NT];
  __shared__ unsigned int wtot[S_NT / 32];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int64_t base = (int64_t)blockIdx.x * S_TILE + (int64_t)t * S_ITEMS;
  int dig[S_ITEMS];
  for (int d = 0; d < 16; ++d) cnt[d][t] = 0;
  for (int k = 0; k < S_ITEMS; ++k) {
    dig[k] = base + k < n ? sort_digit(plane, perm_in, base + k, shift) : -1;
    if (dig[k] >= 0) cnt[dig[k]][t]++;
  }
  __syncthreads();
  // for every digit: exclusive prefix of the per-thread counts over the threads (thread order = element order)
  unsigned int mine[16];
  for (int d = 0; d < 16; ++d) {
    const unsigned int c = cnt[d][t];
    unsigned int incl = c;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
      const unsigned int o = __shfl_up_sync(0xffffffffu, incl, s);
      if (lane >= s) incl += o;
    }
    if (lane == 31) wtot[warp] = incl;
    __syncthreads();
    unsigned int wbase = 0;
    for (int w = 0; w < warp; ++w) wbase += wtot[w];
    mine[d] = wbase + incl - c;
    __syncthreads();
  }
  for (int k = 0; k < S_ITEMS; ++k) {
    if (dig[k] < 0) continue;
    const int d = dig[k];
    perm_out[offs[(int64_t)d * n_blocks + blockIdx.x] + mine[d]++] = perm_in[base + k];
  }
}

static int sort_class(const DCol& c, int* src_width) {
  *src_width = phys_width(c.phys);
  switch (c.phys) {
    case PH_BIT: *src_width = 1; return SC_BIT;
    case PH_I8: case PH_I16: case PH_I32: case PH_I64: case PH_D64: return SC_SIGNED;
    case PH_U8: case PH_U16: case PH_U32: case PH_U64: return SC_UNSIGNED;
    case PH_F32: return SC_F32;
    case PH_F64: return SC_F64;
    case PH_I128: return SC_I128;
    case PH_STR: return SC_STR;
    default: throw_internal("Sort: unsupported sort key type " + c.type.str());
  }
  return 0;
}

View run_sort(PlanNode& node, const View& v) {
  Ctx* ctx = node.ctx;
  const int64_t n = v.num_rows;
  View out = v;
  out.num_batches = 1;  // sort.rs: always exactly one batch
  const int64_t keep = node.sort_limit >= 0 ? std::min<int64_t>(node.sort_limit, n) : n;
  if ((int)node.exprs.size() > S_MAXC) throw_internal("Sort: too many sort expressions");
  if (n <= 1 || node.exprs.empty()) {
    if (keep < n) {
      IdxP idx = iota_idx(ctx, keep);
      out = apply_selection_view(ctx, v, idx);
      out.num_batches = 1;
    }
    node.strategy = "sort(trivial)";
    return out;
  }
  // ---- sort columns ------------------------------------------------------------------------------------------------
  SortArgs a;
  memset(&a, 0, sizeof(a));
  std::vector<DColP> cols;
  int key_bytes = 0;
  for (size_t i = 0; i < node.exprs.size(); ++i) {
    auto c = compile_expr(*node.exprs[i], v.schema);
    DColP col = c->is_column_ref ? materialize(ctx, v.cols[c->column_ref], n) : eval_to_column(ctx, *c, v);
    if (!col) throw_internal("Sort: a sort key column is not resident");
    if (col->phys == PH_NULL) continue;  // an all-NULL key orders nothing
    cols.push_back(col);
    SortCol& sc = a.c[a.n_cols++];
    sc.cls = sort_class(*col, &sc.src_width);
    sc.desc = node.sort_desc[i];
    sc.nulls_first = node.sort_nulls_first[i];
    sc.data = col->data ? col->data->ptr : nullptr;
    sc.offsets = col->offsets ? (const int32_t*)col->offsets->ptr : nullptr;
    sc.validity = (col->validity && col->null_count != 0) ? (const uint32_t*)col->validity->ptr : nullptr;
    if (sc.cls == SC_STR) {
      DBufP m = ctx->alloc_zero(8);
      LAUNCH(ctx, k_str_maxlen, grid_for(ctx, n, 256), 256, 0, sc.offsets, sc.validity, n, (int*)m->ptr);
      sc.width = ctx->read_scalar((const int*)m->ptr) + 4;
    } else {
      sc.width = sc.cls == SC_I128 ? 16 : (sc.cls == SC_BIT ? 1 : sc.src_width);
    }
    sc.key_off = key_bytes;
    key_bytes += 1 + sc.width;
  }
  a.key_bytes = key_bytes;
  IdxP perm = iota_idx(ctx, n);
  int passes = 0;
  if (key_bytes > 0) {
    if ((double)key_bytes * (double)n > 64e9) throw_internal("Sort: the normalised keys would not fit (long string keys on a very large input)");
    DBufP planes = ctx->alloc((size_t)key_bytes * (size_t)n + 16);
    DBufP diff = ctx->alloc_zero((size_t)key_bytes * 4 + 16);
    LAUNCH(ctx, k_sort_keys, grid_for(ctx, n, 256), 256, 0, a, n, (unsigned char*)planes->ptr);
    LAUNCH(ctx, k_sort_plane_diff, grid_for(ctx, n, 256), 256, 0, (const unsigned char*)planes->ptr, n, key_bytes, (unsigned int*)diff->ptr);
    std::vector<unsigned int> hdiff((size_t)key_bytes);
    ctx->d2h_sync(hdiff.data(), diff->ptr, (size_t)key_bytes * 4);
    // ---- stable LSD radix sort of the permutation, least significant byte first, two 4-bit digits per byte -------------
    const int n_blocks = (int)((n + S_TILE - 1) / S_TILE);
    DBufP hist = ctx->alloc((size_t)16 * n_blocks * 8 + 16), offs = ctx->alloc((size_t)16 * n_blocks * 8 + 16);
    IdxP other = std::make_shared<IdxVec>();
    other->length = n;
    other->buf = ctx->alloc((size_t)n * 8 + 16);
    for (int b = key_bytes - 1; b >= 0; --b) {
      if (!hdiff[(size_t)b]) continue;
      const unsigned char* plane = (const unsigned char*)planes->ptr + (size_t)b * (size_t)n;
      for (int shift = 0; shift < 8; shift += 4) {
        LAUNCH(ctx, k_rs_hist, n_blocks, S_NT, 0, plane, (const long long*)perm->buf->ptr, n, shift, (long long*)hist->ptr, n_blocks);
        exclusive_scan_i64(ctx, (const int64_t*)hist->ptr, (int64_t*)offs->ptr, (int64_t)16 * n_blocks);
        LAUNCH(ctx, k_rs_scatter, n_blocks, S_NT, 0, plane, (const long long*)perm->buf->ptr, (long long*)other->buf->ptr, n, shift,
               (const long long*)offs->ptr, n_blocks);
        std::swap(perm, other);
        ++passes;
      }
    }
  }
  perm->length = keep;
  out = apply_selection_view(ctx, v, perm);
  out.num_batches = 1;
  node.strategy = "sort[" + std::to_string(a.n_cols) + " keys, " + std::to_string(key_bytes) + " key bytes, " + std::to_string(passes) +
                  " stable 4-bit radix passes" + (node.sort_limit >= 0 ? ", top " + std::to_string(keep) : std::string()) + "] <- " +
                  node.children[0]->strategy;
  return out;
}

View run_limit(PlanNode& node, const View& v) {
  Ctx* ctx = node.ctx;
  const int64_t n = v.num_rows;
  const int64_t skip = std::min<int64_t>(node.limit_skip, n);
  const int64_t fetch = node.limit_fetch >= 0 ? std::min<int64_t>(node.limit_fetch, n - skip) : n - skip;
  node.strategy = "limit <- " + node.children[0]->strategy;
  if (skip == 0 && fetch == n) return v;
  // rows [skip, skip + fetch) of the concatenated input (limit.rs walks the batches; the rows are the same)
  IdxP all = iota_idx(ctx, skip + fetch);
  IdxP idx = std::make_shared<IdxVec>();
  idx->length = fetch;
  idx->buf = ctx->alloc((size_t)std::max<int64_t>(fetch, 1) * 8 + 16);
  if (fetch > 0)
    CUDA_CHECK(cudaMemcpyAsync(idx->buf->ptr, (const char*)all->buf->ptr + (size_t)skip * 8, (size_t)fetch * 8, cudaMemcpyDeviceToDevice,
                               ctx->stream));
  View out = apply_selection_view(ctx, v, idx);
  out.num_batches = (n > skip && v.num_batches > 0) ? 1 : 0;  // limit.rs: a batch that is skipped entirely yields nothing
  return out;
}

}  // namespace qgpu
