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


// exclusive scan of up to a few 100 k int64 counters by ONE block (no host round trip between the radix passes)
__global__ void __launch_bounds__(1024) k_rs_scan(const long long* __restrict__ in, long long* __restrict__ out, int n) {
  __shared__ long long part[1024];
  const int t = threadIdx.x;
  const int per = (n + 1023) / 1024;
  const int i0 = min(n, t * per), i1 = min(n, i0 + per);
  long long s = 0;
  for (int i = i0; i < i1; ++i) s += in[i];
  part[t] = s;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {
    const long long x = t >= d ? part[t - d] : 0;
    __syncthreads();
    part[t] += x;
    __syncthreads();
  }
  long long run = part[t] - s;
  for (int i = i0; i < i1; ++i) {
    out[i] = run;
    run += in[i];
  }
}

// Small inputs (<= S_TILE rows, e.g. the 4 groups of Q1 or the candidates of a top-N): the whole stable LSD radix sort in
// ONE block.  ids[m] = the rows to order (ascending row ids: the tie-break), planes indexed by row id; constant planes are
// detected and skipped on the fly.  out[i] = i-th row id in sorted order, i < keep.
__global__ void __launch_bounds__(S_NT) k_sort_small(const unsigned char* __restrict__ planes, int64_t n_rows, int key_bytes,
                                                     const long long* __restrict__ ids, int m, long long* __restrict__ out, int keep) {
  __shared__ long long sid[S_TILE];
  __shared__ unsigned short perm[2][S_TILE];
  __shared__ unsigned short cnt[16][S_NT];
  __shared__ unsigned int wtot[S_NT / 32];
  __shared__ unsigned int dbase[16];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  for (int i = t; i < m; i += S_NT) {
    sid[i] = ids ? ids[i] : (long long)i;
    perm[0][i] = (unsigned short)i;
  }
  __syncthreads();
  int cur = 0;
  for (int b = key_bytes - 1; b >= 0; --b) {
    const unsigned char* plane = planes + (size_t)b * (size_t)n_rows;
    const unsigned char first = plane[sid[0]];
    int differs = 0;
    for (int i = t; i < m; i += S_NT) differs |= plane[sid[i]] != first;
    if (!__syncthreads_or(differs)) continue;
    for (int shift = 0; shift < 8; shift += 4) {
      int dig[S_ITEMS];
      for (int d = 0; d < 16; ++d) cnt[d][t] = 0;
      for (int k = 0; k < S_ITEMS; ++k) {
        const int i = t * S_ITEMS + k;
        dig[k] = i < m ? ((plane[sid[perm[cur][i]]] >> shift) & 15) : -1;
        if (dig[k] >= 0) cnt[dig[k]][t]++;
      }
      __syncthreads();
      unsigned int mine[16];
      for (int d = 0; d < 16; ++d) {
        const unsigned int c = cnt[d][t];
        unsigned int incl = c;
#pragma unroll
        for (int s2 = 1; s2 < 32; s2 <<= 1) {
          const unsigned int o = __shfl_up_sync(0xffffffffu, incl, s2);
          if (lane >= s2) incl += o;
        }
        if (lane == 31) wtot[warp] = incl;
        __syncthreads();
        unsigned int wbase = 0, total = 0;
        for (int w = 0; w < S_NT / 32; ++w) {
          if (w < warp) wbase += wtot[w];
          total += wtot[w];
        }
        mine[d] = wbase + incl - c;
        if (t == 0) dbase[d] = total;
        __syncthreads();
      }
      if (t == 0) {
        unsigned int run = 0;
        for (int d = 0; d < 16; ++d) {
          const unsigned int c = dbase[d];
          dbase[d] = run;
          run += c;
        }
      }
      __syncthreads();
      for (int k = 0; k < S_ITEMS; ++k) {
        if (dig[k] < 0) continue;
        const int d = dig[k];
        perm[cur ^ 1][dbase[d] + mine[d]++] = perm[cur][t * S_ITEMS + k];
      }
      __syncthreads();
      cur ^= 1;
    }
  }
  for (int i = t; i < keep && i < m; i += S_NT) out[i] = sid[perm[cur][i]];
}

// top-N prefilter: 16-bit prefix = the two most significant non-constant key bytes
__global__ void __launch_bounds__(256) k_topn_hist(const unsigned char* __restrict__ p0, const unsigned char* __restrict__ p1, int64_t n,
                                                   unsigned int* __restrict__ hist) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < n; row += stride) {
    // most rows share a few prefixes: one atomic per distinct prefix and warp instead of one per row
    const unsigned key = ((unsigned)p0[row] << 8) | (p1 ? (unsigned)p1[row] : 0u);
    const unsigned peers = __match_any_sync(__activemask(), key);
    if ((threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&hist[key], (unsigned)__popc(peers));
  }
}
// smallest prefix t with #rows(prefix <= t) >= limit; res[0] = t, res[1] = that row count
__global__ void __launch_bounds__(1024) k_topn_threshold(const unsigned int* __restrict__ hist, long long limit, long long* __restrict__ res) {
  __shared__ long long part[1024];
  const int t = threadIdx.x;
  long long s = 0;
  for (int i = 0; i < 64; ++i) s += hist[t * 64 + i];
  part[t] = s;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {
    const long long x = t >= d ? part[t - d] : 0;
    __syncthreads();
    part[t] += x;
    __syncthreads();
  }
  long long run = part[t] - s;  // rows with a smaller prefix than this thread's first bin
  if (run < limit && part[t] >= limit) {
    for (int i = 0; i < 64; ++i) {
      run += hist[t * 64 + i];
      if (run >= limit) {
        res[0] = t * 64 + i;
        res[1] = run;
        break;
      }
    }
  }
}
__global__ void __launch_bounds__(256) k_topn_flags(const unsigned char* __restrict__ p0, const unsigned char* __restrict__ p1, int64_t n,
                                                    const long long* __restrict__ res, long long* __restrict__ flags) {
  const unsigned thr = (unsigned)res[0];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < n; row += stride)
    flags[row] = ((((unsigned)p0[row] << 8) | (p1 ? (unsigned)p1[row] : 0u)) <= thr) ? 1 : 0;
}
__global__ void __launch_bounds__(256) k_topn_compact(const long long* __restrict__ flags, const long long* __restrict__ offs, int64_t n,
                                                      long long* __restrict__ ids) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < n; row += stride)
    if (flags[row]) ids[offs[row]] = row;
}
// the general passes work on a candidate list: perm[i] = ids[i] initially

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
      if (n <= S_TILE && col->str_bytes <= 256) {
        sc.width = (int)col->str_bytes + 4;  // no value is longer than all bytes together: no round trip for small inputs
      } else {
        DBufP m = ctx->alloc_zero(8);
        LAUNCH(ctx, k_str_maxlen, grid_for(ctx, n, 256), 256, 0, sc.offsets, sc.validity, n, (int*)m->ptr);
        sc.width = ctx->read_scalar((const int*)m->ptr) + 4;
      }
    } else {
      sc.width = sc.cls == SC_I128 ? 16 : (sc.cls == SC_BIT ? 1 : sc.src_width);
    }
    sc.key_off = key_bytes;
    key_bytes += 1 + sc.width;
  }
  a.key_bytes = key_bytes;
  IdxP perm;
  int passes = 0;
  std::string how = "stable 4-bit radix passes";
  if (key_bytes == 0 || keep == 0) {
    perm = iota_idx(ctx, keep);
  } else {
    if ((double)key_bytes * (double)n > 64e9) throw_internal("Sort: the normalised keys would not fit (long string keys on a very large input)");
    DBufP planes = ctx->alloc((size_t)key_bytes * (size_t)n + 16);
    LAUNCH(ctx, k_sort_keys, grid_for(ctx, n, 256), 256, 0, a, n, (unsigned char*)planes->ptr);
    const unsigned char* pl = (const unsigned char*)planes->ptr;
    auto new_idx = [&](int64_t len) {
      IdxP x = std::make_shared<IdxVec>();
      x->length = len;
      x->buf = ctx->alloc((size_t)std::max<int64_t>(len, 1) * 8 + 16);
      return x;
    };
    if (n <= S_TILE) {
      // ---- small input: keys + ONE single-block sort kernel, no host round trip ---------------------------------------------
      perm = new_idx(keep);
      LAUNCH(ctx, k_sort_small, 1, S_NT, 0, pl, n, key_bytes, (const long long*)nullptr, (int)n, (long long*)perm->buf->ptr, (int)keep);
      how = "single-block stable radix sort";
    } else {
      DBufP diff = ctx->alloc_zero((size_t)key_bytes * 4 + 16);
      LAUNCH(ctx, k_sort_plane_diff, grid_for(ctx, n, 256), 256, 0, pl, n, key_bytes, (unsigned int*)diff->ptr);
      std::vector<unsigned int> hdiff((size_t)key_bytes);
      ctx->d2h_sync(hdiff.data(), diff->ptr, (size_t)key_bytes * 4);
      // ---- top-N prefilter: rows whose 16-bit prefix (the two most significant non-constant bytes) lies above the
      //      limit-th smallest prefix cannot be among the first `limit` rows ------------------------------------------------
      IdxP ids;        // candidate rows in ascending row order (null: all rows)
      int64_t m = n;
      if (keep < n && keep <= S_TILE) {
        int p0 = -1, p1 = -1;
        for (int b = 0; b < key_bytes; ++b)
          if (hdiff[(size_t)b]) {
            if (p0 < 0) p0 = b;
            else if (p1 < 0) p1 = b;
          }
        if (p0 >= 0) {
          DBufP th = ctx->alloc_zero((size_t)65536 * 4 + 64);
          long long* res = (long long*)((char*)th->ptr + 65536 * 4);
          const unsigned char* q0 = pl + (size_t)p0 * (size_t)n;
          const unsigned char* q1 = p1 >= 0 ? pl + (size_t)p1 * (size_t)n : nullptr;
          LAUNCH(ctx, k_topn_hist, grid_for(ctx, n, 256), 256, 0, q0, q1, n, (unsigned int*)th->ptr);
          LAUNCH(ctx, k_topn_threshold, 1, 1024, 0, (const unsigned int*)th->ptr, (long long)keep, res);
          DBufP flags = ctx->alloc((size_t)n * 8 + 16), offs = ctx->alloc((size_t)n * 8 + 16);
          LAUNCH(ctx, k_topn_flags, grid_for(ctx, n, 256), 256, 0, q0, q1, n, (const long long*)res, (long long*)flags->ptr);
          m = exclusive_scan_i64(ctx, (const int64_t*)flags->ptr, (int64_t*)offs->ptr, n);
          ids = new_idx(m);
          LAUNCH(ctx, k_topn_compact, grid_for(ctx, n, 256), 256, 0, (const long long*)flags->ptr, (const long long*)offs->ptr, n,
                 (long long*)ids->buf->ptr);
          how = "top-N prefilter " + std::to_string(n) + " -> " + std::to_string(m) + " rows, " + how;
        }
      }
      if (m <= S_TILE) {
        perm = new_idx(keep);
        LAUNCH(ctx, k_sort_small, 1, S_NT, 0, pl, n, key_bytes, ids ? (const long long*)ids->buf->ptr : (const long long*)nullptr, (int)m,
               (long long*)perm->buf->ptr, (int)keep);
        how += " + single-block stable radix sort";
      } else {
        // ---- stable LSD radix sort of the (candidate) permutation, least significant byte first, two 4-bit digits per byte --
        perm = ids ? ids : iota_idx(ctx, n);
        const int n_blocks = (int)((m + S_TILE - 1) / S_TILE);
        DBufP hist = ctx->alloc((size_t)16 * n_blocks * 8 + 16), offs = ctx->alloc((size_t)16 * n_blocks * 8 + 16);
        IdxP other = new_idx(m);
        for (int b = key_bytes - 1; b >= 0; --b) {
          if (!hdiff[(size_t)b]) continue;
          const unsigned char* plane = pl + (size_t)b * (size_t)n;
          for (int shift = 0; shift < 8; shift += 4) {
            LAUNCH(ctx, k_rs_hist, n_blocks, S_NT, 0, plane, (const long long*)perm->buf->ptr, m, shift, (long long*)hist->ptr, n_blocks);
            if ((int64_t)16 * n_blocks <= ((int64_t)1 << 20))
              LAUNCH(ctx, k_rs_scan, 1, 1024, 0, (const long long*)hist->ptr, (long long*)offs->ptr, 16 * n_blocks);
            else
              exclusive_scan_i64(ctx, (const int64_t*)hist->ptr, (int64_t*)offs->ptr, (int64_t)16 * n_blocks);
            LAUNCH(ctx, k_rs_scatter, n_blocks, S_NT, 0, plane, (const long long*)perm->buf->ptr, (long long*)other->buf->ptr, m, shift,
                   (const long long*)offs->ptr, n_blocks);
            std::swap(perm, other);
            ++passes;
          }
        }
      }
    }
  }
  perm->length = keep;
  out = apply_selection_view(ctx, v, perm);
  out.num_batches = 1;
  node.strategy = "sort[" + std::to_string(a.n_cols) + " keys, " + std::to_string(key_bytes) + " key bytes, " +
                  (passes ? std::to_string(passes) + " " : std::string()) + how +
                  (node.sort_limit >= 0 ? ", top " + std::to_string(keep) : std::string()) + "] <- " + node.children[0]->strategy;
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
