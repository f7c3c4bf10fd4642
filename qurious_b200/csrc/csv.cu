// Delimited text -> HBM-resident columns, parsed on the GPU (on-disk -> HBM ingest).
//
// Replaces, for the path `COPY t FROM 'x.tbl' (DELIMITER '|')` / `read_csv` (planner/sql.rs:324-375 ->
// datasource/file/csv.rs:34-72 -> MemoryTable::insert), the host-side chain  arrow-rs csv Reader (1024-row batches of
// INFERRED types) -> CAST to the table's column types -> append.  Here the raw bytes of the file are copied to the device
// (pread straight into the pinned ring, one cudaMemcpyAsync per slot) and three kernels turn them into the table's
// DECLARED column types in the resident layout:
//   k_csv_count_nl / k_csv_row_starts   newline census per 1 KiB of text -> exclusive scan -> byte offset of every row
//   k_csv_parse                         one thread per row: split at the delimiter, parse every field by its column type
//                                       (Int*/UInt*, Float64, Decimal128(p <= 18, s) -> narrowed int64, Decimal128(p > 18) ->
//                                       i128, Date32 `YYYY-MM-DD`, Boolean, Utf8 -> (position, length) for the copy pass);
//                                       an EMPTY field is NULL for every type (arrow-rs NullRegex default: s.is_empty());
//                                       a trailing delimiter yields one more, empty, field (TPC-H .tbl: the `*_rev` columns)
//   k_csv_copy_strings                  Utf8 columns: lengths -> scan -> Arrow offsets + bytes
// For well-formed input the values equal the reference's infer-then-cast route: integers and dates are exact either way,
// and a decimal literal of <= 15 significant digits survives the reference's detour through Float64 unchanged
// (cast.rs: round(v * 10^s)).  Not supported on the device (InternalError, the host reader remains the way): quoting /
// escaping (CsvReadOptions::quote / escape), fractional digits beyond the column's scale, floats that need more than the
// exact double fast path (> 15 significant digits or |exponent| > 22).
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <cstring>
#include <mutex>
#include <thread>

#include "kernels.h"
#include "launch.h"
#include "plan.h"

namespace qgpu {

namespace {

constexpr int CSV_CHUNK = 1024;   // bytes per newline-census chunk
constexpr int CSV_MAXC = 32;      // columns per table

enum CsvKind : int { CK_SKIP = 0, CK_INT, CK_UINT, CK_F64, CK_F32, CK_D64, CK_D128, CK_DATE32, CK_BOOL, CK_STR, CK_NULL };
enum CsvErr : int { CE_NONE = 0, CE_FIELDS = 1, CE_PARSE = 2, CE_SCALE = 3, CE_FLOAT = 4, CE_RANGE = 5 };

struct CsvCol {
  void* data;            // values (fixed width) / (pos, len) pairs for strings
  uint32_t* validity;    // bitmap words, zero-initialised; set with atomicOr
  unsigned long long* nulls;
  int kind, width, scale, pad;
};
struct CsvParams {
  const unsigned char* text;
  int64_t len;
  const int64_t* row_start;  // n_rows + 1 entries (the last one = end of the last row + 1)
  int64_t n_rows;
  int n_cols;
  unsigned char delim;
  int* err;                  // [0] code, [1] column
  long long* err_row;
  CsvCol cols[CSV_MAXC];
};

__global__ void k_csv_count_nl(const unsigned char* __restrict__ text, int64_t len, int64_t* __restrict__ counts) {
  const int64_t chunk = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int64_t b0 = chunk * CSV_CHUNK;
  if (b0 >= len) return;
  int c = 0;
  for (int64_t i = b0 + lane; i < min(len, b0 + CSV_CHUNK); i += 32) c += text[i] == '\n';
#pragma unroll
  for (int d = 16; d; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
  if (lane == 0) counts[chunk] = c;
}
// row r+1 starts behind the r-th newline
__global__ void k_csv_row_starts(const unsigned char* __restrict__ text, int64_t len, const int64_t* __restrict__ offs,
                                 int64_t* __restrict__ row_start) {
  const int64_t chunk = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int64_t b0 = chunk * CSV_CHUNK;
  if (b0 >= len) return;
  int64_t at = offs[chunk];
  for (int64_t base = b0; base < min(len, b0 + CSV_CHUNK); base += 32) {
    const int64_t i = base + lane;
    const bool nl = i < len && i < b0 + CSV_CHUNK && text[i] == '\n';
    const uint32_t m = __ballot_sync(0xffffffffu, nl);
    if (nl) row_start[at + __popc(m & ((1u << lane) - 1u)) + 1] = i + 1;
    at += __popc(m);
  }
}

__device__ __forceinline__ void csv_fail(const CsvParams& p, int code, int col, int64_t row) {
  if (atomicCAS(&p.err[0], 0, code) == 0) p.err[1] = col;
  atomicMin(p.err_row, (long long)row);
}

// days since 1970-01-01 of a proleptic Gregorian date (Howard Hinnant's days_from_civil)
__device__ __forceinline__ int days_from_civil(int y, int m, int d) {
  y -= m <= 2;
  const int era = (y >= 0 ? y : y - 399) / 400;
  const int yoe = y - era * 400;
  const int doy = (153 * (m + (m > 2 ? -3 : 9)) + 2) / 5 + d - 1;
  const int doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
  return era * 146097 + doe - 719468;
}

__device__ __forceinline__ bool parse_int(const unsigned char* s, int n, bool allow_sign, i128* out) {
  int i = 0;
  bool neg = false;
  if (n > 0 && (s[0] == '-' || s[0] == '+')) {
    if (!allow_sign && s[0] == '-') return false;
    neg = s[0] == '-';
    i = 1;
  }
  if (i >= n) return false;
  i128 v = 0;
  for (; i < n; ++i) {
    const int d = s[i] - '0';
    if (d < 0 || d > 9) return false;
    v = v * 10 + d;
    if (v > ((i128)1 << 100)) return false;
  }
  *out = neg ? -v : v;
  return true;
}

// decimal literal -> unscaled integer at `scale`; CE_SCALE when the literal has more fractional digits than the column
__device__ __forceinline__ int parse_decimal(const unsigned char* s, int n, int scale, i128* out) {
  int i = 0;
  bool neg = false;
  if (n > 0 && (s[0] == '-' || s[0] == '+')) {
    neg = s[0] == '-';
    i = 1;
  }
  if (i >= n) return CE_PARSE;
  i128 v = 0;
  int frac = -1, digits = 0;
  for (; i < n; ++i) {
    if (s[i] == '.') {
      if (frac >= 0) return CE_PARSE;
      frac = 0;
      continue;
    }
    const int d = s[i] - '0';
    if (d < 0 || d > 9) return CE_PARSE;
    v = v * 10 + d;
    ++digits;
    if (frac >= 0) ++frac;
    if (digits > 38) return CE_RANGE;
  }
  if (digits == 0) return CE_PARSE;
  if (frac < 0) frac = 0;
  if (frac > scale) return CE_SCALE;
  for (int k = frac; k < scale; ++k) v *= 10;
  *out = neg ? -v : v;
  return CE_NONE;
}

// correctly rounded for the exact fast path: <= 15 significant digits and |power of ten| <= 22 (both operands exact doubles,
// one IEEE multiplication or division)
__device__ __forceinline__ int parse_f64(const unsigned char* s, int n, double* out) {
  const double P10[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
  int i = 0;
  bool neg = false;
  if (n > 0 && (s[0] == '-' || s[0] == '+')) {
    neg = s[0] == '-';
    i = 1;
  }
  if (i >= n) return CE_PARSE;
  unsigned long long m = 0;
  int sig = 0, frac = -1, e10 = 0;
  bool any = false;
  for (; i < n; ++i) {
    const unsigned char c = s[i];
    if (c == '.') {
      if (frac >= 0) return CE_PARSE;
      frac = 0;
      continue;
    }
    if (c == 'e' || c == 'E') break;
    const int d = c - '0';
    if (d < 0 || d > 9) return CE_PARSE;
    any = true;
    if (m != 0 || d != 0) {
      if (sig >= 15) return CE_FLOAT;
      m = m * 10 + d;
      ++sig;
    }
    if (frac >= 0) ++frac;
  }
  if (!any) return CE_PARSE;
  if (i < n) {  // exponent
    ++i;
    bool eneg = false;
    if (i < n && (s[i] == '-' || s[i] == '+')) {
      eneg = s[i] == '-';
      ++i;
    }
    if (i >= n) return CE_PARSE;
    int e = 0;
    for (; i < n; ++i) {
      const int d = s[i] - '0';
      if (d < 0 || d > 9) return CE_PARSE;
      e = e * 10 + d;
      if (e > 400) return CE_FLOAT;
    }
    e10 = eneg ? -e : e;
  }
  e10 -= frac > 0 ? frac : 0;
  double v = (double)m;
  if (m != 0) {
    if (e10 > 22 || e10 < -22) return CE_FLOAT;
    v = e10 >= 0 ? v * P10[e10] : v / P10[-e10];
  }
  *out = neg ? -v : v;
  return CE_NONE;
}

__global__ void __launch_bounds__(128) k_csv_parse(const __grid_constant__ CsvParams p) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < p.n_rows; row += stride) {
    int64_t b = p.row_start[row], e = p.row_start[row + 1] - 1;  // [b, e): without the newline
    if (e > p.len) e = p.len;
    if (e > b && p.text[e - 1] == '\r') --e;
    int col = 0;
    int64_t f0 = b;
    for (int64_t i = b; i <= e; ++i) {
      if (i < e && p.text[i] != p.delim) continue;
      // field [f0, i)
      if (col >= p.n_cols) {
        csv_fail(p, CE_FIELDS, col, row);
        break;
      }
      const CsvCol& c = p.cols[col];
      const unsigned char* s = p.text + f0;
      const int n = (int)(i - f0);
      if (c.kind != CK_SKIP) {
        bool valid = n > 0 && c.kind != CK_NULL;
        int err = CE_NONE;
        switch (c.kind) {
          case CK_INT:
          case CK_UINT: {
            i128 v = 0;
            if (valid) {
              if (!parse_int(s, n, c.kind == CK_INT, &v)) err = CE_PARSE;
              else {
                const i128 lo = c.kind == CK_UINT ? 0 : -((i128)1 << (8 * c.width - 1));
                const i128 hi = c.kind == CK_UINT ? (((i128)1 << (8 * c.width)) - 1) : (((i128)1 << (8 * c.width - 1)) - 1);
                if (v < lo || v > hi) err = CE_RANGE;
              }
            }
            if (c.width == 8) ((long long*)c.data)[row] = (long long)v;
            else if (c.width == 4) ((int*)c.data)[row] = (int)v;
            else if (c.width == 2) ((short*)c.data)[row] = (short)v;
            else ((signed char*)c.data)[row] = (signed char)v;
            break;
          }
          case CK_D64:
          case CK_D128: {
            i128 v = 0;
            if (valid) err = parse_decimal(s, n, c.scale, &v);
            if (c.kind == CK_D64) {
              if (v > (i128)INT64_MAX || v < (i128)INT64_MIN) err = CE_RANGE;
              ((long long*)c.data)[row] = (long long)v;
            } else {
              ((ulonglong2*)c.data)[row] = make_ulonglong2((unsigned long long)(u128)v, (unsigned long long)((u128)v >> 64));
            }
            break;
          }
          case CK_F64:
          case CK_F32: {
            double v = 0;
            if (valid) err = parse_f64(s, n, &v);
            if (c.kind == CK_F64) ((double*)c.data)[row] = v;
            else ((float*)c.data)[row] = (float)v;
            break;
          }
          case CK_DATE32: {
            int days = 0;
            if (valid) {
              // YYYY-MM-DD (arrow-rs parses Date32 through its ISO date parser; that is what TPC-H carries)
              bool ok = n == 10 && s[4] == '-' && s[7] == '-';
              int y = 0, m = 0, d = 0;
              if (ok) {
                for (int k = 0; k < 10 && ok; ++k) {
                  if (k == 4 || k == 7) continue;
                  const int dg = s[k] - '0';
                  ok = dg >= 0 && dg <= 9;
                  if (k < 4) y = y * 10 + dg;
                  else if (k < 7) m = m * 10 + dg;
                  else d = d * 10 + dg;
                }
              }
              const int mdays[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
              const bool leap = (y % 4 == 0 && y % 100 != 0) || y % 400 == 0;
              ok = ok && m >= 1 && m <= 12 && d >= 1 && d <= mdays[m - 1] + ((m == 2 && leap) ? 1 : 0);
              if (!ok) err = CE_PARSE;
              else days = days_from_civil(y, m, d);
            }
            ((int*)c.data)[row] = days;
            break;
          }
          case CK_BOOL: {
            bool v = false;
            if (valid) {
              const bool t = n == 4 && (s[0] | 32) == 't' && (s[1] | 32) == 'r' && (s[2] | 32) == 'u' && (s[3] | 32) == 'e';
              const bool f = n == 5 && (s[0] | 32) == 'f' && (s[1] | 32) == 'a' && (s[2] | 32) == 'l' && (s[3] | 32) == 's' && (s[4] | 32) == 'e';
              if (!t && !f) err = CE_PARSE;
              v = t;
            }
            ((unsigned char*)c.data)[row] = v ? 1 : 0;  // one byte per row, packed afterwards
            break;
          }
          case CK_STR: {
            ((longlong2*)c.data)[row] = make_longlong2(f0, valid ? n : 0);
            break;
          }
          default: break;
        }
        if (err != CE_NONE) {
          csv_fail(p, err, col, row);
          valid = false;
        }
        if (valid) atomicOr(&c.validity[row >> 5], 1u << (row & 31));
        else atomicAdd(c.nulls, 1ull);
      }
      ++col;
      f0 = i + 1;
    }
    if (col != p.n_cols) csv_fail(p, CE_FIELDS, col, row);
  }
}

__global__ void k_csv_str_len(const longlong2* __restrict__ pl, int64_t n, int64_t* __restrict__ lens) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) lens[i] = pl[i].y;
}
__global__ void k_csv_copy_strings(const unsigned char* __restrict__ text, const longlong2* __restrict__ pl, const int64_t* __restrict__ offs,
                                   int64_t n, int64_t total, int32_t* __restrict__ offsets, char* __restrict__ data) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const longlong2 v = pl[i];
    const int64_t o = offs[i];
    offsets[i] = (int32_t)o;
    for (int64_t k = 0; k < v.y; ++k) data[o + k] = (char)text[v.x + k];
    if (i == n - 1) offsets[n] = (int32_t)total;
  }
}
__global__ void k_csv_pack_bools(const unsigned char* __restrict__ bytes, int64_t n, uint32_t* __restrict__ bits) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t n_words = (n + 31) >> 5;
  for (int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < n_words; w += warps) {
    const int64_t i = (w << 5) + lane;
    const uint32_t m = __ballot_sync(0xffffffffu, i < n && bytes[i] != 0);
    if (lane == 0) bits[w] = m;
  }
}

}  // namespace

// text: DEVICE pointer to `len` bytes
TableChunk parse_csv_device(Ctx* ctx, const Schema& schema, const std::vector<char>& want, const unsigned char* text, int64_t len,
                            const qgpu_csv_options& opt) {
  if (opt.quote || opt.escape) throw_internal("CSV quoting / escaping is not supported by the device parser (CsvReadOptions::quote / escape)");
  const int nc = (int)schema.fields.size();
  if (nc > CSV_MAXC) throw_internal("the device CSV parser handles at most 32 columns");
  TableChunk ch;
  ch.cols.resize((size_t)nc);
  // ---- rows -----------------------------------------------------------------------------------------------------------
  const int64_t n_chunks = (len + CSV_CHUNK - 1) / CSV_CHUNK;
  int64_t n_rows = 0;
  DBufP row_start;
  if (len > 0) {
    DBufP counts = ctx->alloc_zero((size_t)n_chunks * 8), offs = ctx->alloc((size_t)n_chunks * 8);
    const int wpb = 8;  // warps per block
    LAUNCH(ctx, k_csv_count_nl, (int)((n_chunks + wpb - 1) / wpb), wpb * 32, 0, text, len, (int64_t*)counts->ptr);
    const int64_t n_nl = exclusive_scan_i64(ctx, (const int64_t*)counts->ptr, (int64_t*)offs->ptr, n_chunks);
    const unsigned char last = ctx->read_scalar(text + len - 1);
    n_rows = n_nl + (last == '\n' ? 0 : 1);  // a final line without its newline still counts
    row_start = ctx->alloc_zero((size_t)(n_rows + 2) * 8);
    LAUNCH(ctx, k_csv_row_starts, (int)((n_chunks + wpb - 1) / wpb), wpb * 32, 0, text, len, (const int64_t*)offs->ptr, (int64_t*)row_start->ptr);
    if (last != '\n') {
      const int64_t end = len + 1;
      CUDA_CHECK(cudaMemcpyAsync((int64_t*)row_start->ptr + n_rows, &end, 8, cudaMemcpyHostToDevice, ctx->stream));
      CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    }
  }
  const int64_t skip = (opt.has_header && n_rows > 0) ? 1 : 0;
  const int64_t rows = n_rows - skip;
  ch.rows = rows;
  const int64_t n_words = (rows + 31) >> 5;
  // ---- columns ----------------------------------------------------------------------------------------------------------
  CsvParams P;
  memset(&P, 0, sizeof(P));
  P.text = text;
  P.len = len;
  P.row_start = row_start ? (const int64_t*)row_start->ptr + skip : nullptr;
  P.n_rows = rows;
  P.n_cols = nc;
  P.delim = opt.delimiter ? opt.delimiter : (unsigned char)',';
  DBufP flags = ctx->alloc_zero(16 + 8 * (size_t)nc);
  P.err = (int*)flags->ptr;
  P.err_row = (long long*)((char*)flags->ptr + 8);
  const long long big = INT64_MAX;
  CUDA_CHECK(cudaMemcpyAsync(P.err_row, &big, 8, cudaMemcpyHostToDevice, ctx->stream));
  std::vector<DBufP> tmp((size_t)nc);  // strings: (pos, len); booleans: one byte per row
  for (int c = 0; c < nc; ++c) {
    CsvCol& cc = P.cols[c];
    cc.kind = CK_SKIP;
    if (!want[c]) continue;
    const DType& t = schema.fields[c].type;
    auto col = std::make_shared<DCol>();
    col->type = t;
    col->length = rows;
    ch.cols[c] = col;
    cc.nulls = (unsigned long long*)((char*)flags->ptr + 16) + c;
    cc.scale = t.scale;
    switch (t.id) {
      case QGPU_T_NULL: cc.kind = CK_NULL; col->phys = PH_NULL; break;
      case QGPU_T_BOOL: cc.kind = CK_BOOL; cc.width = 1; col->phys = PH_BIT; break;
      case QGPU_T_INT8: cc.kind = CK_INT; cc.width = 1; col->phys = PH_I8; break;
      case QGPU_T_INT16: cc.kind = CK_INT; cc.width = 2; col->phys = PH_I16; break;
      case QGPU_T_INT32: cc.kind = CK_INT; cc.width = 4; col->phys = PH_I32; break;
      case QGPU_T_INT64: cc.kind = CK_INT; cc.width = 8; col->phys = PH_I64; break;
      case QGPU_T_UINT8: cc.kind = CK_UINT; cc.width = 1; col->phys = PH_U8; break;
      case QGPU_T_UINT16: cc.kind = CK_UINT; cc.width = 2; col->phys = PH_U16; break;
      case QGPU_T_UINT32: cc.kind = CK_UINT; cc.width = 4; col->phys = PH_U32; break;
      case QGPU_T_UINT64: throw_internal("the device CSV parser does not read UInt64 columns");
      case QGPU_T_FLOAT32: cc.kind = CK_F32; cc.width = 4; col->phys = PH_F32; break;
      case QGPU_T_FLOAT64: cc.kind = CK_F64; cc.width = 8; col->phys = PH_F64; break;
      case QGPU_T_DATE32: cc.kind = CK_DATE32; cc.width = 4; col->phys = PH_I32; break;
      case QGPU_T_DECIMAL128:
        if (t.precision <= 18) { cc.kind = CK_D64; cc.width = 8; col->phys = PH_D64; }
        else { cc.kind = CK_D128; cc.width = 16; col->phys = PH_I128; }
        break;
      case QGPU_T_UTF8: cc.kind = CK_STR; cc.width = 16; col->phys = PH_STR; break;
      default: throw_internal("the device CSV parser does not read columns of type " + t.str());
    }
    if (cc.kind == CK_NULL) {
      col->null_count = rows;
      cc.validity = nullptr;
      continue;
    }
    col->validity = ctx->alloc_zero(std::max<size_t>((size_t)n_words * 4, 4));
    cc.validity = (uint32_t*)col->validity->ptr;
    if (cc.kind == CK_STR || cc.kind == CK_BOOL) {
      tmp[c] = ctx->alloc(std::max<size_t>((size_t)rows * cc.width, 16));
      cc.data = tmp[c]->ptr;
    } else {
      col->data = ctx->alloc(std::max<size_t>((size_t)rows * cc.width, 16));
      cc.data = col->data->ptr;
    }
  }
  // NULL-typed columns still consume their field; give them a dummy validity target
  if (rows > 0) LAUNCH(ctx, k_csv_parse, grid_for(ctx, rows, 128), 128, 0, P);
  std::vector<unsigned long long> h(2 + (size_t)nc);
  ctx->d2h_sync(h.data(), flags->ptr, h.size() * 8);
  const int err = (int)(h[0] & 0xffffffffu), err_col = (int)(h[0] >> 32);
  if (err != CE_NONE) {
    const std::string where = " (line " + std::to_string((long long)h[1] + 1 + skip) + ", column " + std::to_string(err_col + 1) + ")";
    switch (err) {
      case CE_FIELDS: throw_arrow("Csv error: incorrect number of fields" + where + ", expected " + std::to_string(nc));
      case CE_SCALE: throw_internal("the device CSV parser found more fractional digits than the column's scale" + where);
      case CE_FLOAT: throw_internal("the device CSV parser reads floats of <= 15 significant digits and |exponent| <= 22 only" + where);
      case CE_RANGE: throw_arrow("Parser error: value out of range for the column type" + where);
      default: throw_arrow("Parser error: Error while parsing value" + where);
    }
  }
  // ---- finish the columns ------------------------------------------------------------------------------------------------
  for (int c = 0; c < nc; ++c) {
    if (!want[c] || P.cols[c].kind == CK_NULL) continue;
    DCol& col = *ch.cols[c];
    col.null_count = (int64_t)h[2 + c];
    if (P.cols[c].kind == CK_STR) {
      col.offsets = ctx->alloc_zero((size_t)(rows + 1) * 4);
      int64_t total = 0;
      if (rows > 0) {
        DBufP lens = ctx->alloc((size_t)rows * 8), offs = ctx->alloc((size_t)rows * 8);
        LAUNCH(ctx, k_csv_str_len, grid_for(ctx, rows, 256), 256, 0, (const longlong2*)tmp[c]->ptr, rows, (int64_t*)lens->ptr);
        total = exclusive_scan_i64(ctx, (const int64_t*)lens->ptr, (int64_t*)offs->ptr, rows);
        if (total > 2147483647LL) throw_arrow("Utf8 column exceeds 2 GiB of string data; LargeUtf8 is not supported");
        col.data = ctx->alloc(std::max<size_t>((size_t)total, 4));
        LAUNCH(ctx, k_csv_copy_strings, grid_for(ctx, rows, 256), 256, 0, text, (const longlong2*)tmp[c]->ptr, (const int64_t*)offs->ptr, rows,
               total, (int32_t*)col.offsets->ptr, (char*)col.data->ptr);
      } else {
        col.data = ctx->alloc(4);
      }
      col.str_bytes = total;
    } else if (P.cols[c].kind == CK_BOOL) {
      col.data = ctx->alloc_zero(std::max<size_t>((size_t)n_words * 4, 4));
      if (rows > 0) LAUNCH(ctx, k_csv_pack_bools, grid_for(ctx, rows, 256), 256, 0, (const unsigned char*)tmp[c]->ptr, rows, (uint32_t*)col.data->ptr);
    }
    if (col.null_count == 0) col.validity.reset();
    if (col.null_count == rows && rows > 0 && false) col.phys = PH_NULL;
  }
  ctx->sync();  // `tmp` and the text may go
  return ch;
}

// the file's bytes -> one device buffer: worker threads pread 2 MiB ranges straight into their pinned ring slots (two per
// worker: the DMA of one overlaps the read of the other), one cudaMemcpyAsync per slot -- the same ring and threads as the
// staged Arrow ingest (ingest.cu).  A single pread thread tops out at ~5.4 GB/s from the page cache.
DBufP read_file_to_device(Ctx* ctx, const char* path, int64_t* len_out) {
  const int fd = open(path, O_RDONLY);
  if (fd < 0) throw_internal(std::string("file path: ") + path + ", err: " + strerror(errno));
  struct stat st;
  if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) {
    close(fd);
    throw_internal(std::string("file path: ") + path + " is not a regular file");
  }
  const int64_t len = (int64_t)st.st_size;
  *len_out = len;
  DBufP buf = ctx->alloc(std::max<size_t>((size_t)len, 16));
  const size_t slot_bytes = ctx->ingest_slot_bytes;
  const size_t ring_bytes = ctx->stage_bytes * Ctx::kStageSlots;
  const int64_t n_tasks = (len + (int64_t)slot_bytes - 1) / (int64_t)slot_bytes;
  int threads = ingest_worker_threads(ctx);
  threads = (int)std::min<size_t>((size_t)threads, ring_bytes / (2 * slot_bytes));
  threads = (int)std::max<int64_t>(1, std::min<int64_t>(threads, n_tasks));
  while ((int)ctx->ingest_ev.size() < threads * 2) {
    cudaEvent_t e;
    CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ctx->ingest_ev.push_back(e);
  }
  for (int i = 0; i < Ctx::kStageSlots; ++i) CUDA_CHECK(cudaEventSynchronize(ctx->stage_ev[i]));
  cudaEvent_t ready;
  CUDA_CHECK(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
  CUDA_CHECK(cudaEventRecord(ready, ctx->stream));
  CUDA_CHECK(cudaStreamWaitEvent(ctx->copy_stream, ready, 0));
  CUDA_CHECK(cudaEventDestroy(ready));
  const size_t per_half = ctx->stage_bytes / slot_bytes;
  auto slot_ptr = [&](int s) {
    return (size_t)s < per_half ? (char*)ctx->stage[0] + (size_t)s * slot_bytes : (char*)ctx->stage[1] + ((size_t)s - per_half) * slot_bytes;
  };
  std::atomic<int64_t> next(0);
  std::mutex err_mu;
  std::string err;
  auto worker = [&](int w) {
    if (cudaSetDevice(ctx->device) != cudaSuccess) return;
    int cur = 0;
    bool used[2] = {false, false};
    for (;;) {
      const int64_t t = next.fetch_add(1);
      if (t >= n_tasks) break;
      const int64_t off = t * (int64_t)slot_bytes;
      const size_t n = (size_t)std::min<int64_t>((int64_t)slot_bytes, len - off);
      const int s = w * 2 + cur;
      if (used[cur]) cudaEventSynchronize(ctx->ingest_ev[s]);
      char* slot = slot_ptr(s);
      size_t got = 0;
      while (got < n) {
        const ssize_t r = pread(fd, slot + got, n - got, off + (int64_t)got);
        if (r <= 0) break;
        got += (size_t)r;
      }
      bool bad = got < n;
      if (!bad) {
        bad = cudaMemcpyAsync((char*)buf->ptr + off, slot, n, cudaMemcpyHostToDevice, ctx->copy_stream) != cudaSuccess ||
              cudaEventRecord(ctx->ingest_ev[s], ctx->copy_stream) != cudaSuccess;
        used[cur] = true;
      }
      if (bad) {
        std::lock_guard<std::mutex> lk(err_mu);
        if (err.empty()) err = std::string("read error on ") + path;
        break;
      }
      cur ^= 1;
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
  close(fd);
  CUDA_CHECK(cudaStreamSynchronize(ctx->copy_stream));
  if (!err.empty()) throw_internal(err);
  return buf;
}

}  // namespace qgpu
