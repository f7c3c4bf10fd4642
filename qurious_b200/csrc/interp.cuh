// Row-at-a-time expression interpreter shared by every generic kernel (projection, predicate mask,
// group keys / aggregate arguments, join keys / join filter) and -- compiled for the host -- by the
// planner's constant folder (Literal/Cast subtrees are constants: SURVEY 8a Q12).
//
// Semantics restated from the reference's call sites into arrow-rs:
//   BinaryExpr::evaluate   qurious/src/physical/expr/binary.rs:30-71
//   CastExpr::evaluate     qurious/src/physical/expr/cast.rs:32-38 (CastOptions{safe:false})
//   CaseExpr::evaluate     qurious/src/physical/expr/case.rs:30-47 (nested zip)
//   IsNull/IsNotNull/Negative  is_null.rs:25-33, is_not_null.rs:25-33, negative.rs:27-33
// The operand stack lives in per-thread local memory (L1-resident); the program is warp-uniform so
// there is no divergence on the opcode switch.  The fused kernels in fused.cu bypass this for the
// hot plan shapes; this is the always-correct generic path.
#pragma once
#include "qgpu_internal.h"

namespace qgpu {

enum VClass : uint8_t { VC_BOOL = 0, VC_INT = 1, VC_UINT = 2, VC_DEC = 3, VC_FLT = 4, VC_STR = 5, VC_NULLT = 6 };

enum OpCode : uint8_t {
  OP_COL = 1, OP_CONST, OP_CMP, OP_AND, OP_OR, OP_ARITH, OP_CAST, OP_CASE, OP_ISNULL, OP_ISNOTNULL, OP_NEG, OP_LIKE, OP_EXTRACT
};

enum EvalErr : int { EE_NONE = 0, EE_DIV_ZERO = 1, EE_CAST = 2, EE_OVERFLOW = 3, EE_DEC_PRECISION = 4, EE_PARSE = 5 };

struct Val {
  uint64_t lo;
  uint64_t hi;
  uint32_t valid;
  uint32_t pad;
};

struct ColRef {
  const void* data;
  const int32_t* offsets;
  const uint32_t* validity;
  const int64_t* idx;  // lazy-column row indices (-1 = NULL) or nullptr
  uint8_t phys;
  uint8_t pad[7];
};

struct Op {
  uint8_t code;
  uint8_t sub;     // Operator code 0..12 for CMP/ARITH
  uint8_t vclass;  // operand class (CMP/ARITH/NEG), source class (CAST)
  uint8_t wbits;   // ARITH/NEG on ints: result width in bits; FLT: 32 => round to f32
  int32_t arg;     // COL: slot; CONST: slot; CASE: n_when; ARITH(DEC add/sub/mod): const slot of {lm, rm}
  uint8_t to_id;   // CAST target
  uint8_t to_prec;
  int8_t to_scale;
  int8_t from_scale;  // CAST source decimal scale
  uint8_t from_id;    // CAST source type id
  uint8_t pad[3];
};

#define QGPU_MAX_OPS 96
#define QGPU_MAX_CONSTS 40
#define QGPU_MAX_COLS 24
#define QGPU_STACK 20

struct Program {
  int32_t n_ops;
  int32_t n_consts;
  int32_t n_cols;
  int32_t result_class;
  Op ops[QGPU_MAX_OPS];
  Val consts[QGPU_MAX_CONSTS];
  ColRef cols[QGPU_MAX_COLS];
};

#define QHD __host__ __device__ __forceinline__

QHD i128 val_i128(const Val& v) { return (i128)(((u128)v.hi << 64) | (u128)v.lo); }
QHD void set_i128(Val& v, i128 x) {
  v.lo = (uint64_t)(u128)x;
  v.hi = (uint64_t)((u128)x >> 64);
}
QHD double val_f64(const Val& v) {
  union { uint64_t u; double d; } c;
  c.u = v.lo;
  return c.d;
}
QHD void set_f64(Val& v, double d) {
  union { uint64_t u; double d; } c;
  c.d = d;
  v.lo = c.u;
  v.hi = 0;
}
QHD int64_t f64_total_key(double d) {
  union { int64_t i; double d; } c;
  c.d = d;
  return c.i ^ (int64_t)(((uint64_t)(c.i >> 63)) >> 1);
}
QHD i128 pow10_i128(int e) {
  i128 r = 1;
  for (int i = 0; i < e; ++i) r *= 10;
  return r;
}
QHD double pow10_f64(int e) {
  double r = 1.0;
  for (int i = 0; i < e; ++i) r *= 10.0;
  return r;
}
QHD bool dec_fits_precision(i128 v, int p) {
  i128 lim = pow10_i128(p);
  return v > -lim && v < lim;
}
QHD int64_t wrap_signed(int64_t x, int bits) {
  if (bits >= 64) return x;
  int sh = 64 - bits;
  return (int64_t)((uint64_t)x << sh) >> sh;
}
QHD uint64_t wrap_unsigned(uint64_t x, int bits) {
  if (bits >= 64) return x;
  return x & ((1ull << bits) - 1ull);
}
QHD double i128_to_f64(i128 v) {
  // matches `v as f64` for |v| < 2^53 exactly; above that within 1 ulp (tolerance 1e-12 applies)
  bool neg = v < 0;
  u128 a = neg ? (u128)(-(v + 1)) + 1 : (u128)v;
  double d = (double)(uint64_t)(a >> 64) * 18446744073709551616.0 + (double)(uint64_t)a;
  return neg ? -d : d;
}
// round-half-away-from-zero double -> i128; ok=false when not representable
QHD i128 f64_to_i128(double x, bool* ok) {
  *ok = true;
  if (!(x == x) || x > 1.7e38 || x < -1.7e38) {
    *ok = false;
    return 0;
  }
  bool neg = x < 0;
  double a = neg ? -x : x;
  if (a < 9.2e18) {
    int64_t r = (int64_t)a;
    return neg ? -(i128)r : (i128)r;
  }
  double hi = floor(a / 18446744073709551616.0);
  double lo = a - hi * 18446744073709551616.0;
  u128 r = ((u128)(uint64_t)hi << 64) + (u128)(uint64_t)lo;
  return neg ? -(i128)r : (i128)r;
}
QHD int64_t days_from_civil(int64_t y, unsigned m, unsigned d) {
  y -= m <= 2;
  const int64_t era = (y >= 0 ? y : y - 399) / 400;
  const unsigned yoe = (unsigned)(y - era * 400);
  const unsigned doy = (153 * (m + (m > 2 ? -3 : 9)) + 2) / 5 + d - 1;
  const unsigned doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
  return era * 146097 + (int64_t)doe - 719468;
}
// inverse of days_from_civil (proleptic Gregorian calendar; arrow date_part on Date32/Date64)
QHD void civil_from_days(int64_t z, int64_t* y, unsigned* m, unsigned* d) {
  z += 719468;
  const int64_t era = (z >= 0 ? z : z - 146096) / 146097;
  const unsigned doe = (unsigned)(z - era * 146097);
  const unsigned yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
  const int64_t yy = (int64_t)yoe + era * 400;
  const unsigned doy = doe - (365 * yoe + yoe / 4 - yoe / 100);
  const unsigned mp = (5 * doy + 2) / 153;
  *d = doy - (153 * mp + 2) / 5 + 1;
  *m = mp < 10 ? mp + 3 : mp - 9;
  *y = yy + (*m <= 2);
}
// SQL LIKE as arrow-rs `like` implements it: % = any sequence, _ = exactly one (UTF-8) character, \ escapes the next
// pattern character.  Iterative matcher with one backtracking point (the last %).
QHD int utf8_len_at(const unsigned char* s, int64_t i, int64_t n) {
  const unsigned char c = s[i];
  int l = c < 0x80 ? 1 : ((c >> 5) == 6 ? 2 : ((c >> 4) == 14 ? 3 : ((c >> 3) == 30 ? 4 : 1)));
  return i + l <= n ? l : (int)(n - i);
}
QHD bool like_match(const unsigned char* s, int64_t ls, const unsigned char* p, int64_t lp) {
  int64_t si = 0, pi = 0, star_p = -1, star_s = 0;
  while (si < ls) {
    bool ok = false;
    if (pi < lp) {
      const unsigned char c = p[pi];
      if (c == '%') {
        star_p = pi++;
        star_s = si;
        continue;
      }
      if (c == '_') {
        si += utf8_len_at(s, si, ls);
        ++pi;
        continue;
      }
      if (c == '\\' && pi + 1 < lp) {
        if (p[pi + 1] == s[si]) {
          ++si;
          pi += 2;
          ok = true;
        }
      } else if (c == s[si]) {
        ++si;
        ++pi;
        ok = true;
      }
    }
    if (ok) continue;
    if (star_p < 0) return false;
    pi = star_p + 1;
    star_s += utf8_len_at(s, star_s, ls);
    si = star_s;
  }
  while (pi < lp && p[pi] == '%') ++pi;
  return pi == lp;
}

QHD bool parse_date32(const char* s, int len, int64_t* out) {
  while (len > 0 && (s[0] == ' ')) { ++s; --len; }
  while (len > 0 && (s[len - 1] == ' ')) --len;
  if (len != 10 || s[4] != '-' || s[7] != '-') return false;
  int v[3] = {0, 0, 0};
  const int st[3] = {0, 5, 8}, ln[3] = {4, 2, 2};
  for (int k = 0; k < 3; ++k)
    for (int i = 0; i < ln[k]; ++i) {
      char c = s[st[k] + i];
      if (c < '0' || c > '9') return false;
      v[k] = v[k] * 10 + (c - '0');
    }
  if (v[1] < 1 || v[1] > 12 || v[2] < 1) return false;
  const int mdays[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
  int md = mdays[v[1] - 1];
  bool leap = (v[0] % 4 == 0 && v[0] % 100 != 0) || v[0] % 400 == 0;
  if (v[1] == 2 && leap) md = 29;
  if (v[2] > md) return false;
  *out = days_from_civil(v[0], (unsigned)v[1], (unsigned)v[2]);
  return true;
}
QHD int str_cmp(const Val& a, const Val& b) {
  const unsigned char* p = (const unsigned char*)a.lo;
  const unsigned char* q = (const unsigned char*)b.lo;
  int64_t la = (int64_t)a.hi, lb = (int64_t)b.hi;
  int64_t n = la < lb ? la : lb;
  for (int64_t i = 0; i < n; ++i) {
    if (p[i] != q[i]) return p[i] < q[i] ? -1 : 1;
  }
  return la < lb ? -1 : (la > lb ? 1 : 0);
}

QHD void raise_err(int* err, int code) {
#ifdef __CUDA_ARCH__
  atomicMax(err, code);
#else
  if (*err < code) *err = code;
#endif
}

QHD Val load_col(const ColRef& c, int64_t row) {
  Val v;
  v.lo = 0;
  v.hi = 0;
  v.valid = 1;
  v.pad = 0;
  int64_t r = row;
  if (c.idx) {
    r = c.idx[row];
    if (r < 0) {
      v.valid = 0;
      return v;
    }
  }
  if (c.phys == PH_NULL) {
    v.valid = 0;
    return v;
  }
  if (c.validity && !((c.validity[r >> 5] >> (r & 31)) & 1u)) {
    v.valid = 0;
    return v;
  }
  switch (c.phys) {
    case PH_BIT: v.lo = (((const uint32_t*)c.data)[r >> 5] >> (r & 31)) & 1u; break;
    case PH_I8: v.lo = (uint64_t)(int64_t)((const int8_t*)c.data)[r]; break;
    case PH_I16: v.lo = (uint64_t)(int64_t)((const int16_t*)c.data)[r]; break;
    case PH_I32: v.lo = (uint64_t)(int64_t)((const int32_t*)c.data)[r]; break;
    case PH_I64: v.lo = (uint64_t)((const int64_t*)c.data)[r]; break;
    case PH_U8: v.lo = ((const uint8_t*)c.data)[r]; break;
    case PH_U16: v.lo = ((const uint16_t*)c.data)[r]; break;
    case PH_U32: v.lo = ((const uint32_t*)c.data)[r]; break;
    case PH_U64: v.lo = ((const uint64_t*)c.data)[r]; break;
    case PH_F32: set_f64(v, (double)((const float*)c.data)[r]); break;
    case PH_F64: v.lo = ((const uint64_t*)c.data)[r]; break;
    case PH_I128: {
      const ulonglong2 t = ((const ulonglong2*)c.data)[r];
      v.lo = t.x;
      v.hi = t.y;
      break;
    }
    case PH_D64: {
      int64_t x = ((const int64_t*)c.data)[r];
      v.lo = (uint64_t)x;
      v.hi = x < 0 ? ~0ull : 0ull;
      break;
    }
    case PH_STR: {
      int32_t o0 = c.offsets[r], o1 = c.offsets[r + 1];
      v.lo = (uint64_t)((const char*)c.data + o0);
      v.hi = (uint64_t)(o1 - o0);
      break;
    }
    default: v.valid = 0; break;
  }
  return v;
}

QHD bool cmp_apply(int op, int c /* -1,0,1 */) {
  switch (op) {
    case 0: return c == 0;
    case 1: return c != 0;
    case 2: return c > 0;
    case 3: return c >= 0;
    case 4: return c < 0;
    default: return c <= 0;
  }
}

QHD int val_cmp(const Val& a, const Val& b, int vclass) {
  switch (vclass) {
    case VC_BOOL:
    case VC_UINT: return a.lo < b.lo ? -1 : (a.lo > b.lo ? 1 : 0);
    case VC_INT: {
      int64_t x = (int64_t)a.lo, y = (int64_t)b.lo;
      return x < y ? -1 : (x > y ? 1 : 0);
    }
    case VC_DEC: {
      i128 x = val_i128(a), y = val_i128(b);
      return x < y ? -1 : (x > y ? 1 : 0);
    }
    case VC_FLT: {
      int64_t x = f64_total_key(val_f64(a)), y = f64_total_key(val_f64(b));
      return x < y ? -1 : (x > y ? 1 : 0);
    }
    case VC_STR: return str_cmp(a, b);
    default: return 0;
  }
}

QHD Val do_arith(const Op& op, const Val& a, const Val& b, const Val* consts, int* err) {
  Val r;
  r.lo = 0;
  r.hi = 0;
  r.pad = 0;
  r.valid = a.valid & b.valid;
  if (!r.valid) return r;
  const int o = op.sub;  // 8 add, 9 sub, 10 mul, 11 div, 12 mod
  switch (op.vclass) {
    case VC_INT: {
      int64_t x = (int64_t)a.lo, y = (int64_t)b.lo, z = 0;
      if (o == 8) z = (int64_t)((uint64_t)x + (uint64_t)y);
      else if (o == 9) z = (int64_t)((uint64_t)x - (uint64_t)y);
      else if (o == 10) z = (int64_t)((uint64_t)x * (uint64_t)y);
      else {
        if (y == 0) {
          raise_err(err, EE_DIV_ZERO);
          return r;
        }
        if (y == -1) {
          // MIN / -1 overflows (arrow div_checked); MIN % -1 == 0
          int64_t mn = op.wbits >= 64 ? INT64_MIN : -((int64_t)1 << (op.wbits - 1));
          if (o == 11) {
            if (x == mn) {
              raise_err(err, EE_OVERFLOW);
              return r;
            }
            z = (int64_t)(0ull - (uint64_t)x);
          } else {
            z = 0;
          }
        } else {
          z = (o == 11) ? x / y : x % y;
        }
      }
      r.lo = (uint64_t)wrap_signed(z, op.wbits);
      break;
    }
    case VC_UINT: {
      uint64_t x = a.lo, y = b.lo, z = 0;
      if (o == 8) z = x + y;
      else if (o == 9) z = x - y;
      else if (o == 10) z = x * y;
      else {
        if (y == 0) {
          raise_err(err, EE_DIV_ZERO);
          return r;
        }
        z = (o == 11) ? x / y : x % y;
      }
      r.lo = wrap_unsigned(z, op.wbits);
      break;
    }
    case VC_DEC: {
      u128 x = (u128)val_i128(a), y = (u128)val_i128(b), z = 0;
      if (o == 10) {
        z = x * y;
      } else {
        u128 lm = (u128)val_i128(consts[op.arg]), rm = (u128)val_i128(consts[op.arg + 1]);
        x *= lm;
        y *= rm;
        if (o == 8) z = x + y;
        else if (o == 9) z = x - y;
        else {
          i128 sx = (i128)x, sy = (i128)y;
          if (sy == 0) {
            raise_err(err, EE_DIV_ZERO);
            return r;
          }
          u128 ax = sx < 0 ? (u128)0 - (u128)sx : (u128)sx;
          u128 ay = sy < 0 ? (u128)0 - (u128)sy : (u128)sy;
          u128 m = ax % ay;
          z = sx < 0 ? (u128)0 - m : m;
        }
      }
      set_i128(r, (i128)z);
      break;
    }
    case VC_FLT: {
      double x = val_f64(a), y = val_f64(b), z;
      if (o == 8) z = x + y;
      else if (o == 9) z = x - y;
      else if (o == 10) z = x * y;
      else if (o == 11) z = x / y;
      else z = fmod(x, y);
      if (op.wbits == 32) z = (double)(float)z;
      set_f64(r, z);
      break;
    }
    default: break;
  }
  return r;
}

QHD Val do_cast(const Op& op, const Val& a, int* err) {
  Val r;
  r.lo = 0;
  r.hi = 0;
  r.pad = 0;
  r.valid = a.valid;
  if (op.vclass == VC_NULLT) {
    r.valid = 0;
    return r;
  }
  if (!r.valid) return r;
  const int to = op.to_id;
  const bool to_sint = to >= QGPU_T_INT8 && to <= QGPU_T_INT64;
  const bool to_uint = to >= QGPU_T_UINT8 && to <= QGPU_T_UINT64;
  const int to_bits = to_sint ? (8 << (to - QGPU_T_INT8)) : (to_uint ? (8 << (to - QGPU_T_UINT8)) : 64);
  // integer range of the target
  i128 lo_lim = 0, hi_lim = 0;
  if (to_sint) {
    hi_lim = ((i128)1 << (to_bits - 1)) - 1;
    lo_lim = -hi_lim - 1;
  } else if (to_uint) {
    hi_lim = ((i128)1 << to_bits) - 1;
    lo_lim = 0;
  }
  switch (op.vclass) {
    case VC_BOOL:
    case VC_INT:
    case VC_UINT: {
      i128 x = (op.vclass == VC_INT) ? (i128)(int64_t)a.lo : (i128)(uint64_t)a.lo;
      if (to_sint || to_uint) {
        if (x < lo_lim || x > hi_lim) {
          raise_err(err, EE_CAST);
          return r;
        }
        r.lo = (uint64_t)(int64_t)x;
      } else if (to == QGPU_T_FLOAT64 || to == QGPU_T_FLOAT32) {
        double d = (op.vclass == VC_INT) ? (double)(int64_t)a.lo : (double)(uint64_t)a.lo;
        if (to == QGPU_T_FLOAT32) d = (double)(float)d;
        set_f64(r, d);
      } else if (to == QGPU_T_DECIMAL128) {
        // x * 10^s cannot overflow i128 for |x| < 2^64 and s <= 38 only if checked
        i128 m = pow10_i128(op.to_scale);
        i128 z = x * m;
        if (x != 0 && (z / m != x)) {
          raise_err(err, EE_CAST);
          return r;
        }
        if (!dec_fits_precision(z, op.to_prec)) {
          raise_err(err, EE_DEC_PRECISION);
          return r;
        }
        set_i128(r, z);
      } else if (to == QGPU_T_DATE32 || to == QGPU_T_DATE64) {
        r.lo = (uint64_t)(int64_t)x;
      } else if (to == QGPU_T_BOOL) {
        r.lo = x != 0;
      }
      break;
    }
    case VC_FLT: {
      double d = val_f64(a);
      if (to == QGPU_T_FLOAT64) {
        set_f64(r, d);
      } else if (to == QGPU_T_FLOAT32) {
        set_f64(r, (double)(float)d);
      } else if (to_sint || to_uint) {
        double t = trunc(d);
        if (!(d == d) || t < (double)lo_lim || t > (double)hi_lim) {
          raise_err(err, EE_CAST);
          return r;
        }
        bool ok;
        i128 z = f64_to_i128(t, &ok);
        if (!ok || z < lo_lim || z > hi_lim) {
          raise_err(err, EE_CAST);
          return r;
        }
        r.lo = (uint64_t)(int64_t)z;
      } else if (to == QGPU_T_DECIMAL128) {
        double x = d * pow10_f64(op.to_scale);
        x = (x >= 0) ? floor(x + 0.5) : -floor(-x + 0.5);  // f64::round
        bool ok;
        i128 z = f64_to_i128(x, &ok);
        if (!ok) {
          raise_err(err, EE_CAST);
          return r;
        }
        if (!dec_fits_precision(z, op.to_prec)) {
          raise_err(err, EE_DEC_PRECISION);
          return r;
        }
        set_i128(r, z);
      }
      break;
    }
    case VC_DEC: {
      i128 x = val_i128(a);
      if (to == QGPU_T_DECIMAL128) {
        int ds = (int)op.to_scale - (int)op.from_scale;
        i128 z;
        if (ds >= 0) {
          i128 m = pow10_i128(ds);
          z = x * m;
          if (x != 0 && z / m != x) {
            raise_err(err, EE_CAST);
            return r;
          }
        } else {
          i128 dv = pow10_i128(-ds);
          bool neg = x < 0;
          u128 ax = neg ? (u128)0 - (u128)x : (u128)x;
          u128 q = ax / (u128)dv, rem = ax % (u128)dv;
          if (2 * rem >= (u128)dv) q += 1;  // round half away from zero
          z = neg ? -(i128)q : (i128)q;
        }
        if (!dec_fits_precision(z, op.to_prec)) {
          raise_err(err, EE_DEC_PRECISION);
          return r;
        }
        set_i128(r, z);
      } else if (to == QGPU_T_FLOAT64 || to == QGPU_T_FLOAT32) {
        double d = i128_to_f64(x) / pow10_f64(op.from_scale);
        if (to == QGPU_T_FLOAT32) d = (double)(float)d;
        set_f64(r, d);
      } else if (to_sint || to_uint) {
        i128 dv = pow10_i128(op.from_scale);
        i128 q = x / dv;  // truncation toward zero
        if (q < lo_lim || q > hi_lim) {
          raise_err(err, EE_CAST);
          return r;
        }
        r.lo = (uint64_t)(int64_t)q;
      }
      break;
    }
    case VC_STR: {
      const char* s = (const char*)a.lo;
      int len = (int)a.hi;
      if (to == QGPU_T_DATE32) {
        int64_t d;
        if (!parse_date32(s, len, &d)) {
          raise_err(err, EE_PARSE);
          return r;
        }
        r.lo = (uint64_t)d;
      } else if (to_sint || to_uint) {
        bool neg = false;
        int i = 0;
        if (len > 0 && (s[0] == '-' || s[0] == '+')) {
          neg = s[0] == '-';
          i = 1;
        }
        if (i >= len) {
          raise_err(err, EE_PARSE);
          return r;
        }
        i128 z = 0;
        for (; i < len; ++i) {
          if (s[i] < '0' || s[i] > '9' || z > ((i128)1 << 70)) {
            raise_err(err, EE_PARSE);
            return r;
          }
          z = z * 10 + (s[i] - '0');
        }
        if (neg) z = -z;
        if (z < lo_lim || z > hi_lim) {
          raise_err(err, EE_PARSE);
          return r;
        }
        r.lo = (uint64_t)(int64_t)z;
      }
      break;
    }
    default: break;
  }
  return r;
}

// Evaluate the program for one row.  `err` receives the max EvalErr raised.
QHD Val eval_row(const Program& P, int64_t row, int* err) {
  Val st[QGPU_STACK];
  int sp = 0;
  for (int pc = 0; pc < P.n_ops; ++pc) {
    const Op op = P.ops[pc];
    switch (op.code) {
      case OP_COL: st[sp++] = load_col(P.cols[op.arg], row); break;
      case OP_CONST: st[sp++] = P.consts[op.arg]; break;
      case OP_CMP: {
        const Val b = st[--sp];
        const Val a = st[sp - 1];
        Val r;
        r.hi = 0;
        r.pad = 0;
        r.valid = a.valid & b.valid;
        r.lo = r.valid ? (uint64_t)cmp_apply(op.sub, val_cmp(a, b, op.vclass)) : 0;
        st[sp - 1] = r;
        break;
      }
      case OP_AND: {  // and_kleene: false AND NULL = false
        const Val b = st[--sp];
        const Val a = st[sp - 1];
        bool at = a.valid && a.lo, af = a.valid && !a.lo, bt = b.valid && b.lo, bf = b.valid && !b.lo;
        Val r;
        r.hi = 0;
        r.pad = 0;
        r.lo = at && bt;
        r.valid = (at && bt) || af || bf;
        st[sp - 1] = r;
        break;
      }
      case OP_OR: {  // or_kleene: true OR NULL = true
        const Val b = st[--sp];
        const Val a = st[sp - 1];
        bool at = a.valid && a.lo, af = a.valid && !a.lo, bt = b.valid && b.lo, bf = b.valid && !b.lo;
        Val r;
        r.hi = 0;
        r.pad = 0;
        r.lo = at || bt;
        r.valid = at || bt || (af && bf);
        st[sp - 1] = r;
        break;
      }
      case OP_ARITH: {
        const Val b = st[--sp];
        const Val a = st[sp - 1];
        st[sp - 1] = do_arith(op, a, b, P.consts, err);
        break;
      }
      case OP_CAST: st[sp - 1] = do_cast(op, st[sp - 1], err); break;
      case OP_CASE: {
        // stack: when1 then1 ... whenN thenN else ; result = first WHEN that is true AND valid
        const int n = op.arg;
        const int base = sp - (2 * n + 1);
        Val r = st[sp - 1];
        for (int i = n - 1; i >= 0; --i) {
          const Val w = st[base + 2 * i];
          if (w.valid && w.lo) r = st[base + 2 * i + 1];
        }
        sp = base;
        st[sp++] = r;
        break;
      }
      case OP_ISNULL: {
        Val r;
        r.hi = 0;
        r.pad = 0;
        r.valid = 1;
        r.lo = !st[sp - 1].valid;
        st[sp - 1] = r;
        break;
      }
      case OP_ISNOTNULL: {
        Val r;
        r.hi = 0;
        r.pad = 0;
        r.valid = 1;
        r.lo = st[sp - 1].valid != 0;
        st[sp - 1] = r;
        break;
      }
      case OP_NEG: {
        Val a = st[sp - 1];
        if (a.valid) {
          if (op.vclass == VC_INT) a.lo = (uint64_t)wrap_signed((int64_t)(0ull - a.lo), op.wbits);
          else if (op.vclass == VC_DEC) set_i128(a, (i128)((u128)0 - (u128)val_i128(a)));
          else if (op.vclass == VC_FLT) set_f64(a, -val_f64(a));
        }
        st[sp - 1] = a;
        break;
      }
      case OP_LIKE: {  // like.rs:28-41 -> arrow like / nlike (op.sub = negated); NULL if either side is NULL
        const Val b = st[--sp];
        const Val a = st[sp - 1];
        Val r;
        r.hi = 0;
        r.pad = 0;
        r.valid = a.valid & b.valid;
        r.lo = 0;
        if (r.valid) {
          const bool m = like_match((const unsigned char*)a.lo, (int64_t)a.hi, (const unsigned char*)b.lo, (int64_t)b.hi);
          r.lo = (uint64_t)(m != (op.sub != 0));
        }
        st[sp - 1] = r;
        break;
      }
      case OP_EXTRACT: {  // functions/datetime/extract.rs: date_part(YEAR | MONTH | DAY) cast to Int64
        Val a = st[sp - 1];
        if (a.valid) {
          int64_t days = (int64_t)a.lo;
          if (op.from_id == QGPU_T_DATE64) days = (days >= 0 ? days : days - 86399999) / 86400000;  // ms -> days (floor)
          int64_t y;
          unsigned m, d;
          civil_from_days(days, &y, &m, &d);
          a.lo = (uint64_t)(op.sub == 0 ? y : (op.sub == 1 ? (int64_t)m : (int64_t)d));
          a.hi = 0;
        }
        st[sp - 1] = a;
        break;
      }
      default: break;
    }
  }
  return st[0];
}

}  // namespace qgpu
