// Finalisation of one aggregate value from its accumulator state -- shared by the generic finalise kernel
// (ops.cu: k_agg_finalize) and the single-CTA epilogue of the dense fused path (epilogue.cu).
//   SumAccumulator::evaluate     qurious/src/physical/expr/aggregate/sum.rs:71-77     (NULL when no row was seen)
//   AvgAccumulator::evaluate     qurious/src/physical/expr/aggregate/avg.rs:62-130    (Float64: sum / n; Decimal: sum * 10^(ts - ss) / n, truncating)
//   Min/Max (PrimitiveAccumulator) .../aggregate/{min,max}.rs                           (start value = NATIVE::MAX / MIN, never NULL: quirk Q4)
//   CountAccumulator             .../aggregate/count.rs:40-48
#pragma once
#include "interp.cuh"
#include "ops.h"

namespace qgpu {

struct FinSpec {
  int op;          // qgpu_agg_op
  int kind;        // AccKind
  int out_phys;    // Phys of the output column
  int sum_scale;   // AVG decimal
  int target_scale;
  int target_prec;
  int compat_avg;  // validate the pre-division value (reference quirk Q5)
  int no_input;    // ungrouped aggregate over zero input batches
  const unsigned long long* lo;  // accumulator word of group g = lo[g * stride] (stride 0 is read as 1)
  const unsigned long long* hi;  // may be null: sign extension of lo (0 for f64 / unsigned kinds)
  const unsigned long long* cnt;
  long long stride;
  void* out;
  uint32_t* out_valid;
};

__device__ __forceinline__ double fin_key_to_f64(long long k) {
  union { long long i; double d; } c;
  c.i = k ^ (long long)(((unsigned long long)(k >> 63)) >> 1);
  return c.d;
}

// high word of an accumulator: stored, or (hi == nullptr) the sign extension of lo (0 for f64 / unsigned kinds)
__device__ __forceinline__ unsigned long long fin_hi(const FinSpec& f, int64_t gi, unsigned long long lo) {
  if (f.hi) return f.hi[gi];
  if (f.kind == AK_SUM_F64 || f.kind == AK_MIN_F64 || f.kind == AK_MAX_F64 || f.kind == AK_MIN_U64 || f.kind == AK_MAX_U64) return 0;
  return ((long long)lo < 0) ? ~0ull : 0ull;
}

// value of aggregate `f` for group g -> (lo, hi); returns its validity.  Errors are raised into *err.
__device__ __forceinline__ bool fin_value(const FinSpec& f, int64_t g, int* err, unsigned long long* out_lo, unsigned long long* out_hi) {
  const int64_t gi = g * (f.stride ? f.stride : 1);
  const unsigned long long cnt = f.cnt[gi];
  unsigned long long lo = 0, hi = 0;
  bool valid = false;
  switch (f.op) {
    case QGPU_AGG_COUNT:
      lo = cnt;
      valid = true;
      break;
    case QGPU_AGG_SUM:
      valid = cnt > 0;
      lo = f.lo[gi];
      hi = fin_hi(f, gi, lo);
      break;
    case QGPU_AGG_MIN:
    case QGPU_AGG_MAX:
      // an all-NULL input leaves the type's MAX/MIN start value, not NULL
      valid = !f.no_input;
      lo = f.lo[gi];
      hi = fin_hi(f, gi, lo);
      if (f.kind == AK_MIN_F64 || f.kind == AK_MAX_F64) {
        union { unsigned long long u; double d; } c;
        c.d = fin_key_to_f64((long long)lo);
        lo = c.u;
      }
      break;
    case QGPU_AGG_AVG:
      if (cnt > 0) {
        if (f.kind == AK_SUM_F64) {
          union { unsigned long long u; double d; } c;
          c.u = f.lo[gi];
          c.d = c.d / (double)cnt;
          lo = c.u;
          valid = true;
        } else {
          // avg.rs:89-116: value = sum * 10^(target_scale - sum_scale) (checked); result = value / count
          const unsigned long long slo = f.lo[gi];
          i128 sum = (i128)(((u128)fin_hi(f, gi, slo) << 64) | (u128)slo);
          i128 mul = pow10_i128(f.target_scale - f.sum_scale);
          i128 value = sum * mul;
          bool ovf = sum != 0 && value / mul != sum;
          if (!ovf && f.compat_avg && !dec_fits_precision(value, f.target_prec)) ovf = true;
          if (ovf) {
            // the reference yields a NULL of type Decimal128(38,10) which then fails the schema check
            if (f.compat_avg) raise_err(err, EE_DEC_PRECISION);
            else if (sum != 0 && value / mul != sum) raise_err(err, EE_OVERFLOW);
          } else {
            i128 q = value / (i128)cnt;
            lo = (unsigned long long)(u128)q;
            hi = (unsigned long long)((u128)q >> 64);
            valid = true;
          }
        }
      }
      break;
  }
  if (!valid) lo = hi = 0;
  *out_lo = lo;
  *out_hi = hi;
  return valid;
}

__device__ __forceinline__ void fin_store(const FinSpec& f, int64_t pos, unsigned long long lo, unsigned long long hi) {
  switch (f.out_phys) {
    case PH_I8: case PH_U8: ((uint8_t*)f.out)[pos] = (uint8_t)lo; break;
    case PH_I16: case PH_U16: ((uint16_t*)f.out)[pos] = (uint16_t)lo; break;
    case PH_I32: case PH_U32: ((uint32_t*)f.out)[pos] = (uint32_t)lo; break;
    case PH_I64: case PH_U64: case PH_F64: ((unsigned long long*)f.out)[pos] = lo; break;
    case PH_F32: {
      union { unsigned long long u; double d; } c;
      c.u = lo;
      ((float*)f.out)[pos] = (float)c.d;
      break;
    }
    case PH_I128: ((ulonglong2*)f.out)[pos] = make_ulonglong2(lo, hi); break;
    default: break;
  }
}

}  // namespace qgpu
