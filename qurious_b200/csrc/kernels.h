// Host-callable wrappers around the generic (always-correct) CUDA kernels in kernels.cu.
#pragma once
#include "expr.h"
#include "qgpu_internal.h"

namespace qgpu {

// ---- scan / selection ---------------------------------------------------------------------------
// out[i] = sum(in[0..i)); returns the grand total (synchronises).
int64_t exclusive_scan_i64(Ctx* ctx, const int64_t* in, int64_t* out, int64_t n);

// ---- expression evaluation ------------------------------------------------------------------------
// Evaluate `c` over every row of `v` and materialise the result as a new device column.
DColP eval_to_column(Ctx* ctx, Compiled& c, const View& v);
// Evaluate boolean predicate `c`; return the (ascending) row positions of `v` whose predicate is
// TRUE and not NULL (arrow filter semantics).
IdxP eval_filter(Ctx* ctx, Compiled& c, const View& v);

// ---- gathers ------------------------------------------------------------------------------------
IdxP compose_idx(Ctx* ctx, const IdxP& inner, const IdxP& outer);  // r[i] = outer[i]<0 ? -1 : inner[outer[i]]
LazyCol apply_selection(Ctx* ctx, const LazyCol& col, const IdxP& sel, std::vector<std::pair<IdxP, IdxP>>* cache);
DColP take_column(Ctx* ctx, const DCol& base, const int64_t* idx, int64_t n, bool idx_may_have_null = true);
// Column in canonical Arrow layout (narrowed decimals widened, gathers applied).
DColP materialize_arrow(Ctx* ctx, const LazyCol& col, int64_t n);
// Column with gathers applied but physical narrowing kept (internal consumers).
DColP materialize(Ctx* ctx, const LazyCol& col, int64_t n);
View apply_selection_view(Ctx* ctx, const View& v, const IdxP& sel);

// ---- ingest helpers -------------------------------------------------------------------------------
// Try to narrow a 16 B Decimal128 column to int64; returns nullptr if some value does not fit.
DColP try_narrow_decimal(Ctx* ctx, const DCol& wide);
void ensure_stats(Ctx* ctx, DCol& col);  // min/max over non-null values (ints, dates, decimals)
bool stats_to_device(Ctx* ctx, const DCol& col, long long* dev_mn_mx);  // the same into device words, no host round trip
void copy_bits(Ctx* ctx, uint32_t* dst, int64_t dst_bit, const uint32_t* src, int64_t src_bit, int64_t n);
void fill_bits(Ctx* ctx, uint32_t* dst, int64_t dst_bit, int64_t n, bool value);
void rebase_offsets(Ctx* ctx, int32_t* dst, const int32_t* src, int64_t n_plus_1, int64_t add);
int64_t count_set_bits(Ctx* ctx, const uint32_t* bits, int64_t n);
IdxP iota_idx(Ctx* ctx, int64_t n);

}  // namespace qgpu
