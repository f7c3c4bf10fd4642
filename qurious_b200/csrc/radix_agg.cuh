// radix_agg.cuh -- included by fused.cu (inside namespace qgpu, after the FParams helpers).
//
// High-cardinality group-by (BASELINE.json configs[3]: 1 B rows / 100 M groups) as a RADIX-PARTITIONED aggregate:
// the HBM-resident hash table of FM_HASH costs one random 64 B read-modify-write (and ~7 L2 atomics) per row once
// the table outgrows L2; here the rows are instead streamed twice through a partitioning pass until every bucket
// holds few enough groups for a SHARED-MEMORY table, so that all random accesses stay on chip:
//
//   k_radix_hist1[_tma]  scan (predicate, packed key) -> 256-bin histogram of the hash's top byte + HyperLogLog sketch
//                    (the group-count estimate picks level-2 fan-out, table capacity and output sizes)
//   k_radix_scan1    bucket offsets / cursors / bucket-aligned tile list
//   k_radix_scatter[_tma]<1>  scan again, materialise tuples (hash of the packed key, operand values) and scatter them
//                    into the 256 level-1 buckets: tile of 4096 tuples -> shared-memory counting sort (unordered ranks
//                    from shared atomics) -> one global reservation per (tile, bucket) -> coalesced run write-out
//   k_radix_hist2    per level-1 bucket: histogram of the next b2 hash bits
//   k_radix_scan2    offsets of the 256 << b2 final buckets
//   k_radix_scatter[_tma]<2>  same scatter, tuples -> final buckets
//   k_radix_agg_list<NV>  one CTA per final bucket: the bucket's operand values bulk-copied into shared memory, key table
//                    claimed with 64-bit CAS, every row pushed onto its slot's linked list (native 32-bit shared
//                    atomicExch), one thread walks each group's list, reduces it in registers and writes the group
//                    (k_radix_agg<NV>, QGPU_RADIX_AGG=sort: the earlier form that ranks, scans and re-stages the rows)
//                    (key columns decoded from the inverted hash word; when every aggregate is a copy / sign extension /
//                    f64 mean of an accumulator the result columns themselves -- no finalise pass)
// The _tma kernels are TMA pipelines (input tiles staged in shared memory by 1-D bulk copies one tile ahead); they run
// whenever their stages fit the shared memory (tuples of <= 2 operand values at level 2), the register-staged ones otherwise.
//
// Semantics are those of FM_HASH (hash.rs:45-107,138-170 with grouping by key equality);
// output order is the bucket order (the reference's order is unspecified, quirk Q2).  Anything that does not fit
// (estimate wrong, skewed bucket overflowing its table) raises a flag and the caller re-runs FM_HASH.

constexpr int R_NT = 256;                 // threads per CTA of the partition kernels (== F_NT: load_rows mapping)
constexpr int R_SUB = 4;                  // sub-tiles of F_T rows
constexpr int R_SNT = 512;                // threads per CTA of the scatter kernels
constexpr int R_SPT = R_SUB * F_NT / R_SNT;  // sub-tiles per thread
static_assert(R_SNT == 512 && R_SPT == 2, "scatter: one histogram bin per thread, two sub-tiles per thread");
constexpr int R_T = F_T * R_SUB;          // tuples per tile
constexpr int R_B1 = 8;                   // level-1 digit: top 8 hash bits (9 + 8 instead of 8 + 9 was measured: level 1 +4.5 ms, level 2 -3.1 ms per 1 B rows)
constexpr int R_P1 = 1 << R_B1;
constexpr int R_MAXB2 = 9;
constexpr int R_MAXCOMP = 6;              // tuple components (code + operand values) that fit the sort tile
constexpr int R_HLL_BITS = 12;
constexpr int R_HLL_M = 1 << R_HLL_BITS;
constexpr int R_HLL_SAMPLE = 8;          // the sketch sees the keys whose hash bits 52..54 are zero
constexpr int R_AGG_NT = 1024;            // threads per CTA of the final pass
constexpr int R_U = 8;                    // final pass: global loads in flight per thread
static_assert(R_AGG_NT == 1024, "the final pass scans 32 warp totals with one warp");
static_assert(R_NT == F_NT, "load_rows_g uses the F_NT row mapping");

// The tuples carry fmix64(code), not the code: every later pass (level-2 histogram and scatter, final pass: digit, table
// slot, probe step, key equality) works on the hash itself, and fmix64 is a bijection -- the final pass inverts it once per
// GROUP when it decodes the key columns.
__device__ __forceinline__ uint64_t r_unmix64(uint64_t x) {
  x ^= x >> 33;
  x *= 0x9cb4b2f8129337dbULL;  // inverse of 0xc4ceb9fe1a85ec53 modulo 2^64
  x ^= x >> 33;
  x *= 0x4f74430c22a54005ULL;  // inverse of 0xff51afd7ed558ccd
  x ^= x >> 33;
  return x;
}

struct RComp {           // operand value = coef * prod(a_i + b_i * x_i) (int64, proven not to overflow) or raw f64 bits
  int32_t is_f64, n_factors;
  int64_t coef;
  FFactor f[F_MAXF];
};

struct RParams {
  int32_t n_comp, b2, cap, n_accs;
  int32_t row_cap, pad0;     // final pass: rows a bucket may hold (staging area)
  int32_t comp_of[F_MAXA];   // accumulator -> tuple component (>= 1)
  int32_t kind_of[F_MAXA];
  int32_t cmask[R_MAXCOMP - 1];      // operand value -> kinds reduced over it (bit FK_*)
  int32_t acc_at[R_MAXCOMP - 1][4];  // (operand value, kind) -> accumulator
  RComp comp[R_MAXCOMP - 1]; // component c >= 1 is comp[c - 1]
  unsigned long long* tup_a[R_MAXCOMP];
  unsigned long long* tup_b[R_MAXCOMP];
  unsigned int* hist1;        // [256]
  unsigned int* hll;          // [4096]
  unsigned long long* off1;   // [257]
  unsigned long long* cur1;   // [256]
  unsigned int* tpre;         // [257] first tile of every level-1 bucket (tiles of R_T tuples, bucket-aligned)
  unsigned int* hist2;        // [256 << b2]
  unsigned long long* off2;   // [(256 << b2) + 1]
  unsigned long long* cur2;   // [256 << b2]
  unsigned long long* out_acc[F_MAXA];
  unsigned long long* out_cnt;
  unsigned long long* n_out;
  int64_t out_cap;
  int* overflow;
  // multi-GPU exchange (world > 1): level-1 bucket b belongs to rank b % world; the level-1 scatter writes its tuples
  // straight into the owner's tuple arrays over NVLink (peer_a[rank][component]; the local rank's entry is tup_a)
  // final pass writing the result columns itself (radix_tail: every aggregate is a plain copy / sign extension / f64 mean
  // of an accumulator): accumulator k -> up to two result columns, COUNT columns, key columns decoded from the code
  int32_t direct, n_cnt_dst;
  // tuples of exactly two operand values (TMA scatter on both levels): the two values of a tuple travel side by side as ONE
  // 16-byte element of array 1 (array 2 is unused) -- a run of n tuples is one n * 16 B piece instead of two n * 8 B pieces:
  // fewer, fuller lines per store instruction (the scatter is bound by its L1 -> L2 write requests), one 16-byte load per
  // row in the final pass
  int32_t pair12, pad_pair;
  int32_t fin_mode[F_MAXA][2];   // 0 = none, 1 = 8-byte copy, 2 = 16-byte sign extension (Decimal128), 3 = f64 sum / count
  void* fin_dst[F_MAXA][2];
  unsigned long long* cnt_dst[4];
  int32_t world, nopf;  // nopf: experiment bits (QGPU_RADIX_NOPF) -- 1 = L2 prefetch in scatter<1> ON, 2 / 4 = no L2 prefetch in scatter<2> / k_radix_agg, 8 = one-pass probing in k_radix_agg
  unsigned long long* peer_a[8][R_MAXCOMP];
};

// guarded variant of load_rows for operands read straight from global memory (no staged tile behind the tail); the width
// is dispatched once per call, not once per row
template <typename T>
__device__ __forceinline__ void load_rows_t(const unsigned char* col, int tid, int rows, int64_t (&x)[F_R]) {
  const T* c = (const T*)col;
  if (rows == F_T) {
#pragma unroll
    for (int j = 0; j < F_R; ++j) x[j] = (int64_t)c[j * F_NT + tid];
  } else {
#pragma unroll
    for (int j = 0; j < F_R; ++j) {
      const int r = j * F_NT + tid;
      x[j] = r < rows ? (int64_t)c[r] : 0;
    }
  }
}
__device__ __forceinline__ void load_rows_g(const unsigned char* col, uint32_t wk, int tid, int rows, int64_t (&x)[F_R]) {
  switch (wk) {
    case 8: load_rows_t<int64_t>(col, tid, rows, x); break;
    case 4: load_rows_t<int32_t>(col, tid, rows, x); break;
    case 4 | 256: load_rows_t<uint32_t>(col, tid, rows, x); break;
    case 2: load_rows_t<int16_t>(col, tid, rows, x); break;
    case 2 | 256: load_rows_t<uint16_t>(col, tid, rows, x); break;
    case 1: load_rows_t<int8_t>(col, tid, rows, x); break;
    default: load_rows_t<uint8_t>(col, tid, rows, x); break;
  }
}
__device__ __forceinline__ const unsigned char* gcol(const FParams& p, int col, int64_t row0) {
  return p.cols[col].ptr + (size_t)row0 * p.cols[col].width;
}

// predicate + packed key of the F_R rows this thread owns in the sub-tile starting at row0
__device__ __forceinline__ uint32_t r_pass_code(const FParams& p, int64_t row0, int rows, int tid, uint64_t (&code)[F_R]) {
  uint32_t pass = 0;
#pragma unroll
  for (int j = 0; j < F_R; ++j)
    if (j * F_NT + tid < rows) pass |= 1u << j;
#pragma unroll 1
  for (int k = 0; k < p.n_pred; ++k) {
    int64_t x[F_R];
    load_rows_g(gcol(p, p.pred[k].col, row0), p.pred[k].wk, tid, rows, x);
    const uint64_t lo = (uint64_t)p.pred[k].lo, span = p.pred[k].span;
#pragma unroll
    for (int j = 0; j < F_R; ++j)
      if (((uint64_t)x[j] - lo) > span) pass &= ~(1u << j);
  }
#pragma unroll
  for (int j = 0; j < F_R; ++j) code[j] = 0;
#pragma unroll 1
  for (int k = 0; k < p.n_keys; ++k) {
    int64_t x[F_R];
    load_rows_g(gcol(p, p.keys[k].col, row0), p.keys[k].wk, tid, rows, x);
    const uint64_t base = (uint64_t)p.keys[k].base, mult = p.keys[k].mult;
    if (mult == 1) {
#pragma unroll
      for (int j = 0; j < F_R; ++j) code[j] += (uint64_t)x[j] - base;
    } else {
#pragma unroll
      for (int j = 0; j < F_R; ++j) code[j] += ((uint64_t)x[j] - base) * mult;
    }
  }
  return pass;
}

__global__ void __launch_bounds__(R_NT) k_radix_hist1(const __grid_constant__ FParams p, unsigned int* __restrict__ hist1,
                                                      unsigned int* __restrict__ hll) {
  __shared__ unsigned int sh[R_P1];
  __shared__ unsigned int sl[R_HLL_M];
  const int tid = threadIdx.x;
  for (int i = tid; i < R_P1; i += R_NT) sh[i] = 0;
  for (int i = tid; i < R_HLL_M; i += R_NT) sl[i] = 0;
  __syncthreads();
  for (int64_t t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
    const int64_t row0 = t * F_T;
    const int rows = (int)min((int64_t)F_T, p.n_rows - row0);
    uint64_t code[F_R];
    const uint32_t pass = r_pass_code(p, row0, rows, tid, code);
#pragma unroll
    for (int j = 0; j < F_R; ++j) {
      if (!((pass >> j) & 1)) continue;
      const uint64_t h = fmix64(code[j]);
      atomicAdd(&sh[h >> (64 - R_B1)], 1u);
      // HyperLogLog over the keys whose hash bits 52..54 are zero (a 1/8 sample of the KEY space: a key is in or out
      // with all of its rows, so the estimate is distinct / 8): register = low 12 bits, rank = leading zeros of bits
      // 12..51 (40 bits) + 1
      if ((h >> 52) & (R_HLL_SAMPLE - 1)) continue;
      const uint64_t w = (h >> R_HLL_BITS) & ((1ull << 40) - 1);
      const unsigned rho = w ? (unsigned)(__clzll((long long)w) - 24 + 1) : 41u;
      atomicMax(&sl[h & (R_HLL_M - 1)], rho);
    }
  }
  __syncthreads();
  for (int i = tid; i < R_P1; i += R_NT)
    if (sh[i]) atomicAdd(&hist1[i], sh[i]);
  for (int i = tid; i < R_HLL_M; i += R_NT)
    if (sl[i]) atomicMax(&hll[i], sl[i]);
}

// one CTA of 256 threads: thread b owns level-1 bucket b
__global__ void __launch_bounds__(R_P1) k_radix_scan1(const unsigned int* __restrict__ hist1, unsigned long long* __restrict__ off1,
                                                      unsigned long long* __restrict__ cur1, unsigned int* __restrict__ tpre) {
  __shared__ unsigned long long a[R_P1];
  __shared__ unsigned int b[R_P1];
  const int t = threadIdx.x;
  const unsigned int h = hist1[t];
  a[t] = h;
  b[t] = (h + R_T - 1) / R_T;
  __syncthreads();
  for (int d = 1; d < R_P1; d <<= 1) {
    const unsigned long long x = t >= d ? a[t - d] : 0;
    const unsigned int y = t >= d ? b[t - d] : 0;
    __syncthreads();
    a[t] += x;
    b[t] += y;
    __syncthreads();
  }
  off1[t + 1] = a[t];
  tpre[t + 1] = b[t];
  cur1[t] = a[t] - h;
  if (t == 0) {
    off1[0] = 0;
    tpre[0] = 0;
  }
}

// tile -> (level-1 bucket, first tuple, tuple count) through the bucket-aligned tile list
__device__ __forceinline__ void r_tile_of(const RParams& r, unsigned int tile, int* b1, int64_t* base, int* rows) {
  int lo = 0, hi = R_P1;  // largest b with tpre[b] <= tile
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (r.tpre[mid] <= tile) lo = mid;
    else hi = mid;
  }
  *b1 = lo;
  const int64_t start = (int64_t)r.off1[lo] + (int64_t)(tile - r.tpre[lo]) * R_T;
  *base = start;
  *rows = (int)min((int64_t)R_T, (int64_t)r.off1[lo + 1] - start);
}

// LEVEL 1: input rows -> tuples in level-1 buckets.  LEVEL 2: level-1 tuples -> final buckets.
// dynamic shared memory: sorted[n_comp][R_T] u64 | gbase[512] u64 | hist[512] u32 | toff[512] u32 | bin[R_T] u16
template <int LEVEL>
__global__ void __launch_bounds__(R_SNT, 2) k_radix_scatter(const __grid_constant__ FParams p, const __grid_constant__ RParams r) {
  extern __shared__ __align__(16) unsigned char rsm[];
  unsigned long long* sorted = (unsigned long long*)rsm;
  unsigned long long* gbase = sorted + (size_t)r.n_comp * R_T;
  unsigned int* hist = (unsigned int*)(gbase + 512);
  unsigned int* toff = hist + 512;
  unsigned short* bin = (unsigned short*)(toff + 512);
  __shared__ unsigned int warp_tot[R_SNT / 32];
  // 512 threads: threads 0..255 own sub-tiles 0 and 1 of the 4096-tuple tile, threads 256..511 sub-tiles 2 and 3
  // (row mapping inside a sub-tile = load_rows_g's: row j * 256 + t256)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, t256 = tid & (F_NT - 1), s_first = (tid >> 8) * R_SPT;
  const int NB = LEVEL == 1 ? R_P1 : (1 << r.b2);
  const int shift = LEVEL == 1 ? (64 - R_B1) : (64 - R_B1 - r.b2);
  const unsigned int n_tiles = LEVEL == 1 ? (unsigned int)((p.n_rows + R_T - 1) / R_T) : r.tpre[R_P1];
  unsigned long long* const cursor = LEVEL == 1 ? r.cur1 : r.cur2;
  unsigned long long* const* out = LEVEL == 1 ? r.tup_a : r.tup_b;

  // tiles are taken round-robin (at any time all CTAs append to the same few hundred output runs, which keeps the
  // partially written sectors together in L2); LEVEL 2 walks the bucket-aligned tile list instead of searching it
  int b1 = 0;
  unsigned int b_tile0 = 0, b_tile1 = 0;     // tiles [b_tile0, b_tile1) belong to level-1 bucket b1
  unsigned long long b_off0 = 0, b_off1 = 0; // its tuples
  if (LEVEL == 2) {
    b_tile1 = r.tpre[1];
    b_off1 = r.off1[1];
  }
  for (unsigned int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    int rows;
    int64_t base;
    if (LEVEL == 1) {
      base = (int64_t)t * R_T;
      rows = (int)min((int64_t)R_T, p.n_rows - base);
      // (off by default, QGPU_RADIX_NOPF bit 1 turns it on) pull this CTA's next tile towards L2.  ncu showed the
      // prefetched sectors being fetched from DRAM a second time by the demand loads (7.9 GB read for 4.8 GB of
      // input, L2 read hit rate 24 %); without the prefetch the kernel is 3.6 % faster and reads the input once.
      const int64_t nbase = base + (int64_t)gridDim.x * R_T;
      if (nbase < p.n_rows && tid < p.n_cols && (r.nopf & 1)) {
        const int64_t nrows = min((int64_t)R_T, p.n_rows - nbase);
        const uint32_t bytes = (uint32_t)((nrows * p.cols[tid].width) & ~15ll);
        if (bytes) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gcol(p, tid, nbase)), "r"(bytes) : "memory");
      }
    } else {
      while (t >= b_tile1) {  // next non-empty level-1 bucket
        ++b1;
        b_tile0 = b_tile1;
        b_tile1 = r.tpre[b1 + 1];
        b_off0 = b_off1;
        b_off1 = r.off1[b1 + 1];
      }
      base = (int64_t)b_off0 + (int64_t)(t - b_tile0) * R_T;
      rows = (int)min((int64_t)R_T, (int64_t)b_off1 - base);
      if (t + gridDim.x < b_tile1 && tid < r.n_comp && !(r.nopf & 2)) {  // this CTA's next tile, when it lies in the same bucket
        const int64_t nbase = base + (int64_t)gridDim.x * R_T;
        const uintptr_t a0 = ((uintptr_t)(r.tup_a[tid] + nbase) + 15) & ~(uintptr_t)15;
        const uintptr_t a1 = (uintptr_t)(r.tup_a[tid] + min((int64_t)b_off1, nbase + R_T)) & ~(uintptr_t)15;
        if (a1 > a0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a0), "r"((uint32_t)(a1 - a0)) : "memory");
      }
    }
    for (int i = tid; i < NB; i += R_SNT) hist[i] = 0;
    __syncthreads();
    // ---- phase 1: digit + unordered rank of every tuple (shared atomics).  All global loads of the tile are issued
    //      before the first atomic so that they overlap (16 rows per thread in flight) ---------------------------------
    uint64_t code[R_SPT][F_R];
    uint32_t pr[R_SPT][F_R];
    uint32_t passm[R_SPT];
#pragma unroll
    for (int s = 0; s < R_SPT; ++s) {
      const int srows = max(0, min(F_T, rows - (s_first + s) * F_T));
      if (LEVEL == 1) {
        passm[s] = r_pass_code(p, base + (s_first + s) * F_T, srows, t256, code[s]);
      } else {
        passm[s] = 0;
#pragma unroll
        for (int j = 0; j < F_R; ++j) {
          const int i = j * F_NT + t256;
          code[s][j] = 0;
          if (i < srows) {
            code[s][j] = r.tup_a[0][base + (s_first + s) * F_T + i];
            passm[s] |= 1u << j;
          }
        }
      }
    }
#pragma unroll
    for (int s = 0; s < R_SPT; ++s) {
#pragma unroll
      for (int j = 0; j < F_R; ++j) {
        pr[s][j] = 0xffffffffu;
        if ((passm[s] >> j) & 1) {
          if (LEVEL == 1) code[s][j] = fmix64(code[s][j]);  // from here on the tuple carries the hash
          const unsigned d = (unsigned)(code[s][j] >> shift) & (unsigned)(NB - 1);
          pr[s][j] = (d << 16) | atomicAdd(&hist[d], 1u);
        }
      }
    }
    __syncthreads();
    // ---- exclusive scan of the histogram (one bin per thread) + one global reservation per non-empty bin --------------
    {
      const unsigned h0 = tid < NB ? hist[tid] : 0u;
      unsigned incl = h0;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
      }
      if (lane == 31) warp_tot[warp] = incl;
      __syncthreads();
      unsigned wbase = 0;
#pragma unroll
      for (int w = 0; w < R_SNT / 32; ++w)
        if (w < warp) wbase += warp_tot[w];
      if (tid < NB) {
        toff[tid] = wbase + incl - h0;
        if (h0) gbase[tid] = atomicAdd(&cursor[(LEVEL == 1 ? 0 : (b1 << r.b2)) + tid], (unsigned long long)h0);
      }
    }
    __syncthreads();
    // ---- phase 2: tuples into the shared-memory tile, grouped by bin; per component all 16 loads first -----------------
    uint32_t pos[R_SPT][F_R];
#pragma unroll
    for (int s = 0; s < R_SPT; ++s) {
#pragma unroll
      for (int j = 0; j < F_R; ++j) {
        pos[s][j] = 0xffffffffu;
        if (pr[s][j] != 0xffffffffu) {
          const unsigned d = pr[s][j] >> 16;
          pos[s][j] = toff[d] + (pr[s][j] & 0xffffu);
          sorted[pos[s][j]] = code[s][j];
          bin[pos[s][j]] = (unsigned short)d;
        }
      }
    }
#pragma unroll 1
    for (int c = 1; c < r.n_comp; ++c) {
      int64_t v[R_SPT][F_R];
      if (LEVEL == 1) {
        const RComp& C = r.comp[c - 1];
        if (C.is_f64) {
#pragma unroll
          for (int s = 0; s < R_SPT; ++s)
            load_rows_g(gcol(p, C.f[0].col, base + (s_first + s) * F_T), C.f[0].wk, t256, max(0, min(F_T, rows - (s_first + s) * F_T)), v[s]);
        } else {
#pragma unroll
          for (int s = 0; s < R_SPT; ++s)
#pragma unroll
            for (int j = 0; j < F_R; ++j) v[s][j] = C.coef;
#pragma unroll 1
          for (int f = 0; f < C.n_factors; ++f) {
            int64_t x[R_SPT][F_R];
#pragma unroll
            for (int s = 0; s < R_SPT; ++s)
              load_rows_g(gcol(p, C.f[f].col, base + (s_first + s) * F_T), C.f[f].wk, t256, max(0, min(F_T, rows - (s_first + s) * F_T)), x[s]);
            const int64_t fa = C.f[f].a, fb = C.f[f].b;
#pragma unroll
            for (int s = 0; s < R_SPT; ++s)
#pragma unroll
              for (int j = 0; j < F_R; ++j) v[s][j] *= (fa + fb * x[s][j]);
          }
        }
      } else {
#pragma unroll
        for (int s = 0; s < R_SPT; ++s) {
          const int srows = max(0, min(F_T, rows - (s_first + s) * F_T));
#pragma unroll
          for (int j = 0; j < F_R; ++j) {
            const int i = j * F_NT + t256;
            v[s][j] = i < srows ? (int64_t)r.tup_a[c][base + (s_first + s) * F_T + i] : 0;
          }
        }
      }
#pragma unroll
      for (int s = 0; s < R_SPT; ++s)
#pragma unroll
        for (int j = 0; j < F_R; ++j)
          if (pos[s][j] != 0xffffffffu) sorted[(size_t)c * R_T + pos[s][j]] = (unsigned long long)v[s][j];
    }
    __syncthreads();
    // ---- phase 3: coalesced write-out (consecutive threads = consecutive tuples of one bin) -----------------------
    const unsigned total = toff[NB - 1] + hist[NB - 1];
    for (unsigned i = tid; i < total; i += R_SNT) {
      const unsigned d = bin[i];
      const unsigned long long dest = gbase[d] + (i - toff[d]);
      if (LEVEL == 1 && r.world > 1) {  // peer (or own) memory of the bucket's owner: P2P stores over NVLink
        unsigned long long* const* peer = r.peer_a[d % (unsigned)r.world];
#pragma unroll 1
        for (int c = 0; c < r.n_comp; ++c) peer[c][dest] = sorted[(size_t)c * R_T + i];
      } else {
#pragma unroll 1
        for (int c = 0; c < r.n_comp; ++c) out[c][dest] = sorted[(size_t)c * R_T + i];
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------------------------
// The scatter as a TMA pipeline (the default whenever two input stages of a 4096-tuple tile fit the shared memory; the
// register-staged kernel above remains for wider tuples).  ncu on the kernel above: issue slots 41 % busy, DRAM 50-60 %,
// and the two add up -- with every tile's loads issued by the threads that then wait for them, and two CTAs per SM, the
// memory system idles while a CTA ranks and sorts, and the SM idles while it loads.  Here ONE CTA of 1024 threads per SM
// keeps two input stages: a few lanes issue the 1-D bulk copies (cp.async.bulk -> mbarrier, SASS UBLKCP) of tile i + 1
// before anybody touches tile i, so the input of a whole tile (~96 KB) is in flight during all of tile i's phases.
// The tile is not sorted through shared memory either: the counting sort places only a 16-bit SOURCE INDEX per output
// position; the write-out gathers the tuple from the input stage (LEVEL 1: evaluates the operand values from the staged
// raw columns there) and stores it to its run -- one shared-memory pass over the tuple data instead of three.
//   phase 1  code (LEVEL 1: predicates + packed key) from the stage, digit, unordered rank (native shared atomicAdd)
//   scan     exclusive scan of the bin counts; one global reservation per non-empty bin, its round trip overlapping phase 2
//   phase 2  source index + bin of every output position
//   phase 3  coalesced write-out (consecutive threads = consecutive tuples of one bin)
// dynamic shared memory: full[2] mbarriers (128 B) | stage 0 | stage 1 | src[R_T] u16 | bin[R_T] u16 | delta[512] u64 |
//                        hist[512] u32 | toff[512] u32
constexpr int R2_NT = 1024;
constexpr int R2_PT = R_T / R2_NT;  // tuples per thread and tile
struct RStage {
  uint32_t n_in, stage_bytes;
  uint32_t off[F_MAXC];    // LEVEL 1: staged column c of FParams::cols; LEVEL 2: tuple array c
  uint32_t bytes_per_row[F_MAXC];
  uint32_t col_of[F_MAXC];  // LEVEL 1: copy lane -> column (off / bytes_per_row are indexed by the column)
  uint32_t n_stages, pad;
};
static_assert(F_MAXC >= R_MAXCOMP, "RStage holds columns or components");

// element idx[u] of a staged column, u = 0..R2_PT-1 (width dispatched once)
template <typename T>
__device__ __forceinline__ void lds_rows_t(const unsigned char* col, const int (&idx)[R2_PT], int64_t (&x)[R2_PT]) {
  const T* c = (const T*)col;
#pragma unroll
  for (int u = 0; u < R2_PT; ++u) x[u] = (int64_t)c[idx[u]];
}
__device__ __forceinline__ void lds_rows(const unsigned char* col, uint32_t wk, const int (&idx)[R2_PT], int64_t (&x)[R2_PT]) {
  switch (wk) {
    case 8: lds_rows_t<int64_t>(col, idx, x); break;
    case 4: lds_rows_t<int32_t>(col, idx, x); break;
    case 4 | 256: lds_rows_t<uint32_t>(col, idx, x); break;
    case 2: lds_rows_t<int16_t>(col, idx, x); break;
    case 2 | 256: lds_rows_t<uint16_t>(col, idx, x); break;
    case 1: lds_rows_t<int8_t>(col, idx, x); break;
    default: lds_rows_t<uint8_t>(col, idx, x); break;
  }
}
// packed key of the staged rows idx[] (LEVEL 1)
__device__ __forceinline__ void r2_codes(const FParams& p, const RStage& st, const unsigned char* stage, const int (&idx)[R2_PT],
                                         uint64_t (&code)[R2_PT]) {
#pragma unroll
  for (int u = 0; u < R2_PT; ++u) code[u] = 0;
#pragma unroll 1
  for (int k = 0; k < p.n_keys; ++k) {
    int64_t x[R2_PT];
    lds_rows(stage + st.off[p.keys[k].col], p.keys[k].wk, idx, x);
    const uint64_t base = (uint64_t)p.keys[k].base, mult = p.keys[k].mult;
    if (mult == 1) {  // the first key sits at bit 0: no 64-bit multiply
#pragma unroll
      for (int u = 0; u < R2_PT; ++u) code[u] += (uint64_t)x[u] - base;
    } else {
#pragma unroll
      for (int u = 0; u < R2_PT; ++u) code[u] += ((uint64_t)x[u] - base) * mult;
    }
  }
}

struct R2Tile {
  int64_t base;  // first row (LEVEL 1) / first level-1 tuple (LEVEL 2)
  int rows, b1;
};

template <int LEVEL>
__global__ void __launch_bounds__(R2_NT, 1) k_radix_scatter_tma(const __grid_constant__ FParams p, const __grid_constant__ RParams r,
                                                                const __grid_constant__ RStage st) {
  extern __shared__ __align__(16) unsigned char rsm[];  // (bulk copies and mbarriers need 16 B / 8 B)
  uint64_t* full = (uint64_t*)rsm;
  unsigned char* stage0 = rsm + 128;
  unsigned short* srcidx = (unsigned short*)(stage0 + 2 * (size_t)st.stage_bytes);
  unsigned short* bin = srcidx + R_T;
  unsigned long long* delta = (unsigned long long*)(bin + R_T);
  unsigned int* hist = (unsigned int*)(delta + 512);
  unsigned int* toff = hist + 512;
  __shared__ unsigned int warp_tot[16];
  __shared__ unsigned int tile_total;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int NB = LEVEL == 1 ? R_P1 : (1 << r.b2);
  const int shift = LEVEL == 1 ? (64 - R_B1) : (64 - R_B1 - r.b2);
  const unsigned int n_tiles = LEVEL == 1 ? (unsigned int)((p.n_rows + R_T - 1) / R_T) : r.tpre[R_P1];
  unsigned long long* const cursor = LEVEL == 1 ? r.cur1 : r.cur2;
  unsigned long long* const* out = LEVEL == 1 ? r.tup_a : r.tup_b;

  if (tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  for (int i = tid; i < 512; i += R2_NT) hist[i] = 0;
  __syncthreads();
  const uint64_t l2_policy = l2_evict_first_policy();

  // LEVEL 2 walks the bucket-aligned tile list (tiles [tpre[b], tpre[b + 1]) hold the tuples [off1[b], off1[b + 1]))
  int wb = 0;
  unsigned int w_tile0 = 0, w_tile1 = 0;
  unsigned long long w_off0 = 0, w_off1 = 0;
  if (LEVEL == 2) {
    w_tile1 = r.tpre[1];
    w_off1 = r.off1[1];
  }
  auto describe = [&](unsigned int t) {
    R2Tile d;
    if (LEVEL == 1) {
      d.base = (int64_t)t * R_T;
      d.rows = (int)min((int64_t)R_T, p.n_rows - d.base);
      d.b1 = 0;
    } else {
      while (t >= w_tile1) {  // next non-empty level-1 bucket
        ++wb;
        w_tile0 = w_tile1;
        w_tile1 = r.tpre[wb + 1];
        w_off0 = w_off1;
        w_off1 = r.off1[wb + 1];
      }
      d.base = (int64_t)w_off0 + (int64_t)(t - w_tile0) * R_T;
      d.rows = (int)min((int64_t)R_T, (int64_t)w_off1 - d.base);
      d.b1 = wb;
    }
    return d;
  };
  // lanes 0..n_in-1 of warp 0 each copy one column / tuple array of the tile into stage s.  LEVEL 2 sources of 8-byte
  // elements are only 8 B aligned: the copy starts one tuple early when the tile starts at an odd tuple (the stage is read
  // at + (base & 1)); the 16-byte pair array needs no such shift
  auto issue = [&](const R2Tile& d, int s) {
    if (warp != 0) return;
    const int arr = lane < (int)st.n_in ? (LEVEL == 1 ? (int)st.col_of[lane] : lane) : 0;  // column / tuple array of this lane
    const int a = (LEVEL == 2 && lane < (int)st.n_in && st.bytes_per_row[arr] == 8) ? (int)(d.base & 1) : 0;
    uint32_t bytes = 0;
    if (lane < (int)st.n_in) bytes = ((uint32_t)(d.rows + a) * st.bytes_per_row[arr] + 15u) & ~15u;
    uint32_t total = bytes;
#pragma unroll
    for (int o = 16; o; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
    if (lane == 0) mbar_expect_tx(&full[s], total);
    __syncwarp();  // the expected byte count is registered before any copy can complete
    if (lane < (int)st.n_in) {
      const unsigned char* src = LEVEL == 1 ? p.cols[arr].ptr + (size_t)d.base * st.bytes_per_row[arr]
                                            : (const unsigned char*)r.tup_a[arr] + (size_t)(d.base - a) * st.bytes_per_row[arr];
      bulk_g2s(stage0 + (size_t)s * st.stage_bytes + st.off[arr], src, bytes, &full[s], l2_policy);
    }
  };

  unsigned int t = blockIdx.x;
  R2Tile cur, nxt;
  cur.base = 0; cur.rows = 0; cur.b1 = 0;
  nxt = cur;
  if (t < n_tiles) {
    nxt = describe(t);
    issue(nxt, 0);
  }
  for (unsigned int it = 0; t < n_tiles; ++it, t += gridDim.x) {
    const int s = (int)(it & 1);
    cur = nxt;
    if (t + gridDim.x < n_tiles) {  // stage s ^ 1 was released by the barrier that ended the previous tile
      nxt = describe(t + gridDim.x);
      issue(nxt, s ^ 1);
    }
    mbar_wait(&full[s], (it >> 1) & 1u);
    const unsigned char* stage = stage0 + (size_t)s * st.stage_bytes;
    const int rows = cur.rows;
    const int a = LEVEL == 2 ? (int)(cur.base & 1) : 0;
    // ---- phase 1 ------------------------------------------------------------------------------------------------------
    int idx[R2_PT];
    uint32_t pass = 0;
#pragma unroll
    for (int u = 0; u < R2_PT; ++u) {
      idx[u] = u * R2_NT + tid + a;
      if (u * R2_NT + tid < rows) pass |= 1u << u;
    }
    uint64_t code[R2_PT];
    if (LEVEL == 1) {
#pragma unroll 1
      for (int k = 0; k < p.n_pred; ++k) {
        int64_t x[R2_PT];
        lds_rows(stage + st.off[p.pred[k].col], p.pred[k].wk, idx, x);
        const uint64_t plo = (uint64_t)p.pred[k].lo, span = p.pred[k].span;
#pragma unroll
        for (int u = 0; u < R2_PT; ++u)
          if (((uint64_t)x[u] - plo) > span) pass &= ~(1u << u);
      }
      r2_codes(p, st, stage, idx, code);
    } else {
      const unsigned long long* c0 = (const unsigned long long*)(stage + st.off[0]);
#pragma unroll
      for (int u = 0; u < R2_PT; ++u) code[u] = c0[idx[u]];
    }
    uint32_t pr[R2_PT];
#pragma unroll
    for (int u = 0; u < R2_PT; ++u) {
      pr[u] = 0xffffffffu;
      if ((pass >> u) & 1) {
        const uint64_t h = LEVEL == 1 ? fmix64(code[u]) : code[u];  // level-1 tuples already carry the hash
        const unsigned d = (unsigned)(h >> shift) & (unsigned)(NB - 1);
        pr[u] = (d << 16) | atomicAdd(&hist[d], 1u);
      }
    }
    __syncthreads();
    // ---- scan (threads 0..511: one bin each) + global reservations ---------------------------------------------------------
    unsigned h0 = 0, incl = 0;
    if (tid < 512) {
      h0 = tid < NB ? hist[tid] : 0u;
      hist[tid] = 0;  // for the next tile
      incl = h0;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
      }
      if (lane == 31) warp_tot[warp] = incl;
    }
    __syncthreads();
    unsigned long long reserved = 0;
    unsigned my_toff = 0;
    if (tid < 512) {
      unsigned wbase = 0;
#pragma unroll
      for (int w = 0; w < 16; ++w)
        if (w < warp) wbase += warp_tot[w];
      my_toff = wbase + incl - h0;
      toff[tid] = my_toff;
      if (tid == 511) tile_total = wbase + incl;
      if (h0) reserved = atomicAdd(&cursor[(LEVEL == 1 ? 0 : (cur.b1 << r.b2)) + tid], (unsigned long long)h0);
    }
    __syncthreads();
    // ---- phase 2: source index + bin of every output position ------------------------------------------------------------
#pragma unroll
    for (int u = 0; u < R2_PT; ++u) {
      if (pr[u] != 0xffffffffu) {
        const unsigned d = pr[u] >> 16;
        const unsigned pos = toff[d] + (pr[u] & 0xffffu);
        srcidx[pos] = (unsigned short)(u * R2_NT + tid);
        bin[pos] = (unsigned short)d;
      }
    }
    if (tid < 512) delta[tid] = reserved - my_toff;  // destination of output position i of bin d: delta[d] + i
    __syncthreads();
    // ---- phase 3: write-out ----------------------------------------------------------------------------------------------
    const unsigned total = tile_total;
    unsigned long long dest[R2_PT];
    int jdx[R2_PT];
    unsigned dd[R2_PT];
    uint32_t live = 0;
#pragma unroll
    for (int u = 0; u < R2_PT; ++u) {
      const unsigned i = (unsigned)(u * R2_NT + tid);
      jdx[u] = a;
      dest[u] = 0;
      dd[u] = 0;
      if (i < total) {
        live |= 1u << u;
        dd[u] = bin[i];
        jdx[u] = (int)srcidx[i] + a;
        dest[u] = delta[dd[u]] + i;
      }
    }
    if (LEVEL == 2 && r.pair12) {
      const unsigned long long* in0 = (const unsigned long long*)(stage + st.off[0]);
      const ulonglong2* in1 = (const ulonglong2*)(stage + st.off[1]);
      unsigned long long* o0 = out[0];
      ulonglong2* o1 = (ulonglong2*)out[1];
#pragma unroll
      for (int u = 0; u < R2_PT; ++u)
        if ((live >> u) & 1) o0[dest[u]] = in0[jdx[u]];
#pragma unroll
      for (int u = 0; u < R2_PT; ++u)
        if ((live >> u) & 1) o1[dest[u]] = in1[jdx[u] - a];
    } else if (LEVEL == 2) {
#pragma unroll 1
      for (int c = 0; c < r.n_comp; ++c) {
        const unsigned long long* in = (const unsigned long long*)(stage + st.off[c]);
        unsigned long long* o = out[c];
#pragma unroll
        for (int u = 0; u < R2_PT; ++u)
          if ((live >> u) & 1) o[dest[u]] = in[jdx[u]];
      }
    } else {
      int64_t held[R2_PT];  // pair12: operand value 1, stored together with value 2
#pragma unroll
      for (int u = 0; u < R2_PT; ++u) held[u] = 0;
#pragma unroll 1
      for (int c = 0; c < r.n_comp; ++c) {
        int64_t v[R2_PT];
        if (c == 0) {
          uint64_t cd[R2_PT];
          r2_codes(p, st, stage, jdx, cd);
#pragma unroll
          for (int u = 0; u < R2_PT; ++u) v[u] = (int64_t)fmix64(cd[u]);
        } else {
          const RComp& C = r.comp[c - 1];
          if (C.is_f64) {
            lds_rows(stage + st.off[C.f[0].col], C.f[0].wk, jdx, v);
          } else if (C.n_factors == 1 && C.f[0].plain && C.coef == 1) {  // the bare column
            lds_rows(stage + st.off[C.f[0].col], C.f[0].wk, jdx, v);
          } else {
#pragma unroll
            for (int u = 0; u < R2_PT; ++u) v[u] = C.coef;
#pragma unroll 1
            for (int f = 0; f < C.n_factors; ++f) {
              int64_t x[R2_PT];
              lds_rows(stage + st.off[C.f[f].col], C.f[f].wk, jdx, x);
              const int64_t fa = C.f[f].a, fb = C.f[f].b;
#pragma unroll
              for (int u = 0; u < R2_PT; ++u) v[u] *= (fa + fb * x[u]);
            }
          }
        }
        if (r.pair12 && c == 1) {
#pragma unroll
          for (int u = 0; u < R2_PT; ++u) held[u] = v[u];
          continue;
        }
        if (r.pair12 && c == 2) {
          if (r.world > 1) {
#pragma unroll
            for (int u = 0; u < R2_PT; ++u)
              if ((live >> u) & 1)
                ((ulonglong2*)r.peer_a[dd[u] % (unsigned)r.world][1])[dest[u]] = make_ulonglong2((unsigned long long)held[u], (unsigned long long)v[u]);
          } else {
            ulonglong2* o1 = (ulonglong2*)out[1];
#pragma unroll
            for (int u = 0; u < R2_PT; ++u)
              if ((live >> u) & 1) o1[dest[u]] = make_ulonglong2((unsigned long long)held[u], (unsigned long long)v[u]);
          }
          continue;
        }
        if (r.world > 1) {  // peer (or own) memory of the bucket's owner: P2P stores over NVLink
#pragma unroll
          for (int u = 0; u < R2_PT; ++u)
            if ((live >> u) & 1) r.peer_a[dd[u] % (unsigned)r.world][c][dest[u]] = (unsigned long long)v[u];
        } else {
          unsigned long long* o = out[c];
#pragma unroll
          for (int u = 0; u < R2_PT; ++u)
            if ((live >> u) & 1) o[dest[u]] = (unsigned long long)v[u];
        }
      }
    }
    __syncthreads();  // stage s, src[] and bin[] are free again
  }
}

// Level-1 histogram + sketch as a warp-specialised TMA pipeline (the skeleton of fused_main): the register-staged
// k_radix_hist1 issues a tile's loads and then waits for them (3.1 TB/s).  Here two CTAs per SM keep two to four 4096-row
// stages of the key / predicate columns in flight each: warp 16 is the producer (lane l issues the bulk copy of staged
// column l), warps 0..15 consume (two passes of 2048 rows per tile); full[s]: producer -> consumers (complete_tx bytes), empty[s]: one arrive per consumer warp.
// No CTA-wide barrier per tile (a first version with one ran at 3.0 ms per 1 B rows, slower than the kernel it replaces).
// dynamic shared memory: full[4] + empty[4] mbarriers (128 B) | n_stages stages | hist[R_P1] u32 | hll[R_HLL_M] u32
constexpr int RH_NT = 512;  // consumer threads
__global__ void __launch_bounds__(RH_NT + 32, 2) k_radix_hist1_tma(const __grid_constant__ FParams p, const __grid_constant__ RStage st,
                                                                   unsigned int* __restrict__ hist1, unsigned int* __restrict__ hll) {
  extern __shared__ __align__(16) unsigned char rsm[];  // (bulk copies and mbarriers need 16 B / 8 B)
  uint64_t* full = (uint64_t*)rsm;
  uint64_t* empty = full + 4;
  unsigned char* stage0 = rsm + 128;
  unsigned int* sh = (unsigned int*)(stage0 + (size_t)st.n_stages * st.stage_bytes);
  unsigned int* sl = sh + R_P1;
  const int tid = threadIdx.x, lane = tid & 31;
  const bool producer = tid >= RH_NT;
  const int NS = (int)st.n_stages;
  const unsigned int n_tiles = (unsigned int)((p.n_rows + R_T - 1) / R_T);
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], RH_NT / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  for (int i = tid; i < R_P1; i += RH_NT + 32) sh[i] = 0;
  for (int i = tid; i < R_HLL_M; i += RH_NT + 32) sl[i] = 0;
  __syncthreads();
  const unsigned int my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  if (producer) {
    const uint64_t l2_policy = l2_evict_first_policy();
    const int col = lane < (int)st.n_in ? (int)st.col_of[lane] : 0;
#pragma unroll 1
    for (unsigned int it = 0; it < my_tiles; ++it) {
      const int s = (int)(it % (unsigned)NS);
      const unsigned int use = it / (unsigned)NS;
      if (use > 0) mbar_wait(&empty[s], (use - 1) & 1u);  // every consumer warp released the previous use
      const int64_t base = (int64_t)(blockIdx.x + it * gridDim.x) * R_T;
      const int rows = (int)min((int64_t)R_T, p.n_rows - base);
      uint32_t bytes = 0;
      if (lane < (int)st.n_in) bytes = ((uint32_t)rows * st.bytes_per_row[col] + 15u) & ~15u;
      uint32_t total = bytes;
#pragma unroll
      for (int o = 16; o; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
      if (lane == 0) mbar_expect_tx(&full[s], total);
      __syncwarp();  // the expected byte count is registered before any copy can complete
      if (lane < (int)st.n_in)
        bulk_g2s(stage0 + (size_t)s * st.stage_bytes + st.off[col], p.cols[col].ptr + (size_t)base * st.bytes_per_row[col], bytes, &full[s], l2_policy);
    }
  } else {
#pragma unroll 1
    for (unsigned int it = 0; it < my_tiles; ++it) {
      const int s = (int)(it % (unsigned)NS);
      mbar_wait(&full[s], (it / (unsigned)NS) & 1u);
      const unsigned char* stage = stage0 + (size_t)s * st.stage_bytes;
      const int rows = (int)min((int64_t)R_T, p.n_rows - (int64_t)(blockIdx.x + it * gridDim.x) * R_T);
#pragma unroll 1
      for (int half = 0; half < R_T / (R2_PT * RH_NT); ++half) {
      int idx[R2_PT];
      uint32_t pass = 0;
#pragma unroll
      for (int u = 0; u < R2_PT; ++u) {
        idx[u] = (half * R2_PT + u) * RH_NT + tid;
        if (idx[u] < rows) pass |= 1u << u;
      }
#pragma unroll 1
      for (int k = 0; k < p.n_pred; ++k) {
        int64_t x[R2_PT];
        lds_rows(stage + st.off[p.pred[k].col], p.pred[k].wk, idx, x);
        const uint64_t plo = (uint64_t)p.pred[k].lo, span = p.pred[k].span;
#pragma unroll
        for (int u = 0; u < R2_PT; ++u)
          if (((uint64_t)x[u] - plo) > span) pass &= ~(1u << u);
      }
      uint64_t code[R2_PT];
      r2_codes(p, st, stage, idx, code);
      if (half == R_T / (R2_PT * RH_NT) - 1) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);  // this warp has read its rows of stage s
      }
#pragma unroll
      for (int u = 0; u < R2_PT; ++u) {
        if (!((pass >> u) & 1)) continue;
        const uint64_t h = fmix64(code[u]);
        atomicAdd(&sh[h >> (64 - R_B1)], 1u);
        if ((h >> 52) & (R_HLL_SAMPLE - 1)) continue;  // the sketch's key-space sample (see k_radix_hist1)
        const uint64_t w = (h >> R_HLL_BITS) & ((1ull << 40) - 1);
        const unsigned rho = w ? (unsigned)(__clzll((long long)w) - 24 + 1) : 41u;
        atomicMax(&sl[h & (R_HLL_M - 1)], rho);
      }
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < R_P1; i += RH_NT + 32)
    if (sh[i]) atomicAdd(&hist1[i], sh[i]);
  for (int i = tid; i < R_HLL_M; i += RH_NT + 32)
    if (sl[i]) atomicMax(&hll[i], sl[i]);
}

// level-2 histogram: every CTA owns a contiguous range of tiles so that the shared histogram is flushed only when
// the level-1 bucket changes
__global__ void __launch_bounds__(R_NT) k_radix_hist2(const __grid_constant__ RParams r) {
  __shared__ unsigned int sh[1 << R_MAXB2];
  const int tid = threadIdx.x;
  const int NB = 1 << r.b2;
  const unsigned int n_tiles = r.tpre[R_P1];
  const unsigned int per = (n_tiles + gridDim.x - 1) / gridDim.x;
  const unsigned int t0 = blockIdx.x * per, t1 = min(n_tiles, t0 + per);
  for (int i = tid; i < NB; i += R_NT) sh[i] = 0;
  __syncthreads();
  int cur_b1 = -1;
  for (unsigned int t = t0; t < t1; ++t) {
    int b1, rows;
    int64_t base;
    r_tile_of(r, t, &b1, &base, &rows);
    if (b1 != cur_b1) {
      if (cur_b1 >= 0) {
        __syncthreads();
        for (int i = tid; i < NB; i += R_NT) {
          if (sh[i]) atomicAdd(&r.hist2[(cur_b1 << r.b2) + i], sh[i]);
          sh[i] = 0;
        }
        __syncthreads();
      }
      cur_b1 = b1;
    }
    for (int i = tid; i < rows; i += R_NT) {
      const uint64_t h = r.tup_a[0][base + i];
      atomicAdd(&sh[(unsigned)(h >> (64 - R_B1 - r.b2)) & (unsigned)(NB - 1)], 1u);
    }
  }
  __syncthreads();
  if (cur_b1 >= 0)
    for (int i = tid; i < NB; i += R_NT)
      if (sh[i]) atomicAdd(&r.hist2[(cur_b1 << r.b2) + i], sh[i]);
}

// offsets of the final buckets: the 1 << b2 counters of level-1 bucket b sum to that bucket's size, so bucket b's final
// buckets start at off1[b] and every level-1 bucket is scanned independently -- one CTA of 512 threads each (a single CTA
// scanning all 131072 counters took 0.34 ms per step)
__global__ void __launch_bounds__(1 << R_MAXB2) k_radix_scan2(const unsigned int* __restrict__ hist, int b2, const unsigned long long* __restrict__ off1,
                                                              unsigned long long* __restrict__ off, unsigned long long* __restrict__ cur) {
  __shared__ unsigned int warp_tot[(1 << R_MAXB2) / 32];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5, b = blockIdx.x;
  const int NB = 1 << b2;
  const unsigned int h = t < NB ? hist[((size_t)b << b2) + t] : 0u;
  unsigned int incl = h;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned int o = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += o;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  unsigned int wbase = 0;
#pragma unroll
  for (int w = 0; w < (1 << R_MAXB2) / 32; ++w)
    if (w < warp) wbase += warp_tot[w];
  if (t < NB) {
    const unsigned long long o = off1[b] + (unsigned long long)(wbase + incl - h);
    off[((size_t)b << b2) + t] = o;
    cur[((size_t)b << b2) + t] = o;
  }
  if (b == (int)gridDim.x - 1 && t == 0) off[(size_t)gridDim.x << b2] = off1[gridDim.x];
}

struct RKeys {
  int n_keys;
  int shift[F_MAXK], bits[F_MAXK], width[F_MAXK];
  long long base[F_MAXK];
  void* out[F_MAXK];
};
// packed code -> key k (base_k + bits [shift_k, shift_k + bits_k) of the code), stored at row g of its column
__device__ __forceinline__ void r_store_keys(const RKeys& rk, unsigned long long c, unsigned long long g) {
  for (int k = 0; k < rk.n_keys; ++k) {
    const unsigned long long field = rk.bits[k] >= 64 ? c : ((c >> rk.shift[k]) & ((1ull << rk.bits[k]) - 1ull));
    const long long v = (long long)((unsigned long long)rk.base[k] + field);
    switch (rk.width[k]) {
      case 8: ((long long*)rk.out[k])[g] = v; break;
      case 4: ((int*)rk.out[k])[g] = (int)v; break;
      case 2: ((short*)rk.out[k])[g] = (short)v; break;
      default: ((signed char*)rk.out[k])[g] = (signed char)v; break;
    }
  }
}
// accumulator k of output row o: into the accumulator array (finish_aggregate finalises it) or straight into the result
// column(s) it feeds (SumAccumulator / Min / Max evaluate: the value; AvgAccumulator Float64: sum / n, avg.rs:62-75)
__device__ __forceinline__ void r_emit(const RParams& r, int k, unsigned long long o, unsigned long long v, unsigned c) {
  if (!r.direct) {
    r.out_acc[k][o] = v;
    return;
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int mode = r.fin_mode[k][j];
    if (mode == 1) ((unsigned long long*)r.fin_dst[k][j])[o] = v;
    else if (mode == 2) ((ulonglong2*)r.fin_dst[k][j])[o] = make_ulonglong2(v, (long long)v < 0 ? ~0ull : 0ull);
    else if (mode == 3) ((double*)r.fin_dst[k][j])[o] = __longlong_as_double((long long)v) / (double)c;
  }
}

// Final pass: one CTA per final bucket.  No accumulator atomics: the bucket's rows are counting-sorted by GROUP in
// shared memory and every group is then reduced sequentially in registers by one thread.
//   A  stream the packed codes: claim / find the group's slot in an open-addressing key table (64-bit CAS only for
//      the first row of a group), rank = native 32-bit shared atomicAdd on the slot's row count; remember (slot, rank)
//   S  block scan of the slot row counts -> first staged row of every slot; occupied slots -> output positions
//   B  stream the operand values (coalesced, all operands of a row together) into the staging area at start[slot] + rank
//   R  thread-per-slot: reduce the slot's contiguous staged rows (every operand in the same walk), write the group straight
//      to the output arrays
// dynamic shared memory: stage[row_cap][n_comp - 1] u64 | keys[C1] u64 | cnt[C1] u32 | start[C1] u32 | pk[row_cap] u32 |
//                        occ[C1] u16
template <int NV>
__global__ void __launch_bounds__(R_AGG_NT, 1) k_radix_agg(const __grid_constant__ RParams r, const __grid_constant__ RKeys rk) {
  extern __shared__ __align__(16) unsigned char rsm[];
  const int cap = r.cap, C1 = cap + 1, RC = r.row_cap;
  unsigned long long* stage = (unsigned long long*)rsm;  // 16 B aligned: staged row = NV consecutive words
  unsigned long long* keys = stage + (size_t)NV * RC;
  unsigned int* cnt = (unsigned int*)(keys + C1);
  unsigned int* start = cnt + C1;
  unsigned int* pk = start + C1;
  unsigned short* occ = (unsigned short*)(pk + RC);
  __shared__ unsigned long long warp_tot[R_AGG_NT / 32];
  __shared__ unsigned long long out_base, bucket_total;
  __shared__ int bucket_overflow;
  __shared__ unsigned int n_defer;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_buckets = R_P1 << r.b2;
  const int per = (C1 + R_AGG_NT - 1) / R_AGG_NT;
  const int s0 = min(C1, tid * per), s1 = min(C1, s0 + per);

  for (int b = blockIdx.x; b < n_buckets; b += gridDim.x) {
    const int64_t lo = (int64_t)r.off2[b];
    const int64_t n64 = (int64_t)r.off2[b + 1] - lo;
    if (n64 == 0) continue;  // uniform for the CTA
    if (n64 > RC) {          // skewed bucket: more rows than the staging area holds
      if (tid == 0) *r.overflow = 3;
      continue;
    }
    const int n = (int)n64;
    // the next bucket of this CTA: pull its tuples towards L2 while this one is processed
    if (b + (int)gridDim.x < n_buckets && tid < (r.pair12 ? 2 : r.n_comp) && !(r.nopf & 4)) {
      const int64_t words = (r.pair12 && tid == 1) ? 2 : 1;  // 8-byte words per element of tuple array `tid`
      const int64_t nlo = (int64_t)r.off2[b + gridDim.x] * words;
      const int64_t nn = min((int64_t)r.off2[b + gridDim.x + 1] * words - nlo, (int64_t)RC * words);
      const uintptr_t a0 = ((uintptr_t)(r.tup_b[tid] + nlo) + 15) & ~(uintptr_t)15;   // 16 B aligned address and size
      const uintptr_t a1 = (uintptr_t)(r.tup_b[tid] + nlo + nn) & ~(uintptr_t)15;
      if (a1 > a0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a0), "r"((uint32_t)(a1 - a0)) : "memory");
    }
    for (int i = tid; i < C1; i += R_AGG_NT) {
      keys[i] = F_EMPTY;
      cnt[i] = 0;
    }
    if (tid == 0) bucket_overflow = 0;
    __syncthreads();
    // ---- A: slot + rank of every row (R_U loads in flight per thread: the bucket is read at HBM latency) ------------
    if ((r.nopf & 8) || NV < 2) {  // one-pass version (the deferred queue of the two-pass one needs 1.5 staging columns)
    for (int i0 = 0; i0 < n; i0 += R_U * R_AGG_NT) {
      unsigned long long c8[R_U];
#pragma unroll
      for (int u = 0; u < R_U; ++u) {
        const int i = i0 + u * R_AGG_NT + tid;
        c8[u] = i < n ? __ldcs(&r.tup_b[0][lo + i]) : 0ull;
      }
#pragma unroll
      for (int u = 0; u < R_U; ++u) {
        const int i = i0 + u * R_AGG_NT + tid;
        if (i >= n) continue;
        const unsigned long long code = c8[u];
        int slot = cap;  // the key whose code equals the EMPTY marker owns the extra slot
        if (code != F_EMPTY) {
          const uint64_t h = code;  // the tuple's first word IS the hash (r_unmix64)
          slot = (int)(h & (uint64_t)(cap - 1));
          const int step = (int)((h >> 13) & (uint64_t)(cap - 1)) | 1;  // double hashing: odd step, no clustering tails
          int probes = 0;
          while (true) {
            unsigned long long cur = *(volatile unsigned long long*)&keys[slot];
            if (cur == code) break;
            if (cur == F_EMPTY) {
              cur = atomicCAS(&keys[slot], F_EMPTY, code);
              if (cur == F_EMPTY || cur == code) break;
            }
            slot = (slot + step) & (cap - 1);
            if (++probes >= cap) {  // table full: this bucket holds more groups than the estimate allowed for
              slot = -1;
              break;
            }
          }
        }
        if (slot < 0) {
          bucket_overflow = 1;
          continue;
        }
        pk[i] = (unsigned)slot | (atomicAdd(&cnt[slot], 1u) << 13);
      }
    }
    } else {
    // ---- A (two passes): pass 1 gives every row ONE probe -- convergent, no loop; ~3 of 4 rows settle there (their
    //      group's key already sits in the home slot, or the slot is free).  The others are queued (code + row, in the
    //      still unused staging area) and pass 2 walks their probe chains with dense lanes: the warp-wide while loop of
    //      the one-pass version ran ~3.5 iterations per row for an average chain of 1.3.
    unsigned long long* dcode = stage;                       // [RC] codes of the deferred rows
    unsigned int* drow = (unsigned int*)(stage + RC);        // [RC] their row numbers
    if (tid == 0) n_defer = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += R_U * R_AGG_NT) {
      unsigned long long c8[R_U];
#pragma unroll
      for (int u = 0; u < R_U; ++u) {
        const int i = i0 + u * R_AGG_NT + tid;
        c8[u] = i < n ? __ldcs(&r.tup_b[0][lo + i]) : 0ull;
      }
#pragma unroll
      for (int u = 0; u < R_U; ++u) {
        const int i = i0 + u * R_AGG_NT + tid;
        const unsigned long long code = c8[u];
        int slot = cap;
        bool done = true, live = i < n;
        if (live && code != F_EMPTY) {
          slot = (int)(code & (uint64_t)(cap - 1));
          unsigned long long cur = *(volatile unsigned long long*)&keys[slot];
          if (cur == F_EMPTY) cur = atomicCAS(&keys[slot], F_EMPTY, code);
          done = cur == code || cur == F_EMPTY;
        }
        if (live && done) pk[i] = (unsigned)slot | (atomicAdd(&cnt[slot], 1u) << 13);
        const unsigned dm = __ballot_sync(0xffffffffu, live && !done);
        if (dm) {
          unsigned at = 0;
          if (lane == 0) at = atomicAdd(&n_defer, (unsigned)__popc(dm));
          at = __shfl_sync(0xffffffffu, at, 0) + __popc(dm & ((1u << lane) - 1u));
          if (live && !done) {
            dcode[at] = code;
            drow[at] = (unsigned)i;
          }
        }
      }
    }
    __syncthreads();
    const int nd = (int)n_defer;
    for (int j = tid; j < nd; j += R_AGG_NT) {
      const unsigned long long code = dcode[j];
      const uint64_t h = code;
      const int step = (int)((h >> 13) & (uint64_t)(cap - 1)) | 1;
      int slot = ((int)(h & (uint64_t)(cap - 1)) + step) & (cap - 1);  // the home slot holds another key
      int probes = 1;
      while (true) {
        unsigned long long cur = *(volatile unsigned long long*)&keys[slot];
        if (cur == code) break;
        if (cur == F_EMPTY) {
          cur = atomicCAS(&keys[slot], F_EMPTY, code);
          if (cur == F_EMPTY || cur == code) break;
        }
        slot = (slot + step) & (cap - 1);
        if (++probes >= cap) {
          slot = -1;
          break;
        }
      }
      if (slot < 0) {
        bucket_overflow = 1;
        continue;
      }
      pk[drow[j]] = (unsigned)slot | (atomicAdd(&cnt[slot], 1u) << 13);
    }
    }
    __syncthreads();
    if (bucket_overflow) {
      if (tid == 0) *r.overflow = 1;
      __syncthreads();
      continue;
    }
    // ---- S: (rows, occupied slots) exclusive scan over the slots, `per` consecutive slots per thread ------------------
    unsigned long long mine = 0;  // rows in the low word, occupied slots in the high word
    for (int s = s0; s < s1; ++s) mine += (unsigned long long)cnt[s] + (cnt[s] ? (1ull << 32) : 0ull);
    unsigned long long incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned long long o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {  // second level: exclusive scan of the 32 warp totals
      const unsigned long long wt = warp_tot[lane];
      unsigned long long wi = wt;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long o = __shfl_up_sync(0xffffffffu, wi, d);
        if (lane >= d) wi += o;
      }
      warp_tot[lane] = wi - wt;
      if (lane == 31) {
        bucket_total = wi;
        out_base = atomicAdd(r.n_out, wi >> 32);
      }
    }
    __syncthreads();
    const unsigned long long excl = warp_tot[warp] + incl - mine;
    {
      unsigned run = (unsigned)excl;
      unsigned g = (unsigned)(excl >> 32);
      for (int s = s0; s < s1; ++s) {
        start[s] = run;
        run += cnt[s];
        if (cnt[s]) occ[g++] = (unsigned short)s;   // dense list of the occupied slots
      }
    }
    const unsigned n_occ = (unsigned)(bucket_total >> 32);
    const unsigned long long ob = out_base;
    if (ob + n_occ > (unsigned long long)r.out_cap) {  // uniform
      if (tid == 0) *r.overflow = 2;
      __syncthreads();
      continue;
    }
    __syncthreads();
    // ---- B: operand values into the staging area, grouped by slot; a staged row holds its NV operands side by side ---------
    if (NV > 0) {
      constexpr int UB = NV >= 3 ? 2 : 4;
      for (int i0 = 0; i0 < n; i0 += UB * R_AGG_NT) {
        unsigned long long v8[NV > 0 ? NV : 1][UB];
        if (NV == 2 && r.pair12) {
#pragma unroll
          for (int u = 0; u < UB; ++u) {
            const int i = i0 + u * R_AGG_NT + tid;
            const ulonglong2 q = i < n ? __ldcs((const ulonglong2*)r.tup_b[1] + lo + i) : make_ulonglong2(0ull, 0ull);
            v8[0][u] = q.x;
            v8[NV > 1 ? 1 : 0][u] = q.y;
          }
        } else {
#pragma unroll
          for (int c = 0; c < NV; ++c)
#pragma unroll
            for (int u = 0; u < UB; ++u) {
              const int i = i0 + u * R_AGG_NT + tid;
              v8[c][u] = i < n ? __ldcs(&r.tup_b[c + 1][lo + i]) : 0ull;
            }
        }
#pragma unroll
        for (int u = 0; u < UB; ++u) {
          const int i = i0 + u * R_AGG_NT + tid;
          if (i >= n) continue;
          const unsigned p = pk[i];
          unsigned long long* dst = stage + (size_t)(start[p & 8191u] + (p >> 13)) * NV;
          if (NV == 2) {
            *(ulonglong2*)dst = make_ulonglong2(v8[0][u], v8[NV > 1 ? 1 : 0][u]);
          } else {
#pragma unroll
            for (int c = 0; c < NV; ++c) dst[c] = v8[c][u];
          }
        }
      }
    }
    __syncthreads();
    // ---- R: one thread per GROUP (dense over the lanes) reduces its contiguous staged rows in registers -- all operands in
    //         one walk, SUM / MIN / MAX of the same operand sharing it -- and writes the group; consecutive threads write
    //         consecutive output rows
    int msk[NV > 0 ? NV : 1];
#pragma unroll
    for (int cc = 0; cc < NV; ++cc) msk[cc] = r.cmask[cc];
    for (unsigned g = tid; g < n_occ; g += R_AGG_NT) {
      const int sl = occ[g];
      const unsigned c = cnt[sl];
      const unsigned long long* src = stage + (size_t)start[sl] * NV;
      const unsigned long long o = ob + g;
      r_store_keys(rk, r_unmix64(sl == cap ? F_EMPTY : keys[sl]), o);
      if (!r.direct) r.out_cnt[o] = c;
      else
        for (int j = 0; j < r.n_cnt_dst; ++j) r.cnt_dst[j][o] = c;
      long long sm[NV > 0 ? NV : 1], mn[NV > 0 ? NV : 1], mx[NV > 0 ? NV : 1];
      double fs[NV > 0 ? NV : 1];
#pragma unroll
      for (int cc = 0; cc < NV; ++cc) {
        const long long v0 = (long long)src[cc];
        sm[cc] = mn[cc] = mx[cc] = v0;
        fs[cc] = __longlong_as_double(v0);
      }
      for (unsigned i = 1; i < c; ++i) {
        long long v[NV > 0 ? NV : 1];
        if (NV == 2) {
          const ulonglong2 q = *(const ulonglong2*)(src + (size_t)i * NV);
          v[0] = (long long)q.x;
          v[NV > 1 ? 1 : 0] = (long long)q.y;
        } else {
#pragma unroll
          for (int cc = 0; cc < NV; ++cc) v[cc] = (long long)src[(size_t)i * NV + cc];
        }
#pragma unroll
        for (int cc = 0; cc < NV; ++cc) {
          if (msk[cc] & 8) {
            fs[cc] += __longlong_as_double(v[cc]);
          } else {
            sm[cc] += v[cc];
            mn[cc] = min(mn[cc], v[cc]);
            mx[cc] = max(mx[cc], v[cc]);
          }
        }
      }
#pragma unroll
      for (int cc = 0; cc < NV; ++cc) {
        const int m = msk[cc];
        if (m & 1) r_emit(r, r.acc_at[cc][0], o, (unsigned long long)sm[cc], c);
        if (m & 2) r_emit(r, r.acc_at[cc][1], o, (unsigned long long)mn[cc], c);
        if (m & 4) r_emit(r, r.acc_at[cc][2], o, (unsigned long long)mx[cc], c);
        if (m & 8) r_emit(r, r.acc_at[cc][3], o, (unsigned long long)__double_as_longlong(fs[cc]), c);
      }
    }
    __syncthreads();
  }
}

// Final pass, list form (the default; QGPU_RADIX_AGG=sort selects the form above).  The sorting form reads a bucket three
// times with exposed latency (codes, then the operand values twice over: into the staging area at start[slot] + rank, then
// out of it) around a scan of the slot counts.  Here the operand values of the WHOLE bucket are brought into shared memory
// by 1-D bulk copies (cp.async.bulk -> mbarrier) issued before the first row is looked at -- they stay in arrival order --
// and phase A threads every row onto a per-slot linked list instead of ranking it: prev = atomicExch(&head[slot], row)
// (native 32-bit shared atomic), next[row] = prev.  No start[] scan, no second pass over the values; phase R walks a
// group's list (a dependent 4-byte LDS per row) and reads the values where the copy put them.
//   0  init keys / heads; warp 0 issues the bulk copies of the bucket's value arrays
//   A  (two passes like the sorting form) slot of every row, row pushed onto the slot's list; the deferred queue lives in
//      whatever the copies leave free of the staging area -- rows that find it full walk their chain on the spot
//   S  occupied slots -> dense list + output positions; wait for the copies
//   R  thread-per-group: walk the list, reduce, write the group
// dynamic shared memory: stage (value arrays of this bucket, packed) | keys[C1] u64 | head[C1] u32 | next[row_cap] u32 |
//                        occ[C1] u16   -- within the sorting form's size
constexpr unsigned R_NIL = 0xffffffffu;
template <int NV>
__global__ void __launch_bounds__(R_AGG_NT, 1) k_radix_agg_list(const __grid_constant__ RParams r, const __grid_constant__ RKeys rk) {
  extern __shared__ __align__(16) unsigned char rsm[];
  const int cap = r.cap, C1 = cap + 1, RC = r.row_cap;
  const size_t stage_bytes = (size_t)NV * RC * 8;
  unsigned char* stage = rsm;
  unsigned long long* keys = (unsigned long long*)(rsm + stage_bytes);
  unsigned int* head = (unsigned int*)(keys + C1);
  unsigned int* next = head + C1;
  unsigned short* occ = (unsigned short*)(next + RC);
  __shared__ __align__(8) uint64_t full;
  __shared__ unsigned int warp_tot[R_AGG_NT / 32];
  __shared__ unsigned long long out_base;
  __shared__ unsigned int n_occ_sh;
  __shared__ int bucket_overflow;
  __shared__ unsigned int n_defer;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_buckets = R_P1 << r.b2;
  const int per = (C1 + R_AGG_NT - 1) / R_AGG_NT;
  const int s0 = min(C1, tid * per), s1 = min(C1, s0 + per);
  const bool pairs = NV == 2 && r.pair12;
  const int n_arr = NV == 0 ? 0 : (pairs ? 1 : NV);  // value arrays to copy
  if (tid == 0) {
    mbar_init(&full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  const uint64_t l2_policy = l2_evict_first_policy();
  uint32_t phase = 0;

  for (int b = blockIdx.x; b < n_buckets; b += gridDim.x) {
    const int64_t lo = (int64_t)r.off2[b];
    const int64_t n64 = (int64_t)r.off2[b + 1] - lo;
    if (n64 == 0) continue;  // uniform for the CTA
    if (n64 + 2 > RC) {      // skewed bucket: more rows than the staging area holds (+ the alignment shift of its copies)
      if (tid == 0) *r.overflow = 3;
      continue;
    }
    const int n = (int)n64;
    // value arrays of 8-byte elements are only 8 B aligned: copied from the 16 B boundary below (read at + sh)
    const int sh = pairs ? 0 : (int)(lo & 1);
    const uint32_t arr_bytes = pairs ? (uint32_t)n * 16u : (((uint32_t)(n + sh) * 8u + 15u) & ~15u);
    const size_t used = (size_t)n_arr * arr_bytes;
    // ---- 0: table init; the bucket's values start moving (the barrier that ended the previous bucket ordered its reads of
    //         the staging area before these writes)
    if (warp == 0 && n_arr > 0) {
      if (lane == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&full, (uint32_t)used);
      }
      __syncwarp();
      if (lane < n_arr) {
        const unsigned char* src = pairs ? (const unsigned char*)r.tup_b[1] + (size_t)lo * 16
                                         : (const unsigned char*)(r.tup_b[lane + 1] + (lo - sh));
        bulk_g2s(stage + (size_t)lane * arr_bytes, src, arr_bytes, &full, l2_policy);
      }
    }
    // the next bucket of this CTA: pull its codes towards L2 while this one is processed
    if (b + (int)gridDim.x < n_buckets && tid == 32 && !(r.nopf & 4)) {
      const int64_t nlo = (int64_t)r.off2[b + gridDim.x];
      const int64_t nn = min((int64_t)r.off2[b + gridDim.x + 1] - nlo, (int64_t)RC);
      const uintptr_t a0 = ((uintptr_t)(r.tup_b[0] + nlo) + 15) & ~(uintptr_t)15;
      const uintptr_t a1 = (uintptr_t)(r.tup_b[0] + nlo + nn) & ~(uintptr_t)15;
      if (a1 > a0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a0), "r"((uint32_t)(a1 - a0)) : "memory");
    }
    for (int i = tid; i < C1; i += R_AGG_NT) {
      keys[i] = F_EMPTY;
      head[i] = R_NIL;
    }
    if (tid == 0) {
      bucket_overflow = 0;
      n_defer = 0;
    }
    // deferred queue (hash word + row) in what the copies leave free of the staging area
    unsigned long long* dcode = (unsigned long long*)(stage + ((used + 15) & ~(size_t)15));
    const int dq_cap = (int)((stage_bytes - ((used + 15) & ~(size_t)15)) / 12);
    unsigned int* drow = (unsigned int*)(dcode + dq_cap);
    __syncthreads();
    // ---- A, pass 1: one convergent probe per row; settled rows go onto their slot's list -------------------------------------
    for (int i0 = 0; i0 < n; i0 += R_U * R_AGG_NT) {
      unsigned long long c8[R_U];
#pragma unroll
      for (int u = 0; u < R_U; ++u) {
        const int i = i0 + u * R_AGG_NT + tid;
        c8[u] = i < n ? __ldcs(&r.tup_b[0][lo + i]) : 0ull;
      }
#pragma unroll
      for (int u = 0; u < R_U; ++u) {
        const int i = i0 + u * R_AGG_NT + tid;
        const unsigned long long code = c8[u];
        int slot = cap;  // the key whose hash word equals the EMPTY marker owns the extra slot
        bool done = true, live = i < n;
        if (live && code != F_EMPTY) {
          slot = (int)(code & (uint64_t)(cap - 1));
          unsigned long long cur = *(volatile unsigned long long*)&keys[slot];
          if (cur == F_EMPTY) cur = atomicCAS(&keys[slot], F_EMPTY, code);
          done = cur == code || cur == F_EMPTY;
        }
        if (live && done) next[i] = atomicExch(&head[slot], (unsigned)i);
        const unsigned dm = __ballot_sync(0xffffffffu, live && !done);
        if (dm) {
          unsigned at = 0;
          if (lane == 0) at = atomicAdd(&n_defer, (unsigned)__popc(dm));
          at = __shfl_sync(0xffffffffu, at, 0) + __popc(dm & ((1u << lane) - 1u));
          if (live && !done) {
            if ((int)at < dq_cap) {
              dcode[at] = code;
              drow[at] = (unsigned)i;
            } else {  // no room in the queue (a bucket that nearly fills the staging area): walk the chain here
              const int step = (int)((code >> 13) & (uint64_t)(cap - 1)) | 1;
              int probes = 1;
              slot = (slot + step) & (cap - 1);
              while (true) {
                unsigned long long cur = *(volatile unsigned long long*)&keys[slot];
                if (cur == code) break;
                if (cur == F_EMPTY) {
                  cur = atomicCAS(&keys[slot], F_EMPTY, code);
                  if (cur == F_EMPTY || cur == code) break;
                }
                slot = (slot + step) & (cap - 1);
                if (++probes >= cap) {
                  slot = -1;
                  break;
                }
              }
              if (slot < 0) bucket_overflow = 1;
              else next[i] = atomicExch(&head[slot], (unsigned)i);
            }
          }
        }
      }
    }
    __syncthreads();
    // ---- A, pass 2: the queued rows walk their probe chains with dense lanes ---------------------------------------------------
    const int nd = min((int)n_defer, dq_cap);
    for (int j = tid; j < nd; j += R_AGG_NT) {
      const unsigned long long code = dcode[j];
      const int step = (int)((code >> 13) & (uint64_t)(cap - 1)) | 1;
      int slot = ((int)(code & (uint64_t)(cap - 1)) + step) & (cap - 1);  // the home slot holds another key
      int probes = 1;
      while (true) {
        unsigned long long cur = *(volatile unsigned long long*)&keys[slot];
        if (cur == code) break;
        if (cur == F_EMPTY) {
          cur = atomicCAS(&keys[slot], F_EMPTY, code);
          if (cur == F_EMPTY || cur == code) break;
        }
        slot = (slot + step) & (cap - 1);
        if (++probes >= cap) {
          slot = -1;
          break;
        }
      }
      if (slot < 0) {
        bucket_overflow = 1;
        continue;
      }
      const unsigned i = drow[j];
      next[i] = atomicExch(&head[slot], i);
    }
    __syncthreads();
    if (n_arr > 0) {
      mbar_wait(&full, phase);  // the values have landed (long ago: phase A is most of a bucket's time)
      phase ^= 1u;
    }
    if (bucket_overflow) {
      if (tid == 0) *r.overflow = 1;
      __syncthreads();
      continue;
    }
    // ---- S: occupied slots -> dense list + output positions ---------------------------------------------------------------------
    unsigned mine = 0;
    for (int sl = s0; sl < s1; ++sl) mine += head[sl] != R_NIL ? 1u : 0u;
    unsigned incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      const unsigned wt = warp_tot[lane];
      unsigned wi = wt;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned o = __shfl_up_sync(0xffffffffu, wi, d);
        if (lane >= d) wi += o;
      }
      warp_tot[lane] = wi - wt;
      if (lane == 31) {
        n_occ_sh = wi;
        out_base = atomicAdd(r.n_out, (unsigned long long)wi);
      }
    }
    __syncthreads();
    {
      unsigned g = warp_tot[warp] + incl - mine;
      for (int sl = s0; sl < s1; ++sl)
        if (head[sl] != R_NIL) occ[g++] = (unsigned short)sl;
    }
    const unsigned n_occ = n_occ_sh;
    const unsigned long long ob = out_base;
    if (ob + n_occ > (unsigned long long)r.out_cap) {  // uniform
      if (tid == 0) *r.overflow = 2;
      __syncthreads();
      continue;
    }
    __syncthreads();
    // ---- R: one thread per GROUP walks its list and reduces the rows where the copies put them ------------------------------------
    int msk[NV > 0 ? NV : 1];
#pragma unroll
    for (int cc = 0; cc < NV; ++cc) msk[cc] = r.cmask[cc];
    for (unsigned g = tid; g < n_occ; g += R_AGG_NT) {
      const int sl = occ[g];
      const unsigned long long o = ob + g;
      long long sm[NV > 0 ? NV : 1], mn[NV > 0 ? NV : 1], mx[NV > 0 ? NV : 1];
      double fs[NV > 0 ? NV : 1];
#pragma unroll
      for (int cc = 0; cc < NV; ++cc) {
        sm[cc] = 0;
        mn[cc] = INT64_MAX;
        mx[cc] = INT64_MIN;
        fs[cc] = 0.0;
      }
      unsigned c = 0;
      for (unsigned i = head[sl]; i != R_NIL; i = next[i]) {
        long long v[NV > 0 ? NV : 1];
        if (pairs) {
          const ulonglong2 q = ((const ulonglong2*)stage)[i];
          v[0] = (long long)q.x;
          v[NV > 1 ? 1 : 0] = (long long)q.y;
        } else {
#pragma unroll
          for (int cc = 0; cc < NV; ++cc) v[cc] = (long long)((const unsigned long long*)(stage + (size_t)cc * arr_bytes))[i + sh];
        }
#pragma unroll
        for (int cc = 0; cc < NV; ++cc) {
          if (msk[cc] & 8) {
            fs[cc] += __longlong_as_double(v[cc]);
          } else {
            sm[cc] += v[cc];
            mn[cc] = min(mn[cc], v[cc]);
            mx[cc] = max(mx[cc], v[cc]);
          }
        }
        ++c;
      }
      r_store_keys(rk, r_unmix64(sl == cap ? F_EMPTY : keys[sl]), o);
      if (!r.direct) r.out_cnt[o] = c;
      else
        for (int j = 0; j < r.n_cnt_dst; ++j) r.cnt_dst[j][o] = c;
#pragma unroll
      for (int cc = 0; cc < NV; ++cc) {
        const int m = msk[cc];
        if (m & 1) r_emit(r, r.acc_at[cc][0], o, (unsigned long long)sm[cc], c);
        if (m & 2) r_emit(r, r.acc_at[cc][1], o, (unsigned long long)mn[cc], c);
        if (m & 4) r_emit(r, r.acc_at[cc][2], o, (unsigned long long)mx[cc], c);
        if (m & 8) r_emit(r, r.acc_at[cc][3], o, (unsigned long long)__double_as_longlong(fs[cc]), c);
      }
    }
    __syncthreads();
  }
}

// (A second form of the final pass -- shared memory holding only the group table, every row accumulating straight into
// its slot -- was built and measured in round 2 and removed: 64-bit shared-memory atomics are CAS loops in SASS
// (ATOMS.CAST.SPIN.64), one per accumulator and row made it 44 ms per 1 B rows against 17 ms for the sorting form; with a
// per-slot micro-lock in the row count's top bit (one native ATOMS.OR, plain updates, releasing store) 1024 threads
// working on ~1500 groups collide on half of their rows and it still took 38 ms.  Native 32-bit atomics, which the
// sorting form uses for its ranks, cost about as much as a random LDS: scripts/ubench/rank_ubench.cu.)

// HyperLogLog estimate from the 4096 registers (host side)
static double hll_estimate(const unsigned int* reg) {
  const double m = (double)R_HLL_M;
  double sum = 0;
  int zeros = 0;
  for (int i = 0; i < R_HLL_M; ++i) {
    sum += ldexp(1.0, -(int)reg[i]);
    zeros += reg[i] == 0;
  }
  const double alpha = 0.7213 / (1.0 + 1.079 / m);
  double e = alpha * m * m / sum;
  if (e <= 2.5 * m && zeros > 0) e = m * log(m / (double)zeros);
  return e * R_HLL_SAMPLE;
}
