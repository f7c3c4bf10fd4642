// Device code of the fused scan -> predicate -> aggregate pipeline (see fused.cu for the design): parameter block, PTX
// helpers (mbarrier + 1-D TMA bulk copy), the two tile bodies and the kernel skeleton.  This header is compiled twice:
//   * ahead of time, included by fused.cu (generic interpreted body for every mode, the registered specialised shapes);
//   * at run time by NVRTC (fused_jit.cu, QGPU_JIT defined): the SAME source instantiated for the shape signature of a
//     plan that has no registered kernel -- `SpecBody<signature>` -- so that every DENSE plan runs a specialised body.
// It therefore contains device code only and no host includes.
#pragma once
#ifdef QGPU_JIT
typedef signed char int8_t;
typedef unsigned char uint8_t;
typedef short int16_t;
typedef unsigned short uint16_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
#define INT64_MAX 0x7fffffffffffffffLL
#define INT64_MIN (-INT64_MAX - 1)
namespace qgpu {
typedef __int128 i128;
typedef unsigned __int128 u128;
}
#endif

namespace qgpu {

constexpr int F_T = 1024;   // rows per tile
constexpr int F_NT = 256;   // threads per CTA
constexpr int F_R = F_T / F_NT;
constexpr int F_MAXC = 10, F_MAXP = 8, F_MAXK = 4, F_MAXA = 8, F_MAXF = 3;
constexpr int F_SMEM_MAX = 232448 - 1024;  // 227 KB opt-in limit minus static slack
enum { FK_SUM = 0, FK_MIN = 1, FK_MAX = 2, FK_SUMF = 3 };
enum { FM_DENSE = 0, FM_HASH = 1, FM_PROBE = 2, FM_EMIT = 3, FM_BUILD = 4 };
#define F_EMPTY 0xffffffffffffffffULL

struct FCol {
  const unsigned char* ptr;
  uint32_t width;     // 1, 2, 4, 8
  uint32_t kind;      // 0 signed, 1 unsigned
  uint32_t smem_off;  // offset of this column's tile inside a stage
  uint32_t pad;
};
// every operand reference carries the staged tile's offset and its width/sign (resolved on the host)
struct FPred {
  int32_t col;
  uint32_t off, wk, pad;
  int64_t lo;
  uint64_t span;  // pass iff (uint64)(x - lo) <= span
};
struct FKey {
  int32_t col;
  uint32_t off, wk, pad;
  int64_t base;
  uint64_t mult;  // code += (uint64)(x - base) * mult
};
struct FFactor {
  int32_t col;
  uint32_t off, wk;
  int32_t plain;  // a == 0 && b == 1
  int64_t a, b;   // a + b * x
};
struct FAcc {
  int32_t kind;
  int32_t chain;  // 1: value = previous accumulator's value * factors
  int32_t n_factors;
  int32_t unit;   // !chain && coef == 1 && f[0].plain: value starts as the bare column
  int64_t coef;
  FFactor f[F_MAXF];
};
struct FParams {
  int64_t n_rows, n_tiles;
  int32_t n_cols, n_pred, n_keys, n_accs;
  int32_t mode, stages, dense_groups, carry;
  uint32_t stage_bytes, priv_off;
  uint64_t cap_mask;  // HASH: capacity - 1 (slot `capacity` is reserved for the key whose code == F_EMPTY)
  // DENSE: lo/hi[g * (n_accs + 2) + k];  HASH: lo/hi[k * (capacity + 1) + slot]
  unsigned long long* g_lo;
  unsigned long long* g_hi;
  unsigned long long* n_groups;
  int* abort_flag;
  // FM_PROBE: join table on the BUILD side (unique keys): slot = (tag32 << 32) | (build row + 1), 0 = empty;
  // equality is verified against the immutable build key column; the accumulator slot of a probe row is its
  // matching build row (accumulator stride = acc_stride)
  const unsigned long long* jt_slots;
  uint64_t jt_mask;
  // exact membership bitmap of the build keys over [jt_kmin, jt_kmin + jt_kspan] (null: none): most probe rows
  // miss, and they are rejected here -- a few MB that stay in L1/L2 -- before the random access to the join table
  const uint32_t* jt_bitmap;
  int64_t jt_kmin;
  uint64_t jt_kspan;
  const void* bkey;
  int32_t bkey_width, pad_probe;
  // FM_BUILD: the same table / bitmap, writable
  unsigned long long* jt_wslots;
  uint32_t* jt_wbitmap;
  // accumulator word of (accumulator k, slot g) = g_lo[acc_base + k * acc_kstride + g * acc_gstride]
  //   HASH : array of records {key, acc[0..n_accs), count, first row, pad} (acc_base 1, kstride 1, gstride = record words):
  //          one 64 B record per group => one cache line per row instead of one per accumulator
  //   PROBE: structure of arrays over build rows (acc_base 0, kstride = build rows, gstride 1)
  int64_t acc_base, acc_kstride, acc_gstride;
  // DENSE + specialised body: accumulators in `pack_mask` are bit fields of the row-count word of the thread's private
  // slot (count in bits [0, pack_cnt_bits), accumulator k in [pack_shift[k], pack_shift[k] + pack_bits[k])): the host
  // proved from the column statistics that they are non-negative and that a thread's partial sums fit their fields
  uint32_t pack_mask, pack_cnt_bits;
  uint8_t pack_shift[F_MAXA], pack_bits[F_MAXA];
  // DENSE private table: priv[(slot * priv_lines + line) * F_NT + thread]; line of accumulator / count / first row k
  // (0xff: packed into the count word) and the reverse map
  int32_t priv_lines;
  uint8_t priv_line_of[F_MAXA + 2], priv_k_of[F_MAXA + 2];
  FCol cols[F_MAXC];
  FPred pred[F_MAXP];
  FKey keys[F_MAXK];
  FAcc accs[F_MAXA];
};

// ------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D TMA bulk copy (SASS: UBLKCP / SYNCS)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  const uint32_t a = smem_u32(bar);
  while (!ok) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
  }
}
// streamed-once column tiles are tagged evict-first in L2 so that they do not push out the HBM-resident hash /
// join tables that the same kernel probes at random
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
               : "memory");
}

__device__ __forceinline__ uint64_t fmix64(uint64_t x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return x;
}

// Vectorised operand fetch: the F_R rows this thread owns in the staged tile (row j*F_NT + tid), one
// width/sign dispatch for all of them (the dispatch is warp-uniform).
__device__ __forceinline__ void load_rows(const unsigned char* col, uint32_t wk, int tid, int64_t (&x)[F_R]) {
  switch (wk) {
    case 8:
#pragma unroll
      for (int j = 0; j < F_R; ++j) x[j] = ((const int64_t*)col)[j * F_NT + tid];
      break;
    case 4:
#pragma unroll
      for (int j = 0; j < F_R; ++j) x[j] = (int64_t)((const int32_t*)col)[j * F_NT + tid];
      break;
    case 4 | 256:
#pragma unroll
      for (int j = 0; j < F_R; ++j) x[j] = (int64_t)((const uint32_t*)col)[j * F_NT + tid];
      break;
    case 2:
#pragma unroll
      for (int j = 0; j < F_R; ++j) x[j] = (int64_t)((const int16_t*)col)[j * F_NT + tid];
      break;
    case 2 | 256:
#pragma unroll
      for (int j = 0; j < F_R; ++j) x[j] = (int64_t)((const uint16_t*)col)[j * F_NT + tid];
      break;
    case 1:
#pragma unroll
      for (int j = 0; j < F_R; ++j) x[j] = (int64_t)((const int8_t*)col)[j * F_NT + tid];
      break;
    default:
#pragma unroll
      for (int j = 0; j < F_R; ++j) x[j] = (int64_t)((const uint8_t*)col)[j * F_NT + tid];
      break;
  }
}

__device__ __forceinline__ long long acc_init(const FParams& p, int k) {
  if (k < p.n_accs) {
    const int kind = p.accs[k].kind;
    return kind == FK_MIN ? INT64_MAX : (kind == FK_MAX ? INT64_MIN : 0);
  }
  return k == p.n_accs ? 0 : INT64_MAX;  // row count, first row
}

__device__ __forceinline__ void add128_global(unsigned long long* lo, unsigned long long* hi, unsigned long long vlo,
                                              unsigned long long vhi) {
  const unsigned long long old = atomicAdd(lo, vlo);
  const unsigned long long carry = (old + vlo < old) ? 1ull : 0ull;
  if (vhi + carry) atomicAdd(hi, vhi + carry);
}

template <int KIND>
__device__ __forceinline__ long long comb(long long a, long long b) {
  if (KIND == FK_SUM) return a + b;
  if (KIND == FK_MIN) return min(a, b);
  if (KIND == FK_MAX) return max(a, b);
  return __double_as_longlong(__longlong_as_double(a) + __longlong_as_double(b));
}

// merge bits: 0:(0<-1) 1:(0<-2) 2:(0<-3) 3:(1<-2) 4:(1<-3) 5:(2<-3); rep bits: row j is the representative
// of its group among this thread's rows.  After the merge the representatives own DISTINCT private slots,
// so all loads are issued before the stores.
template <int KIND>
__device__ __forceinline__ void dense_update(long long* const (&a)[F_R], uint32_t koff, int64_t (&v)[F_R], uint32_t merge,
                                             uint32_t rep) {
  if (merge & 1) v[0] = comb<KIND>(v[0], v[1]);
  if (merge & 2) v[0] = comb<KIND>(v[0], v[2]);
  if (merge & 4) v[0] = comb<KIND>(v[0], v[3]);
  if (merge & 8) v[1] = comb<KIND>(v[1], v[2]);
  if (merge & 16) v[1] = comb<KIND>(v[1], v[3]);
  if (merge & 32) v[2] = comb<KIND>(v[2], v[3]);
  long long cur[F_R];
#pragma unroll
  for (int j = 0; j < F_R; ++j) cur[j] = ((rep >> j) & 1) ? a[j][koff] : 0;
#pragma unroll
  for (int j = 0; j < F_R; ++j)
    if ((rep >> j) & 1) a[j][koff] = comb<KIND>(cur[j], v[j]);
}

static_assert(F_R == 4, "the in-thread duplicate-group merge below is written for 4 rows per thread");

#ifndef QGPU_JIT
// ------------------------------------------------------------------------------------------------
// tile body #1: the generic (interpreted) one -- every structural property is a runtime parameter
// ------------------------------------------------------------------------------------------------
template <int MODE>
struct GenericBody {
  static __device__ __forceinline__ void tile(const FParams& p, const unsigned char* stage, const int64_t row0, const int rows,
                                              const int tid, long long* priv, const int NA2) {
    // ---- predicate: range tests, operand-major over the thread's F_R rows -----------------------------
    uint32_t pass = 0;
#pragma unroll
    for (int j = 0; j < F_R; ++j)
      if (j * F_NT + tid < rows) pass |= 1u << j;
#pragma unroll 1
    for (int k = 0; k < p.n_pred; ++k) {
      int64_t x[F_R];
      load_rows(stage + p.pred[k].off, p.pred[k].wk, tid, x);
      const uint64_t lo = (uint64_t)p.pred[k].lo, span = p.pred[k].span;
#pragma unroll
      for (int j = 0; j < F_R; ++j)
        if (((uint64_t)x[j] - lo) > span) pass &= ~(1u << j);
    }
    if (__any_sync(0xffffffffu, pass != 0)) {
      // ---- packed group key ---------------------------------------------------------------------------
      uint64_t code[F_R] = {0, 0, 0, 0};
#pragma unroll 1
      for (int k = 0; k < p.n_keys; ++k) {
        int64_t x[F_R];
        load_rows(stage + p.keys[k].off, p.keys[k].wk, tid, x);
        const uint64_t base = (uint64_t)p.keys[k].base, mult = p.keys[k].mult;
#pragma unroll
        for (int j = 0; j < F_R; ++j) code[j] += ((uint64_t)x[j] - base) * mult;
      }
      const bool p0 = pass & 1, p1 = pass & 2, p2 = pass & 4, p3 = pass & 8;
      int64_t prev[F_R] = {0, 0, 0, 0};

      if (MODE == FM_DENSE) {
        // Rows of this thread that fall into the same group are merged in registers first, so that the
        // read-modify-writes of the representatives touch DISTINCT private slots: all loads can then be
        // issued before the stores (no serialisation on may-alias shared-memory accesses).
        const bool m01 = p0 && p1 && code[0] == code[1];
        const bool m02 = p0 && p2 && code[0] == code[2];
        const bool m03 = p0 && p3 && code[0] == code[3];
        const bool r1 = p1 && !m01;
        const bool m12 = r1 && p2 && code[1] == code[2];
        const bool m13 = r1 && p3 && code[1] == code[3];
        const bool r2 = p2 && !m02 && !m12;
        const bool m23 = r2 && p3 && code[2] == code[3];
        const bool r3 = p3 && !m03 && !m13 && !m23;
        const uint32_t merge = (uint32_t)m01 | ((uint32_t)m02 << 1) | ((uint32_t)m03 << 2) | ((uint32_t)m12 << 3) |
                               ((uint32_t)m13 << 4) | ((uint32_t)m23 << 5);
        const uint32_t rep = (uint32_t)p0 | ((uint32_t)r1 << 1) | ((uint32_t)r2 << 2) | ((uint32_t)r3 << 3);
        long long* a[F_R];
#pragma unroll
        for (int j = 0; j < F_R; ++j)
          a[j] = priv + (((rep >> j) & 1) ? (uint32_t)code[j] * (uint32_t)(NA2 * F_NT) : 0u) + (uint32_t)tid;
#pragma unroll 1
        for (uint32_t k = 0; k < (uint32_t)p.n_accs; ++k) {
          const FAcc& A = p.accs[k];
          const int kind = A.kind;
          int64_t v[F_R];
          if (kind == FK_SUMF || A.unit) {
            load_rows(stage + A.f[0].off, A.f[0].wk, tid, v);
          } else {
#pragma unroll
            for (int j = 0; j < F_R; ++j) v[j] = A.chain ? prev[j] : A.coef;
          }
          if (kind != FK_SUMF) {
#pragma unroll 1
            for (int f = A.unit ? 1 : 0; f < A.n_factors; ++f) {
              int64_t x[F_R];
              load_rows(stage + A.f[f].off, A.f[f].wk, tid, x);
              if (A.f[f].plain) {
#pragma unroll
                for (int j = 0; j < F_R; ++j) v[j] *= x[j];
              } else {
                const int64_t fa = A.f[f].a, fb = A.f[f].b;
#pragma unroll
                for (int j = 0; j < F_R; ++j) v[j] *= (fa + fb * x[j]);
              }
            }
#pragma unroll
            for (int j = 0; j < F_R; ++j) prev[j] = v[j];
          }
          const uint32_t koff = k * F_NT;
          if (kind == FK_SUM) dense_update<FK_SUM>(a, koff, v, merge, rep);
          else if (kind == FK_MIN) dense_update<FK_MIN>(a, koff, v, merge, rep);
          else if (kind == FK_MAX) dense_update<FK_MAX>(a, koff, v, merge, rep);
          else dense_update<FK_SUMF>(a, koff, v, merge, rep);
        }
        // row count and first row of each representative's class
        int64_t cnt[F_R] = {1, 1, 1, 1};
        dense_update<FK_SUM>(a, (uint32_t)p.n_accs * F_NT, cnt, merge, rep);
        int64_t fr[F_R];
#pragma unroll
        for (int j = 0; j < F_R; ++j) fr[j] = row0 + j * F_NT + tid;
        dense_update<FK_MIN>(a, (uint32_t)(p.n_accs + 1) * F_NT, fr, 0u, rep);  // the representative is the smallest row
      } else if (MODE == FM_BUILD) {
        // ---- join BUILD over the filtered scan (hash_join.rs:148-175 build_hash_table): every qualifying row claims a
        // slot of the HBM-resident linear-probing table (slot = hash tag | row + 1, atomicCAS) and sets its bit of the
        // membership bitmap; an equal key already present raises the duplicate flag (the fused probes need unique keys).
        // No selection vector, no gathered key column: the table holds BASE-table rows and the key column is the base's.
#pragma unroll
        for (int j = 0; j < F_R; ++j) {
          if (!((pass >> j) & 1)) continue;
          const int64_t key = (int64_t)code[j];
          const uint64_t h = fmix64((uint64_t)key);
          const uint32_t tag = (uint32_t)(h >> 32);
          const unsigned long long mine = ((unsigned long long)tag << 32) | (unsigned long long)(row0 + j * F_NT + tid + 1);
          if (p.jt_wbitmap) {
            const uint64_t kk = (uint64_t)key - (uint64_t)p.jt_kmin;
            atomicOr(&p.jt_wbitmap[kk >> 5], 1u << (kk & 31));
          }
          uint64_t sl = h & p.jt_mask;
          while (true) {
            unsigned long long cur = *(volatile unsigned long long*)&p.jt_wslots[sl];
            if (cur == 0) {
              cur = atomicCAS(&p.jt_wslots[sl], 0ull, mine);
              if (cur == 0) break;
            }
            if ((uint32_t)(cur >> 32) == tag) {
              const uint64_t other = (cur & 0xffffffffull) - 1;
              const int64_t ok = p.bkey_width == 8 ? ((const long long*)p.bkey)[other] : (int64_t)((const int*)p.bkey)[other];
              if (ok == key) {
                *p.abort_flag = 1;
                break;
              }
            }
            sl = (sl + 1) & p.jt_mask;
          }
        }
      } else {
        // ---- HBM-resident open-addressing table on the packed key -------------------------------------------
        uint64_t slot[F_R];
#pragma unroll
        for (int j = 0; j < F_R; ++j) {
          slot[j] = F_EMPTY;
          if (!((pass >> j) & 1)) continue;
          const uint64_t c = code[j];
          if (MODE == FM_PROBE || MODE == FM_EMIT) continue;  // resolved below (the first probes of all rows are issued together)
          if (c == F_EMPTY) {
            slot[j] = p.cap_mask + 1;
            continue;
          }
          uint64_t sl = fmix64(c) & p.cap_mask;
          int probes = 0;
          while (true) {
            unsigned long long* keyp = p.g_lo + sl * (uint64_t)p.acc_gstride;
            unsigned long long cur = *(volatile unsigned long long*)keyp;
            if (cur == c) break;
            if (((++probes) & 63) == 0 && *(volatile int*)p.abort_flag) {  // table (nearly) full: host retries larger
              sl = F_EMPTY;
              break;
            }
            if (cur == F_EMPTY) {
              cur = atomicCAS(keyp, F_EMPTY, (unsigned long long)c);
              if (cur == F_EMPTY) {
                const unsigned long long ng = atomicAdd(p.n_groups, 1ull);
                if (2 * (ng + 1) > p.cap_mask + 1) *p.abort_flag = 1;
                break;
              }
              if (cur == c) break;
            }
            sl = (sl + 1) & p.cap_mask;
          }
          slot[j] = sl;
        }
        if (MODE == FM_PROBE || MODE == FM_EMIT) {
          // hash-join probe (hash_join.rs:70-107,177-216): the packed code IS the probe key value.  Most probe rows
          // miss (Q3: ~1% match), so the cost is the latency of the first slot load: issue all F_R of them first.
          uint64_t hh[F_R], sl[F_R];
          unsigned long long cur[F_R];
          if (p.jt_bitmap) {
            uint32_t word[F_R];
            uint64_t kk[F_R];
#pragma unroll
            for (int j = 0; j < F_R; ++j) {
              kk[j] = code[j] - (uint64_t)p.jt_kmin;
              word[j] = (((pass >> j) & 1) && kk[j] <= p.jt_kspan) ? __ldg(&p.jt_bitmap[kk[j] >> 5]) : 0u;
            }
#pragma unroll
            for (int j = 0; j < F_R; ++j)
              if (!((word[j] >> (kk[j] & 31)) & 1u)) pass &= ~(1u << j);
            if (!__any_sync(0xffffffffu, pass != 0)) return;  // no row of this warp has a build partner
          }
#pragma unroll
          for (int j = 0; j < F_R; ++j) {
            hh[j] = fmix64(code[j]);
            sl[j] = hh[j] & p.jt_mask;
            cur[j] = ((pass >> j) & 1) ? __ldg(&p.jt_slots[sl[j]]) : 0ull;
          }
          // the build keys of the first candidates are verified together as well (F_R loads in flight): a row whose first
          // slot settles it -- empty, or tag + key match -- never enters the dependent chain below (load factor <= 0.5)
          int64_t bk0[F_R];
#pragma unroll
          for (int j = 0; j < F_R; ++j) {
            const bool cand = cur[j] != 0 && (uint32_t)(cur[j] >> 32) == (uint32_t)(hh[j] >> 32);
            const uint64_t row = (cur[j] & 0xffffffffull) - 1;
            bk0[j] = !cand ? 0 : (p.bkey_width == 8 ? __ldg((const long long*)p.bkey + row) : (int64_t)__ldg((const int*)p.bkey + row));
            if (!cand && cur[j] != 0) bk0[j] = ~(int64_t)code[j];  // tag mismatch: cannot equal the probe key
          }
#pragma unroll
          for (int j = 0; j < F_R; ++j) {
            const uint32_t tag = (uint32_t)(hh[j] >> 32);
            unsigned long long c = cur[j];
            if (c == 0) continue;
            if (bk0[j] == (int64_t)code[j]) {
              slot[j] = (c & 0xffffffffull) - 1;
              continue;
            }
            uint64_t s2 = (sl[j] + 1) & p.jt_mask;
            c = __ldg(&p.jt_slots[s2]);
            while (c != 0) {
              if ((uint32_t)(c >> 32) == tag) {
                const uint64_t row = (c & 0xffffffffull) - 1;
                const int64_t bk = p.bkey_width == 8 ? __ldg((const long long*)p.bkey + row) : (int64_t)__ldg((const int*)p.bkey + row);
                if (bk == (int64_t)code[j]) {
                  slot[j] = row;
                  break;
                }
              }
              s2 = (s2 + 1) & p.jt_mask;
              c = __ldg(&p.jt_slots[s2]);
            }
          }
        }
        if (MODE == FM_EMIT) {
          // order-free join output: (build row, probe row) pairs appended with one atomic per warp and row slot
          const unsigned lane = threadIdx.x & 31;
          unsigned m[F_R];
          unsigned total = 0;
#pragma unroll
          for (int j = 0; j < F_R; ++j) {
            m[j] = __ballot_sync(0xffffffffu, slot[j] != F_EMPTY);
            total += __popc(m[j]);
          }
          if (total == 0) return;
          unsigned long long base = 0;
          if (lane == 0) base = atomicAdd(p.n_groups, (unsigned long long)total);  // ONE atomic per warp and tile pass
          base = __shfl_sync(0xffffffffu, base, 0);
#pragma unroll
          for (int j = 0; j < F_R; ++j) {
            if (slot[j] != F_EMPTY) {
              const unsigned long long pos = base + __popc(m[j] & ((1u << lane) - 1u));
              p.g_lo[pos] = slot[j];
              p.g_hi[pos] = (unsigned long long)(row0 + j * F_NT + tid);
            }
            base += __popc(m[j]);
          }
          return;
        }
        const size_t kstride = (size_t)p.acc_kstride, gstride = (size_t)p.acc_gstride;
        unsigned long long* const acc0 = p.g_lo + p.acc_base;
        unsigned long long* const acc0_hi = p.g_hi + p.acc_base;
        bool any_slot = false;
#pragma unroll
        for (int j = 0; j < F_R; ++j) any_slot |= slot[j] != F_EMPTY;
        if (MODE == FM_PROBE && !__any_sync(0xffffffffu, any_slot)) return;  // nothing in this warp matched a build row
#pragma unroll 1
        for (int k = 0; k < p.n_accs; ++k) {
          const FAcc& A = p.accs[k];
          const int kind = A.kind;
          int64_t v[F_R];
          if (kind == FK_SUMF || A.unit) {
            load_rows(stage + A.f[0].off, A.f[0].wk, tid, v);
          } else {
#pragma unroll
            for (int j = 0; j < F_R; ++j) v[j] = A.chain ? prev[j] : A.coef;
          }
          if (kind != FK_SUMF) {
#pragma unroll 1
            for (int f = A.unit ? 1 : 0; f < A.n_factors; ++f) {
              int64_t x[F_R];
              load_rows(stage + A.f[f].off, A.f[f].wk, tid, x);
              const int64_t fa = A.f[f].a, fb = A.f[f].b;
#pragma unroll
              for (int j = 0; j < F_R; ++j) v[j] *= (fa + fb * x[j]);
            }
#pragma unroll
            for (int j = 0; j < F_R; ++j) prev[j] = v[j];
          }
#pragma unroll
          for (int j = 0; j < F_R; ++j) {
            if (slot[j] == F_EMPTY) continue;
            unsigned long long* lo = acc0 + k * kstride + slot[j] * gstride;
            if (kind == FK_SUM) {
              if (p.carry) add128_global(lo, acc0_hi + k * kstride + slot[j] * gstride, (unsigned long long)v[j], v[j] < 0 ? ~0ull : 0ull);
              else atomicAdd(lo, (unsigned long long)v[j]);
            } else if (kind == FK_MIN) {
              atomicMin((long long*)lo, (long long)v[j]);
            } else if (kind == FK_MAX) {
              atomicMax((long long*)lo, (long long)v[j]);
            } else {
              atomicAdd((double*)lo, __longlong_as_double(v[j]));
            }
          }
        }
#pragma unroll
        for (int j = 0; j < F_R; ++j) {
          if (slot[j] == F_EMPTY) continue;
          atomicAdd(acc0 + (size_t)p.n_accs * kstride + slot[j] * gstride, 1ull);
          atomicMin((long long*)(acc0 + (size_t)(p.n_accs + 1) * kstride + slot[j] * gstride), (long long)(row0 + j * F_NT + tid));
        }
      }
    }
  }
};

#endif  // !QGPU_JIT

// ------------------------------------------------------------------------------------------------
// kernel skeleton: TMA/mbarrier tile pipeline + per-CTA reduction of the private tables; `Body::tile`
// consumes one staged tile
// ------------------------------------------------------------------------------------------------
template <int MODE, class Body>
__device__ __forceinline__ void fused_main(const FParams& p) {
  // Warp-specialised TMA pipeline: warps 0..7 (F_NT threads) consume tiles, warp 8 is the producer (one elected
  // lane issues the bulk copies).  full[s]: producer -> consumers (complete_tx bytes); empty[s]: consumers ->
  // producer (one arrive per consumer warp).  No CTA-wide barrier per tile: a fast warp runs up to `stages`
  // tiles ahead of a slow one.
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = (uint64_t*)smem;
  uint64_t* empty = full + 8;
  unsigned char* tiles = smem + 128;
  long long* priv = (long long*)(smem + p.priv_off);
  const int tid = threadIdx.x;
  const int NA2 = p.n_accs + 2;
  const int NL = MODE == FM_DENSE ? p.priv_lines : NA2;  // private lines per slot (packed accumulators have none)
  const bool producer = tid >= F_NT;

  if (MODE == FM_DENSE && !producer) {
    const int total = p.dense_groups * NL * F_NT;
    for (int i = tid; i < total; i += F_NT) priv[i] = acc_init(p, p.priv_k_of[(i / F_NT) % NL]);
  }
  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], F_NT / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  const int64_t first_tile = blockIdx.x;
  const int64_t my_tiles = first_tile < p.n_tiles ? (p.n_tiles - first_tile + gridDim.x - 1) / gridDim.x : 0;

  if (producer) {
    // Lane c of the producer warp issues column c's bulk copy: the copies of one tile are issued concurrently.  (One
    // elected thread issuing all of them back to back capped the whole pipeline at ~5.2 TB/s -- Q1 SF10 0.48 ms; one
    // lane per column reaches the box's read bandwidth, 6.7-6.9 TB/s -- 0.34 ms.)
    const int lane = tid - F_NT;
    const uint64_t l2_policy = l2_evict_first_policy();
    const unsigned char* src = lane < p.n_cols ? p.cols[lane].ptr : nullptr;
    const uint32_t width = lane < p.n_cols ? p.cols[lane].width : 0u;
    const uint32_t smem_off = lane < p.n_cols ? p.cols[lane].smem_off : 0u;
#pragma unroll 1
    for (int64_t i = 0; i < my_tiles; ++i) {
      const int s = (int)(i % p.stages);
      const int64_t use = i / p.stages;
      if (use > 0) mbar_wait(&empty[s], (uint32_t)((use - 1) & 1));  // every consumer warp released the previous use
      const int64_t row0 = (first_tile + i * gridDim.x) * F_T;
      const int64_t rows = min((int64_t)F_T, p.n_rows - row0);
      unsigned char* dst = tiles + (size_t)s * p.stage_bytes;
      const uint32_t bytes = ((uint32_t)(rows * width) + 15u) & ~15u;
      uint32_t total = bytes;
#pragma unroll
      for (int d = 16; d; d >>= 1) total += __shfl_xor_sync(0xffffffffu, total, d);
      if (lane == 0) mbar_expect_tx(&full[s], total);
      __syncwarp();  // the expected byte count is registered before any copy can complete
      if (lane < p.n_cols) bulk_g2s(dst + smem_off, src + (size_t)row0 * width, bytes, &full[s], l2_policy);
    }
  } else {
#pragma unroll 1
    for (int64_t i = 0; i < my_tiles; ++i) {
      const int s = (int)(i % p.stages);
      mbar_wait(&full[s], (uint32_t)((i / p.stages) & 1));
      const unsigned char* stage = tiles + (size_t)s * p.stage_bytes;
      const int64_t row0 = (first_tile + i * gridDim.x) * F_T;
      const int rows = (int)min((int64_t)F_T, p.n_rows - row0);
      // HASH: once the table overflowed the host retries with a larger one; the rest of this pass only drains
      if (!(MODE == FM_HASH && *(volatile int*)p.abort_flag != 0)) Body::tile(p, stage, row0, rows, tid, priv, NL);
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(&empty[s]);  // this warp is done with stage s
    }
  }

  // ---- DENSE: reduce the per-thread private tables once per CTA into the 128-bit global table ----
  if (MODE == FM_DENSE) {
    __syncthreads();
    const int lane = tid & 31, warp = tid >> 5;
    const int n_lines = p.dense_groups * NA2;
    for (int line = producer ? n_lines : warp; line < n_lines; line += F_NT / 32) {
      const int k = line % NA2;
      const int g = line / NA2;
      const int pl = p.priv_line_of[k];  // 0xff: packed into the count word
      const long long* src = priv + ((size_t)g * NL + (pl == 0xff ? 0 : pl)) * F_NT;
      int kind = FK_SUM;
      if (k < p.n_accs) kind = p.accs[k].kind;
      else if (k == p.n_accs + 1) kind = FK_MIN;
      // skip groups this CTA never saw
      long long cnt = 0;
      const long long* csrc = priv + ((size_t)g * NL + p.priv_line_of[p.n_accs]) * F_NT;
      const unsigned long long cmask = p.pack_mask ? ((1ull << p.pack_cnt_bits) - 1ull) : ~0ull;
      {
        for (int j = lane; j < F_NT; j += 32) cnt += (long long)((unsigned long long)csrc[j] & cmask);
#pragma unroll
        for (int d = 16; d; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
      }
      if (cnt == 0) continue;
      unsigned long long* glo = p.g_lo + line;
      unsigned long long* ghi = p.g_hi + line;
      if (kind == FK_SUM) {
        // packed accumulators and the row count are fields of the count word (see FParams::pack_mask)
        const bool packed = k < p.n_accs && ((p.pack_mask >> k) & 1u);
        const long long* rsrc = packed ? csrc : src;
        const int sh = packed ? (int)p.pack_shift[k] : 0;
        const unsigned long long fm = packed ? ((1ull << p.pack_bits[k]) - 1ull) : (k == p.n_accs ? cmask : ~0ull);
        i128 s = 0;
        for (int j = lane; j < F_NT; j += 32) s += (i128)(long long)(((unsigned long long)rsrc[j] >> sh) & fm);
        unsigned long long lo = (unsigned long long)(u128)s, hi = (unsigned long long)((u128)s >> 64);
#pragma unroll
        for (int d = 16; d; d >>= 1) {
          const unsigned long long olo = __shfl_xor_sync(0xffffffffu, lo, d), ohi = __shfl_xor_sync(0xffffffffu, hi, d);
          const u128 t = (((u128)hi << 64) | lo) + (((u128)ohi << 64) | olo);
          lo = (unsigned long long)t;
          hi = (unsigned long long)(t >> 64);
        }
        if (lane == 0) add128_global(glo, ghi, lo, hi);
      } else if (kind == FK_SUMF) {
        double s = 0;
        for (int j = lane; j < F_NT; j += 32) s += __longlong_as_double(src[j]);
#pragma unroll
        for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
        if (lane == 0) atomicAdd((double*)glo, s);
      } else {
        long long m = kind == FK_MIN ? INT64_MAX : INT64_MIN;
        for (int j = lane; j < F_NT; j += 32) m = kind == FK_MIN ? min(m, src[j]) : max(m, src[j]);
#pragma unroll
        for (int d = 16; d; d >>= 1) {
          const long long o = __shfl_xor_sync(0xffffffffu, m, d);
          m = kind == FK_MIN ? min(m, o) : max(m, o);
        }
        if (lane == 0) {
          if (kind == FK_MIN) atomicMin((long long*)glo, m);
          else atomicMax((long long*)glo, m);
        }
      }
    }
  }
}


// ------------------------------------------------------------------------------------------------
// tile body #2: compile-time specialised on the plan's SHAPE SIGNATURE (operand widths, accumulator
// kinds / chaining / factor counts); constants (bounds, coefficients, tile offsets) stay runtime
// parameters.  The host computes the signature of every fused plan and launches the matching
// instantiation when one is registered (the shapes of TPC-H Q1 and Q6 are); everything else runs the
// generic body above.  Registering a shape = one line in SPEC_SHAPES below.
//
// signature: s[0] = n_pred(4) | 8 x pred width code(3) | n_keys(3) | 4 x key width code(3) | n_accs(4)
//            s[1..3]: 3 accumulators each, 18 bits: kind(2) chain(1) unit(1) n_factors(2) 3 x {width code(3) plain(1)}
// width codes: 0 i8, 1 u8, 2 i16, 3 u16, 4 i32, 5 u32, 6 i64
// ------------------------------------------------------------------------------------------------
struct FSig {
  uint64_t s[4];
};
__host__ __device__ constexpr uint32_t wk_to_code(uint32_t wk) {
  return wk == 1 ? 0u : wk == (1u | 256u) ? 1u : wk == 2 ? 2u : wk == (2u | 256u) ? 3u : wk == 4 ? 4u : wk == (4u | 256u) ? 5u : 6u;
}
enum : uint32_t { WC_I8 = 0, WC_U8 = 1, WC_I16 = 2, WC_U16 = 3, WC_I32 = 4, WC_U32 = 5, WC_I64 = 6 };
__host__ __device__ constexpr uint64_t sig_factor(uint32_t wc, bool plain) { return (uint64_t)wc | ((uint64_t)(plain ? 1 : 0) << 3); }
__host__ __device__ constexpr uint64_t sig_acc(int kind, bool chain, bool unit, int nf, uint64_t f0 = 0, uint64_t f1 = 0, uint64_t f2 = 0) {
  return (uint64_t)kind | ((uint64_t)(chain ? 1 : 0) << 2) | ((uint64_t)(unit ? 1 : 0) << 3) | ((uint64_t)nf << 4) | (f0 << 6) | (f1 << 10) |
         (f2 << 14);
}
__host__ __device__ constexpr uint64_t sig_head(int n_pred, uint64_t preds, int n_keys, uint64_t keys, int n_accs) {
  return (uint64_t)n_pred | (preds << 4) | ((uint64_t)n_keys << 28) | (keys << 31) | ((uint64_t)n_accs << 43);
}
__host__ __device__ constexpr uint64_t sig_list(uint32_t a = 0, uint32_t b = 0, uint32_t c = 0, uint32_t d = 0, uint32_t e = 0,
                                                uint32_t f = 0, uint32_t g = 0, uint32_t h = 0) {
  return (uint64_t)a | ((uint64_t)b << 3) | ((uint64_t)c << 6) | ((uint64_t)d << 9) | ((uint64_t)e << 12) | ((uint64_t)f << 15) |
         ((uint64_t)g << 18) | ((uint64_t)h << 21);
}
__host__ __device__ constexpr uint64_t sig_accs3(uint64_t a0 = 0, uint64_t a1 = 0, uint64_t a2 = 0) { return a0 | (a1 << 18) | (a2 << 36); }

template <uint64_t S0, uint64_t S1, uint64_t S2, uint64_t S3>
struct SigView {
  static constexpr int n_pred = (int)(S0 & 15);
  static constexpr int n_keys = (int)((S0 >> 28) & 7);
  static constexpr int n_accs = (int)((S0 >> 43) & 15);
  static constexpr uint32_t pred_wc(int k) { return (uint32_t)((S0 >> (4 + 3 * k)) & 7); }
  static constexpr uint32_t key_wc(int k) { return (uint32_t)((S0 >> (31 + 3 * k)) & 7); }
  static constexpr uint64_t acc(int k) { return ((k < 3 ? S1 : (k < 6 ? S2 : S3)) >> (18 * (k % 3))) & 0x3ffff; }
  static constexpr int kind(int k) { return (int)(acc(k) & 3); }
  static constexpr bool chain(int k) { return (acc(k) >> 2) & 1; }
  static constexpr bool unit(int k) { return (acc(k) >> 3) & 1; }
  static constexpr int n_factors(int k) { return (int)((acc(k) >> 4) & 3); }
  static constexpr uint32_t f_wc(int k, int f) { return (uint32_t)((acc(k) >> (6 + 4 * f)) & 7); }
  static constexpr bool f_plain(int k, int f) { return (acc(k) >> (6 + 4 * f + 3)) & 1; }
};

template <uint32_t WC>
__device__ __forceinline__ void load_rows_t(const unsigned char* col, int tid, int64_t (&x)[F_R]) {
#pragma unroll
  for (int j = 0; j < F_R; ++j) {
    if (WC == WC_I64) x[j] = ((const int64_t*)col)[j * F_NT + tid];
    else if (WC == WC_I32) x[j] = (int64_t)((const int32_t*)col)[j * F_NT + tid];
    else if (WC == WC_U32) x[j] = (int64_t)((const uint32_t*)col)[j * F_NT + tid];
    else if (WC == WC_I16) x[j] = (int64_t)((const int16_t*)col)[j * F_NT + tid];
    else if (WC == WC_U16) x[j] = (int64_t)((const uint16_t*)col)[j * F_NT + tid];
    else if (WC == WC_I8) x[j] = (int64_t)((const int8_t*)col)[j * F_NT + tid];
    else x[j] = (int64_t)((const uint8_t*)col)[j * F_NT + tid];
  }
}

template <uint64_t S0, uint64_t S1, uint64_t S2, uint64_t S3, uint32_t PACK>
struct SpecBody {
  typedef SigView<S0, S1, S2, S3> G;
  static constexpr bool packed(int k) { return (PACK >> k) & 1u; }
  static constexpr int popc(uint32_t x) { return x == 0 ? 0 : (int)(x & 1u) + popc(x >> 1); }
  static constexpr int line_of(int k) { return k - popc(PACK & ((1u << k) - 1u)); }  // FParams::priv_line_of
  static constexpr int n_lines = G::n_accs + 2 - popc(PACK);
  // value of accumulator K for the thread's rows (compile-time recursion keeps every index static)
  template <int K, int F>
  static __device__ __forceinline__ void factors(const FParams& p, const unsigned char* stage, int tid, int64_t (&v)[F_R]) {
    if constexpr (F < G::n_factors(K)) {
      int64_t x[F_R];
      load_rows_t<G::f_wc(K, F)>(stage + p.accs[K].f[F].off, tid, x);
      if constexpr (G::f_plain(K, F)) {
#pragma unroll
        for (int j = 0; j < F_R; ++j) v[j] *= x[j];
      } else {
        const int64_t fa = p.accs[K].f[F].a, fb = p.accs[K].f[F].b;
#pragma unroll
        for (int j = 0; j < F_R; ++j) v[j] *= (fa + fb * x[j]);
      }
      factors<K, F + 1>(p, stage, tid, v);
    }
  }
  template <int K>
  static __device__ __forceinline__ void values(const FParams& p, const unsigned char* stage, int tid, int64_t (&val)[F_MAXA][F_R]) {
    if constexpr (K < G::n_accs) {
      if constexpr (G::kind(K) == FK_SUMF || G::unit(K)) {
        load_rows_t<G::f_wc(K, 0)>(stage + p.accs[K].f[0].off, tid, val[K]);
        if constexpr (G::kind(K) != FK_SUMF) factors<K, 1>(p, stage, tid, val[K]);
      } else {
        if constexpr (G::chain(K)) {
#pragma unroll
          for (int j = 0; j < F_R; ++j) val[K][j] = val[K - 1][j];
        } else {
          const int64_t c = p.accs[K].coef;
#pragma unroll
          for (int j = 0; j < F_R; ++j) val[K][j] = c;
        }
        factors<K, 0>(p, stage, tid, val[K]);
      }
      values<K + 1>(p, stage, tid, val);
    }
  }
  template <int K>
  static __device__ __forceinline__ void preds(const FParams& p, const unsigned char* stage, int tid, uint32_t& pass) {
    if constexpr (K < G::n_pred) {
      int64_t x[F_R];
      load_rows_t<G::pred_wc(K)>(stage + p.pred[K].off, tid, x);
      const uint64_t lo = (uint64_t)p.pred[K].lo, span = p.pred[K].span;
#pragma unroll
      for (int j = 0; j < F_R; ++j)
        if (((uint64_t)x[j] - lo) > span) pass &= ~(1u << j);
      preds<K + 1>(p, stage, tid, pass);
    }
  }
  template <int K>
  static __device__ __forceinline__ void keys(const FParams& p, const unsigned char* stage, int tid, uint32_t (&code)[F_R]) {
    if constexpr (K < G::n_keys) {
      int64_t x[F_R];
      load_rows_t<G::key_wc(K)>(stage + p.keys[K].off, tid, x);
      const uint32_t base = (uint32_t)p.keys[K].base, mult = (uint32_t)p.keys[K].mult;
#pragma unroll
      for (int j = 0; j < F_R; ++j) code[j] += ((uint32_t)x[j] - base) * mult;  // dense index < 4096: 32-bit is exact
      keys<K + 1>(p, stage, tid, code);
    }
  }
  template <int K>
  static __device__ __forceinline__ void rmw_load(const long long* a, long long (&cur)[F_MAXA]) {
    if constexpr (K < G::n_accs) {
      if constexpr (!packed(K)) cur[K] = a[line_of(K) * F_NT];
      rmw_load<K + 1>(a, cur);
    }
  }
  template <int K>
  static __device__ __forceinline__ void rmw_store(long long* a, const long long (&cur)[F_MAXA], const int64_t (&val)[F_MAXA][F_R], int j) {
    if constexpr (K < G::n_accs) {
      if constexpr (!packed(K)) a[line_of(K) * F_NT] = comb<G::kind(K)>(cur[K], val[K][j]);
      rmw_store<K + 1>(a, cur, val, j);
    }
  }
  // increment of the count word: 1 row + the packed accumulators' values in their bit fields
  template <int K>
  static __device__ __forceinline__ void pack_inc(const FParams& p, const int64_t (&val)[F_MAXA][F_R], int j, long long& inc) {
    if constexpr (K < G::n_accs) {
      if constexpr (packed(K)) inc += (long long)((unsigned long long)val[K][j] << p.pack_shift[K]);
      pack_inc<K + 1>(p, val, j, inc);
    }
  }

  static __device__ __forceinline__ void tile(const FParams& p, const unsigned char* stage, const int64_t row0, const int rows,
                                              const int tid, long long* priv, const int /*NA2*/) {
    constexpr int CL = n_lines - 2, FL = n_lines - 1;  // count word, first row
    uint32_t pass = 0;
#pragma unroll
    for (int j = 0; j < F_R; ++j)
      if (j * F_NT + tid < rows) pass |= 1u << j;
    preds<0>(p, stage, tid, pass);
    if (!__any_sync(0xffffffffu, pass != 0)) return;
    uint32_t code[F_R] = {0, 0, 0, 0};
    keys<0>(p, stage, tid, code);
    int64_t val[F_MAXA][F_R];
    values<0>(p, stage, tid, val);
    // Row-sequential read-modify-write of the thread's private slots: the K loads of one row are issued
    // together (distinct slots), rows follow in program order (two rows may share a group).
#pragma unroll
    for (int j = 0; j < F_R; ++j) {
      if ((pass >> j) & 1) {
        long long* a = priv + code[j] * (uint32_t)(n_lines * F_NT) + (uint32_t)tid;
        long long cur[F_MAXA];
        rmw_load<0>(a, cur);
        const long long cc = a[CL * F_NT];
        rmw_store<0>(a, cur, val, j);
        long long inc = 1;
        pack_inc<0>(p, val, j, inc);
        a[CL * F_NT] = cc + inc;
        // a thread meets its rows in ascending order (static round-robin tiles): the first row of a private slot is
        // the row that finds its count word still zero
        if (cc == 0) a[FL * F_NT] = (long long)(row0 + j * F_NT + tid);
      }
    }
  }
};

}  // namespace qgpu
