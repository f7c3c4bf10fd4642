// k_dense_epilogue: see epilogue.h.  One CTA of 1024 threads.
//   A  compaction of the occupied slots of the dense accumulator table (slot order)
//   B  this rank's state block: one record per group = key VALUES, first row, row count, accumulator words
//   X  (world > 1) peer stores of the block into every rank's symmetric buffer + epoch-flag barrier
//   M  (world > 1) exact merge of the gathered records (leader search by key words, fold in rank order)
//   F  first-occurrence ranking, finalisation of every aggregate, key columns, metadata
//   R  re-initialisation of the accumulator table for the next execution
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "comm.h"
#include "epilogue.h"
#include "launch.h"

namespace qgpu {

namespace {

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// shared-memory layout for `cap` records / groups (dynamic: the small cases -- Q1: 4 groups x 8 ranks -- leave room for a
// scan CTA of the NEXT execution on the same SM, so the epilogue overlaps it; see run_dense in fused.cu)
struct EpiSmem {
  long long* first;        // merged group -> first row                                   [cap]
  unsigned int* pos;       // gathered record -> word offset in this rank's buffer        [cap]
  int* pre;                // compaction prefix                                           [EPI_NT]
  unsigned short* slot;    // local group -> dense slot                                   [cap]
  unsigned short* leader;  // gathered record -> first record with the same key           [cap]
  unsigned short* gid;     // leader record -> merged group                               [cap]
  unsigned short* order;   // output position -> merged group                             [cap]
  unsigned long long* nulls;      // [EPI_MAXAGG]
  unsigned long long* key_nulls;  // [EPI_MAXK]
  int* base;               // [COMM_MAX_WORLD + 1]
  int* scalars;            // n_local, n_rec, n_groups, eval_err, x_err
  __device__ EpiSmem(unsigned char* raw, int cap) {
    first = (long long*)raw;
    nulls = (unsigned long long*)(first + cap);
    key_nulls = nulls + EPI_MAXAGG;
    pos = (unsigned int*)(key_nulls + EPI_MAXK);
    pre = (int*)(pos + cap);
    base = pre + EPI_NT;
    scalars = base + COMM_MAX_WORLD + 1;
    slot = (unsigned short*)(scalars + 7);
    leader = slot + cap;
    gid = leader + cap;
    order = gid + cap;
  }
};
}  // namespace

__global__ void __launch_bounds__(EPI_NT) k_dense_epilogue(const __grid_constant__ EpiParams p) {
  extern __shared__ __align__(16) unsigned char epi_smem_raw[];
  EpiSmem s(epi_smem_raw, p.smem_cap);
  int& s_n_local = s.scalars[0];
  int& s_n_rec = s.scalars[1];
  int& s_n_groups = s.scalars[2];
  int& s_eval_err = s.scalars[3];
  int& s_x_err = s.scalars[4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int NA2 = p.n_accs + 2;
  const int nk = p.n_keys, rw = p.rw;
  const int off_first = 3 * nk, off_agg = 3 * nk + 1;  // record: keys | first row | n_aggs x {lo, hi, count}

  if (tid == 0) {
    s_eval_err = 0;
    s_x_err = 0;
  }
  if (tid < EPI_MAXAGG) s.nulls[tid] = 0;
  if (tid < EPI_MAXK) s.key_nulls[tid] = 0;

  // ---- A: occupied slots ---------------------------------------------------------------------------------------
  if (p.src == EPI_SRC_PACKED) {
    if (tid == 0) s_n_local = (int)p.rec[1] + (p.rec[4] ? p.max_groups + 1 : 0);  // k_pack_records filled the block
    __syncthreads();
  } else if (!p.grouped) {
    if (tid == 0) {
      s.slot[0] = 0;  // the single group is exported even when no row qualified (NULL sums, MIN/MAX start values)
      s_n_local = 1;
    }
    __syncthreads();
  } else {
    const int per = (p.n_slots + EPI_NT - 1) / EPI_NT;
    const int b0 = tid * per;
    int c = 0;
    unsigned mask = 0;
    for (int i = 0; i < per; ++i) {
      const int g = b0 + i;
      const bool occ = g < p.n_slots && p.g_lo[(size_t)g * NA2 + p.n_accs] != 0;
      c += occ;
      mask |= (unsigned)occ << i;
    }
    s.pre[tid] = c;
    __syncthreads();
    for (int d = 1; d < EPI_NT; d <<= 1) {
      const int v = tid >= d ? s.pre[tid - d] : 0;
      __syncthreads();
      s.pre[tid] += v;
      __syncthreads();
    }
    int o = s.pre[tid] - c;
    for (int i = 0; i < per; ++i)
      if ((mask >> i) & 1) s.slot[o++] = (unsigned short)(b0 + i);
    if (tid == EPI_NT - 1) s_n_local = s.pre[EPI_NT - 1];
    __syncthreads();
  }
  const int n_local = s_n_local;
  const int n_send = min(n_local, p.max_groups);

  // ---- B: state block --------------------------------------------------------------------------------------------
  if (tid == 0) {
    p.rec[0] = EPI_MAGIC;
    p.rec[1] = (unsigned long long)n_send;
    p.rec[2] = (unsigned long long)nk;
    p.rec[3] = (unsigned long long)p.n_aggs;
    p.rec[4] = n_local > p.max_groups ? 1ull : 0ull;
    p.rec[5] = (unsigned long long)rw;
    p.rec[6] = p.epoch;
    p.rec[7] = 0;
  }
  for (int gi = tid; gi < n_send && p.src == EPI_SRC_DENSE; gi += EPI_NT) {
    const unsigned slot = s.slot[gi];
    unsigned long long* r = p.rec + EPI_HDR + (size_t)gi * rw;
    for (int k = 0; k < nk; ++k) {
      const EpiKey& K = p.key[k];
      const unsigned digit = (slot / K.mult) % K.range;
      unsigned long long tag = 1, w1 = 0, w2 = 0;
      if (K.is_dict) {
        const int o0 = K.dict_offs[digit], len = K.dict_offs[digit + 1] - o0;
        for (int i = 0; i < len && i < 16; ++i) {
          const unsigned long long b = (unsigned char)K.dict_data[o0 + i];
          if (i < 8) w1 |= b << (8 * i);
          else w2 |= b << (8 * (i - 8));
        }
        tag = 2 + (unsigned long long)len;
      } else {
        const long long v = K.base + (long long)digit;
        w1 = (unsigned long long)v;
        w2 = v < 0 ? ~0ull : 0ull;
      }
      r[3 * k] = tag;
      r[3 * k + 1] = w1;
      r[3 * k + 2] = w2;
    }
    const unsigned long long* t = p.g_lo + (size_t)slot * NA2;
    const unsigned long long rows = t[p.n_accs];
    r[off_first] = rows ? (unsigned long long)((long long)t[p.n_accs + 1] + p.row_offset) : (unsigned long long)INT64_MAX;
    for (int a = 0; a < p.n_aggs; ++a) {
      const int k = p.agg_acc[a];
      unsigned long long lo = 0, hi = 0;
      if (k >= 0) {
        lo = t[k];
        if (p.acc_wide[k]) hi = p.g_hi[(size_t)slot * NA2 + k];
        else if (p.acc_kind[k] != 3) hi = ((long long)lo < 0) ? ~0ull : 0ull;
        if (rows == 0 && (p.acc_kind[k] == 1 || p.acc_kind[k] == 2)) {
          // ungrouped aggregate, no qualifying row on this shard: the TYPED start value (min.rs / max.rs NATIVE::MAX / MIN,
          // quirk Q4) instead of the scan kernel's 64-bit working seed
          lo = p.sent_lo[a];
          hi = p.sent_hi[a];
        }
      }
      unsigned long long* q = r + off_agg + 3 * a;
      q[0] = lo;
      q[1] = hi;
      q[2] = rows;  // the fused path only takes provably NULL-free arguments: every aggregate saw every row
    }
  }
  __syncthreads();

  const unsigned long long* R = p.rec + EPI_HDR;  // records the finalisation reads (world == 1: this rank's)
  int G = n_send;
  int x_err = (n_local > p.max_groups) ? EPI_ERR_OVERFLOW : 0;

  if (p.world > 1) {
    // ---- X: every rank's block into every rank's buffer, then the epoch flags ------------------------------------
    const size_t slot_words = COMM_SLOT_BYTES / 8, flag_words = COMM_FLAG_BYTES / 8;
    const unsigned parity = (unsigned)(p.epoch & 1);
    const int words = EPI_HDR + n_send * rw;
    for (int i = tid; i < words * p.world; i += EPI_NT) {
      const int q = i / words, w = i - q * words;
      p.peer[q][flag_words + ((size_t)parity * p.world + p.rank) * slot_words + w] = p.rec[w];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < p.world) st_release_sys(p.peer[tid] + parity * COMM_MAX_WORLD + p.rank, p.epoch);
    if (tid < p.world) {
      const unsigned long long* f = p.peer[p.rank] + parity * COMM_MAX_WORLD + tid;
      const unsigned long long t0 = global_ns();
      unsigned spins = 0;
      while (ld_acquire_sys(f) != p.epoch) {
        __nanosleep(64);
        if (((++spins) & 1023u) == 0 && global_ns() - t0 > p.timeout_ns) {
          atomicMax(&s_x_err, (int)EPI_ERR_TIMEOUT);
          break;
        }
      }
    }
    __syncthreads();
    // ---- M: merge ------------------------------------------------------------------------------------------------
    const volatile unsigned long long* in = p.peer[p.rank] + flag_words + (size_t)parity * p.world * slot_words;
    if (tid == 0) {
      int t = 0, bad = s_x_err;
      for (int r = 0; r < p.world && !bad; ++r) {
        const volatile unsigned long long* h = in + (size_t)r * slot_words;
        s.base[r] = t;
        if (h[0] != EPI_MAGIC || (int)h[2] != nk || (int)h[3] != p.n_aggs || (int)h[5] != rw || h[6] != p.epoch) bad = EPI_ERR_BLOCK;
        else if (h[4]) bad = EPI_ERR_OVERFLOW;
        else t += (int)h[1];
      }
      if (!bad && t > p.smem_cap) bad = EPI_ERR_TOO_MANY;
      s.base[p.world] = bad ? 0 : t;
      s_n_rec = bad ? 0 : t;
      s_x_err = bad;
    }
    __syncthreads();
    const int M = s_n_rec;
    x_err = s_x_err;
    for (int r = 0; r < p.world && M > 0; ++r) {
      const int n = s.base[r + 1] - s.base[r];
      for (int g = tid; g < n; g += EPI_NT) s.pos[s.base[r] + g] = (unsigned int)((size_t)r * slot_words + EPI_HDR + (size_t)g * rw);
    }
    __syncthreads();
    for (int i = tid; i < M; i += EPI_NT) {
      const volatile unsigned long long* ri = in + s.pos[i];
      int l = i;
      for (int j = 0; j < i; ++j) {
        const volatile unsigned long long* rj = in + s.pos[j];
        bool eq = true;
        for (int w = 0; w < 3 * nk && eq; ++w) eq = ri[w] == rj[w];
        if (eq) {
          l = j;
          break;
        }
      }
      s.leader[i] = (unsigned short)l;
    }
    __syncthreads();
    if (tid == 0) {
      int ng = 0;
      for (int i = 0; i < M; ++i)
        if (s.leader[i] == i) s.gid[i] = (unsigned short)ng++;
      if (ng > p.g_max) {
        s_x_err = EPI_ERR_TOO_MANY;
        ng = 0;
      }
      s_n_groups = ng;
    }
    __syncthreads();
    G = s_n_groups;
    x_err = s_x_err;
    // one thread per merged group folds that group's records in (rank, slot) order: deterministic
    for (int i = tid; i < M && G > 0; i += EPI_NT) {
      if (s.leader[i] != i) continue;
      unsigned long long* out = p.mrec + (size_t)s.gid[i] * rw;
      const volatile unsigned long long* r0 = in + s.pos[i];
      for (int w = 0; w < 3 * nk; ++w) out[w] = r0[w];
      long long first = INT64_MAX;
      for (int j = i; j < M; ++j)
        if (s.leader[j] == i) first = min(first, (long long)in[s.pos[j] + off_first]);
      out[off_first] = (unsigned long long)first;
      for (int a = 0; a < p.n_aggs; ++a) {
        const int kind = p.agg[a].kind;
        unsigned long long lo = 0, hi = 0, cnt = 0;
        bool have = false;
        for (int j = i; j < M; ++j) {
          if (s.leader[j] != i) continue;
          const volatile unsigned long long* q = in + s.pos[j] + off_agg + 3 * a;
          const unsigned long long l2 = q[0], h2 = q[1];
          cnt += q[2];
          if (!have) {
            lo = l2;
            hi = h2;
            have = true;
            continue;
          }
          switch (kind) {
            case AK_COUNT: break;
            case AK_SUM_I64: lo += l2; break;
            case AK_SUM_DEC: {
              const u128 t = (((u128)hi << 64) | lo) + (((u128)h2 << 64) | l2);
              lo = (unsigned long long)t;
              hi = (unsigned long long)(t >> 64);
              break;
            }
            case AK_SUM_F64: lo = (unsigned long long)__double_as_longlong(__longlong_as_double((long long)lo) + __longlong_as_double((long long)l2)); break;
            case AK_MIN_I64: case AK_MIN_F64: if ((long long)l2 < (long long)lo) { lo = l2; hi = h2; } break;  // f64 travels as its total-order key
            case AK_MAX_I64: case AK_MAX_F64: if ((long long)l2 > (long long)lo) { lo = l2; hi = h2; } break;
            case AK_MIN_U64: if (l2 < lo) lo = l2; break;
            case AK_MAX_U64: if (l2 > lo) lo = l2; break;
            case AK_MIN_DEC: if ((long long)h2 < (long long)hi || (h2 == hi && l2 < lo)) { lo = l2; hi = h2; } break;
            case AK_MAX_DEC: if ((long long)hi < (long long)h2 || (h2 == hi && lo < l2)) { lo = l2; hi = h2; } break;
            default: break;
          }
        }
        unsigned long long* q = out + off_agg + 3 * a;
        q[0] = lo;
        q[1] = hi;
        q[2] = cnt;
      }
    }
    __syncthreads();
    R = p.mrec;
  }
  if (x_err) G = 0;

  // ---- F: order = first occurrence (the reference's order is unspecified) -------------------------
  for (int t = tid; t < G; t += EPI_NT) s.first[t] = (long long)R[(size_t)t * rw + off_first];
  __syncthreads();
  for (int t = tid; t < G; t += EPI_NT) {
    const long long mine = s.first[t];
    int rank = 0;
    for (int h = 0; h < G; ++h) rank += (s.first[h] < mine) || (s.first[h] == mine && h < t);
    s.order[rank] = (unsigned short)t;
  }
  __syncthreads();
  // aggregates: one warp per (aggregate, 32 output rows): the validity word is one ballot
  const int n_words = (G + 31) >> 5;
  for (int item = warp; item < p.n_aggs * n_words; item += EPI_NT / 32) {
    const int a = item / n_words, w = item - a * n_words;
    FinSpec f = p.agg[a];
    f.lo = R + off_agg + 3 * a;
    f.hi = f.lo + 1;
    f.cnt = f.lo + 2;
    f.stride = rw;
    const int pos = (w << 5) + lane;
    bool valid = false;
    unsigned long long lo = 0, hi = 0;
    if (pos < G) {
      const int t = s.order[pos];
      valid = fin_value(f, t, &s_eval_err, &lo, &hi);
    }
    const uint32_t vw = __ballot_sync(0xffffffffu, valid);
    if (lane == 0) {
      f.out_valid[w] = vw;
      const int live = min(32, G - (w << 5));
      if (live - __popc(vw)) atomicAdd(&s.nulls[a], (unsigned long long)(live - __popc(vw)));
    }
    if (pos < G) fin_store(f, pos, lo, hi);
  }
  // key columns: the key VALUES travel in the records
  for (int k = 0; k < nk; ++k) {
    const EpiKey& K = p.key[k];
    if (K.out_valid) {
      for (int w = warp; w < n_words; w += EPI_NT / 32) {
        const int pos = (w << 5) + lane;
        const bool valid = pos < G && R[(size_t)s.order[pos] * rw + 3 * k] != 0;
        const uint32_t vw = __ballot_sync(0xffffffffu, valid);
        if (lane == 0) {
          K.out_valid[w] = vw;
          const int live = min(32, G - (w << 5));
          if (live - __popc(vw)) atomicAdd(&s.key_nulls[k], (unsigned long long)(live - __popc(vw)));
        }
      }
    }
    if (K.out_phys == PH_STR) {
      if (tid == 0) {
        int off = 0;
        for (int pos = 0; pos < G; ++pos) {
          K.out_offsets[pos] = off;
          const unsigned long long tag = R[(size_t)s.order[pos] * rw + 3 * k];
          off += tag >= 2 ? (int)(tag - 2) : 0;
        }
        K.out_offsets[G] = off;
      }
      __syncthreads();
      for (int pos = tid; pos < G; pos += EPI_NT) {
        const unsigned long long* r = R + (size_t)s.order[pos] * rw + 3 * k;
        const int len = r[0] >= 2 ? (int)(r[0] - 2) : 0;
        char* dst = (char*)K.out + K.out_offsets[pos];
        for (int b = 0; b < len; ++b) dst[b] = (char)(((b < 8 ? r[1] : r[2]) >> (8 * (b & 7))) & 0xff);
      }
    } else {
      for (int pos = tid; pos < G; pos += EPI_NT) {
        const unsigned long long* r = R + (size_t)s.order[pos] * rw + 3 * k;
        const unsigned long long w1 = r[1], w2 = r[2];
        switch (K.out_phys) {
          case PH_I8: case PH_U8: ((uint8_t*)K.out)[pos] = (uint8_t)w1; break;
          case PH_I16: case PH_U16: ((uint16_t*)K.out)[pos] = (uint16_t)w1; break;
          case PH_I32: case PH_U32: ((uint32_t*)K.out)[pos] = (uint32_t)w1; break;
          case PH_I64: case PH_U64: ((unsigned long long*)K.out)[pos] = w1; break;
          case PH_I128: ((ulonglong2*)K.out)[pos] = make_ulonglong2(w1, w2); break;
          default: break;
        }
      }
    }
  }
  __syncthreads();
  if (tid == 0) {
    p.meta[0] = (unsigned long long)G;
    p.meta[1] = (unsigned long long)s_eval_err;
    p.meta[2] = (unsigned long long)x_err;
  }
  if (tid < p.n_aggs) p.meta[3 + tid] = s.nulls[tid];
  if (tid < nk) p.meta[3 + EPI_MAXAGG + tid] = s.key_nulls[tid];

  // ---- R: the accumulator table is ready for the next execution -----------------------------------------------------------
  const int total = p.src == EPI_SRC_DENSE ? p.n_slots * NA2 : 0;
  for (int i = tid; i < total; i += EPI_NT) {
    p.g_lo[i] = (unsigned long long)p.init[i % NA2];
    p.g_hi[i] = 0;
  }
}

size_t epilogue_smem_bytes(int cap) {
  return (size_t)cap * (8 + 4 + 4 * 2) + (EPI_MAXAGG + EPI_MAXK) * 8 + (EPI_NT + COMM_MAX_WORLD + 1 + 7) * 4 + 64;
}

void launch_dense_epilogue(Ctx* ctx, const EpiParams& p) {
  const size_t smem = epilogue_smem_bytes(p.smem_cap);
  if (smem > ctx->epi_smem_set) {
    CUDA_CHECK(cudaFuncSetAttribute(k_dense_epilogue, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 48 << 10)));
    ctx->epi_smem_set = std::max<size_t>(smem, 48 << 10);
  }
  // on the epilogue stream, behind everything queued on the compute stream so far
  CUDA_CHECK(cudaEventRecord(ctx->epi_ready, ctx->stream));
  CUDA_CHECK(cudaStreamWaitEvent(ctx->epi_stream, ctx->epi_ready, 0));
  ctx->launches++;
  k_dense_epilogue<<<1, EPI_NT, smem, ctx->epi_stream>>>(p);
  CUDA_CHECK(cudaGetLastError());
  debug_sync_launch(ctx, "k_dense_epilogue");
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
void epilogue_describe(Ctx* ctx, const std::vector<DType>& key_types, const std::vector<AggSpec>& specs, const std::vector<int>& kinds,
                       const Schema& out_schema, EpiParams& E) {
  const size_t nk = key_types.size(), na = specs.size();
  if (nk > (size_t)EPI_MAXK || na > (size_t)EPI_MAXAGG) throw_internal("too many keys / aggregates for the fused epilogue");
  if (out_schema.fields.size() != nk + na) throw_arrow("aggregate output schema has the wrong number of fields");
  E.n_keys = (int)nk;
  E.n_aggs = (int)na;
  E.grouped = nk > 0 ? 1 : 0;
  E.rw = epi_rec_words((int)nk, (int)na);
  for (size_t k = 0; k < nk; ++k) {
    if (key_types[k] != out_schema.fields[k].type)
      throw_arrow("column types must match schema types, expected " + out_schema.fields[k].type.str() + " but found " +
                  key_types[k].str() + " at column index " + std::to_string(k));
    E.key[k].out_phys = out_phys_of(key_types[k]);
    if (E.key[k].out_phys == PH_NULL || E.key[k].out_phys == PH_BIT || E.key[k].out_phys == PH_F32 || E.key[k].out_phys == PH_F64)
      throw_internal("unsupported group key type for sharded execution: " + key_types[k].str());
  }
  for (size_t i = 0; i < na; ++i) {
    const AggSpec& a = specs[i];
    const Field& of = out_schema.fields[nk + i];
    DType produced = a.return_type;
    if (a.op == QGPU_AGG_COUNT) produced = mk_type(QGPU_T_INT64);
    if (produced != of.type)
      throw_arrow("column types must match schema types, expected " + of.type.str() + " but found " + produced.str() +
                  " at column index " + std::to_string(nk + i));
    FinSpec& f = E.agg[i];
    memset(&f, 0, sizeof(f));
    f.op = a.op;
    f.kind = kinds[i];
    f.out_phys = out_phys_of(produced);
    f.sum_scale = a.expr_type.scale;
    f.target_scale = a.return_type.scale;
    f.target_prec = a.return_type.precision;
    f.compat_avg = ctx->compat_avg_precision ? 1 : 0;
    f.no_input = 0;
    E.sent_lo[i] = E.sent_hi[i] = 0;
    if (a.op == QGPU_AGG_MIN || a.op == QGPU_AGG_MAX) minmax_sentinel(a.arg->result_type, a.op == QGPU_AGG_MIN, &E.sent_lo[i], &E.sent_hi[i]);
  }
}

View epilogue_execute(Ctx* ctx, EpiParams E, const std::vector<DType>& key_types, const std::vector<AggSpec>& specs,
                      const Schema& out_schema, bool sharded, int64_t row_offset, int max_groups, DBufP packed_rec) {
  const int nk = E.n_keys, na = E.n_aggs;
  Comm* comm = sharded ? ctx->comm.get() : nullptr;
  if (sharded && !comm) throw QError(QGPU_ERR_NCCL, "NcclError: no communicator: call qgpu_comm_init first");
  E.world = comm ? comm->world : 1;
  E.rank = comm ? comm->rank : 0;
  if (!E.grouped) max_groups = 1;
  if (max_groups < 1 || max_groups > EPI_MAXG) throw_internal("max_groups must be in [1, 4096]");
  E.max_groups = max_groups;
  E.g_max = E.grouped ? (int)std::min<int64_t>((int64_t)EPI_MAXG, (int64_t)E.world * max_groups) : 1;
  E.row_offset = row_offset;
  // the shared-memory arrays hold up to world x max_groups gathered records (ungrouped: one record per rank)
  E.smem_cap = std::max(std::max(E.g_max, max_groups), (int)std::min<int64_t>((int64_t)EPI_MAXG, (int64_t)E.world * max_groups));
  const size_t rec_words = (size_t)EPI_HDR + (size_t)max_groups * E.rw;
  if (comm && rec_words * 8 > COMM_SLOT_BYTES)
    throw_internal("sharded aggregate: the state block of " + std::to_string(max_groups) + " groups exceeds the " +
                   std::to_string(COMM_SLOT_BYTES >> 10) + " KB exchange slot; use hash repartition for high-cardinality keys");
  const int64_t g = E.g_max;
  const size_t words = (size_t)((g + 31) >> 5);
  // ---- one allocation: output columns, state block, merged records, metadata ----------------------------------------
  size_t total = Slab::need(EPI_META_WORDS * 8) + (packed_rec ? 0 : Slab::need(rec_words * 8)) +
                 (E.world > 1 ? Slab::need((size_t)g * E.rw * 8) : 0);
  for (int k = 0; k < nk; ++k) {
    const Phys ph = (Phys)E.key[k].out_phys;
    total += ph == PH_STR ? Slab::need((size_t)g * 16 + 16) + Slab::need((size_t)(g + 1) * 4) : Slab::need((size_t)g * phys_width(ph) + 16);
    if (E.src == EPI_SRC_PACKED) total += Slab::need(words * 4 + 4);
  }
  for (int a = 0; a < na; ++a)
    total += Slab::need((size_t)g * std::max(phys_width((Phys)E.agg[a].out_phys), 1) + 16) + Slab::need(words * 4 + 4);
  Slab slab(ctx, total, false);
  slab.buf->free_stream = ctx->epi_stream;  // last written by the epilogue: freed in that stream's order
  if (packed_rec) packed_rec->free_stream = ctx->epi_stream;
  DBufP meta = slab.take(EPI_META_WORDS * 8);
  E.meta = (unsigned long long*)meta->ptr;
  DBufP rec = packed_rec ? packed_rec : slab.take(rec_words * 8);
  E.rec = (unsigned long long*)rec->ptr;
  if (E.world > 1) E.mrec = (unsigned long long*)slab.take((size_t)g * E.rw * 8)->ptr;
  std::vector<DColP> key_cols, agg_cols;
  for (int k = 0; k < nk; ++k) {
    auto c = std::make_shared<DCol>();
    c->type = key_types[k];
    c->phys = (Phys)E.key[k].out_phys;
    c->length = g;
    if (c->phys == PH_STR) {
      c->data = slab.take((size_t)g * 16 + 16);
      c->offsets = slab.take((size_t)(g + 1) * 4);
      c->str_bytes = g * 16;
      c->str_bytes_is_bound = true;
      c->max_str_len = 16;  // state records carry Utf8 keys of <= 16 bytes
      E.key[k].out_offsets = (int32_t*)c->offsets->ptr;
    } else {
      c->data = slab.take((size_t)g * phys_width(c->phys) + 16);
    }
    E.key[k].out = c->data->ptr;
    E.key[k].out_valid = nullptr;
    if (E.src == EPI_SRC_PACKED) {  // generic sources may carry NULL keys
      c->validity = slab.take(words * 4 + 4);
      E.key[k].out_valid = (uint32_t*)c->validity->ptr;
    }
    key_cols.push_back(c);
  }
  for (int a = 0; a < na; ++a) {
    auto c = std::make_shared<DCol>();
    c->type = specs[a].op == QGPU_AGG_COUNT ? mk_type(QGPU_T_INT64) : specs[a].return_type;
    c->phys = (Phys)E.agg[a].out_phys;
    c->length = g;
    c->data = slab.take((size_t)g * std::max(phys_width(c->phys), 1) + 16);
    c->validity = slab.take(words * 4 + 4);
    E.agg[a].out = c->data->ptr;
    E.agg[a].out_valid = (uint32_t*)c->validity->ptr;
    agg_cols.push_back(c);
  }
  if (comm) {
    E.epoch = ++comm->epoch;
    for (int r = 0; r < comm->world; ++r) E.peer[r] = (unsigned long long*)comm->peer[r];
    const char* t = getenv("QGPU_PEER_TIMEOUT_MS");
    E.timeout_ns = (unsigned long long)(t ? std::max(1, atoi(t)) : 20000) * 1000000ull;
  }
  launch_dense_epilogue(ctx, E);
  // ---- result: metadata pending --------------------------------------------------------------------------------------
  View out;
  out.schema = out_schema;
  out.num_rows = g;
  out.num_batches = 1;
  for (auto& c : key_cols) out.cols.push_back({c, nullptr});
  for (auto& c : agg_cols) out.cols.push_back({c, nullptr});
  std::vector<int> ops;
  for (auto& a : specs) ops.push_back(a.op);
  const bool compat_empty = ctx->compat_empty_decimal_sum;
  out.pending = make_pending(ctx, E.meta, EPI_META_WORDS, [key_cols, agg_cols, ops, compat_empty, meta](const unsigned long long* m, Pending& P) {
    const int64_t G = (int64_t)m[0];
    const int x = (int)m[2];
    if (x == EPI_ERR_TIMEOUT)
      throw QError(QGPU_ERR_NCCL, "NcclError: a peer's state block did not arrive in time (is every rank executing the same sharded plan?)");
    if (x == EPI_ERR_OVERFLOW || x == EPI_ERR_TOO_MANY)
      throw_internal("sharded aggregate: a shard produced more than max_groups groups (or a key longer than 16 bytes); use hash "
                     "repartition for high-cardinality keys");
    if (x) throw_internal("sharded aggregate: malformed state block (" + std::to_string(x) + ")");
    if ((int)m[1]) throw_eval_error((int)m[1]);
    for (size_t k = 0; k < key_cols.size(); ++k) {
      DCol& c = *key_cols[k];
      c.length = G;
      c.null_count = c.validity ? (int64_t)m[3 + EPI_MAXAGG + k] : 0;
      if (c.null_count == 0) c.validity.reset();
    }
    for (size_t a = 0; a < agg_cols.size(); ++a) {
      DCol& c = *agg_cols[a];
      c.length = G;
      c.null_count = (int64_t)m[3 + a];
      if (c.null_count == 0) c.validity.reset();
      if (c.null_count > 0 && ops[a] == QGPU_AGG_SUM && c.type.is_decimal() && compat_empty)
        throw_arrow("column types must match schema types, expected " + c.type.str() + " but found Decimal128(38, 10)");
      if (c.null_count > 0 && (ops[a] == QGPU_AGG_MIN || ops[a] == QGPU_AGG_MAX) && compat_empty)
        throw_arrow("column types must match schema types, expected " + c.type.str() + " but found Null");
    }
    P.num_rows = G;
  }, ctx->epi_stream);
  return out;
}

}  // namespace qgpu
