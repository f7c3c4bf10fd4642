// qref_cpu.cpp -- single-threaded C++ restatement ("port") of the STEPS qurious's CPU operators perform
// for TPC-H Q6 / Q1 / Q3, used ONLY as (a) the reported cpu_baseline / `bench.py --impl reference` arm
// and (b) a second checker in tests/.  TEST/BENCH INFRASTRUCTURE: nothing in qurious_b200/ links or
// calls this file.  The reference itself is Rust and cannot be built here (no cargo/rustc in the image),
// so this is kind = "port", never "reference".
//
// What is restated, step by step (reference file:line, relative to /root/reference/qurious/src):
//   * MemoryTable::scan                 datasource/memory.rs:69-98   per batch: predicate.evaluate +
//                                       filter_record_batch (one output batch per input batch)
//   * Literal::evaluate                 physical/expr/literal.rs:19-23  an N-row array PER BATCH
//   * CastExpr (Utf8->Date32, ...)      physical/expr/cast.rs:32-38     re-parsed for every row of every batch
//   * BinaryExpr cmp / and_kleene / arithmetic  physical/expr/binary.rs:30-71: one fresh array per operator
//   * NoGroupingAggregate               physical/plan/aggregate/no_grouping.rs:30-62
//   * HashAggregate + GroupAccumulator  physical/plan/aggregate/hash.rs:45-107,138-170: concat_batches,
//                                       SipHash-1-3 per row (utils/array.rs:171-210), HashMap<u64,usize>,
//                                       per-group row-id vectors, per group x aggregate take + accumulate
//   * Sum/Avg/Count accumulators        physical/expr/aggregate/{sum,avg,count}.rs
//   * HashJoinExec                      physical/plan/join/hash_join.rs:148-216,354-385: concat build side,
//                                       hashes, chained JoinHashMap (reverse insertion), probe per batch,
//                                       key equality re-check, take of every carried column
// Only the columns a query references are carried (the reference carries all 17/10/9 columns through
// filter/take/concat, so this port UNDER-estimates the reference's cost; said so in bench.py's `sample`).
#include <cstdint>
#include <algorithm>
#include <climits>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

typedef __int128 i128;
typedef unsigned __int128 u128;

// ---------------------------------------------------------------------------------------------
// SipHash-1-3 with zero keys == Rust std::hash::DefaultHasher::new() (hash.rs:18, hash_join.rs:25)
// ---------------------------------------------------------------------------------------------
struct Sip13 {
  uint64_t v0, v1, v2, v3, tail;
  size_t length, ntail;
  Sip13() { reset(); }
  void reset() {
    v0 = 0x736f6d6570736575ULL;
    v1 = 0x646f72616e646f6dULL;
    v2 = 0x6c7967656e657261ULL;
    v3 = 0x7465646279746573ULL;
    tail = 0;
    length = 0;
    ntail = 0;
  }
  static inline uint64_t rotl(uint64_t x, int b) { return (x << b) | (x >> (64 - b)); }
  inline void round() {
    v0 += v1; v1 = rotl(v1, 13); v1 ^= v0; v0 = rotl(v0, 32);
    v2 += v3; v3 = rotl(v3, 16); v3 ^= v2;
    v0 += v3; v3 = rotl(v3, 21); v3 ^= v0;
    v2 += v1; v1 = rotl(v1, 17); v1 ^= v2; v2 = rotl(v2, 32);
  }
  void write(const uint8_t* p, size_t n) {
    length += n;
    size_t i = 0;
    if (ntail) {
      while (ntail < 8 && i < n) tail |= (uint64_t)p[i++] << (8 * ntail++);
      if (ntail < 8) return;
      v3 ^= tail; round(); v0 ^= tail;
      tail = 0; ntail = 0;
    }
    for (; i + 8 <= n; i += 8) {
      uint64_t m;
      memcpy(&m, p + i, 8);
      v3 ^= m; round(); v0 ^= m;
    }
    for (; i < n; ++i) tail |= (uint64_t)p[i] << (8 * ntail++);
  }
  uint64_t finish() const {
    Sip13 s = *this;
    uint64_t b = ((uint64_t)s.length << 56) | s.tail;
    s.v3 ^= b; s.round(); s.v0 ^= b;
    s.v2 ^= 0xff;
    s.round(); s.round(); s.round();
    return s.v0 ^ s.v1 ^ s.v2 ^ s.v3;
  }
};
struct SipU64Hash {  // std HashMap<u64,_> hashes its key with SipHash-1-3 again (RandomState; keys irrelevant to cost)
  size_t operator()(uint64_t k) const {
    Sip13 s;
    s.write((const uint8_t*)&k, 8);
    return (size_t)s.finish();
  }
};

static int64_t days_from_civil(int64_t y, unsigned m, unsigned d) {
  y -= m <= 2;
  const int64_t era = (y >= 0 ? y : y - 399) / 400;
  const unsigned yoe = (unsigned)(y - era * 400);
  const unsigned doy = (153 * (m + (m > 2 ? -3 : 9)) + 2) / 5 + d - 1;
  const unsigned doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
  return era * 146097 + (int64_t)doe - 719468;
}
static int32_t parse_date(const std::string& s) {
  int y = atoi(s.substr(0, 4).c_str()), m = atoi(s.substr(5, 2).c_str()), d = atoi(s.substr(8, 2).c_str());
  return (int32_t)days_from_civil(y, (unsigned)m, (unsigned)d);
}
// CAST(Literal(Utf8) AS Date32) evaluated on one batch: materialise N strings, parse each (literal.rs + cast.rs)
static std::vector<int32_t> date_literal_array(const char* lit, size_t n) {
  std::vector<std::string> strs(n, std::string(lit));
  std::vector<int32_t> out(n);
  for (size_t i = 0; i < n; ++i) out[i] = parse_date(strs[i]);
  return out;
}
template <typename T>
static std::vector<T> const_array(T v, size_t n) { return std::vector<T>(n, v); }

enum { EQ = 0, NE, GT, GE, LT, LE };
template <typename T>
static std::vector<uint8_t> cmp(const T* a, const std::vector<T>& b, size_t n, int op) {
  std::vector<uint8_t> r(n);
  switch (op) {
    case GE: for (size_t i = 0; i < n; ++i) r[i] = a[i] >= b[i]; break;
    case GT: for (size_t i = 0; i < n; ++i) r[i] = a[i] > b[i]; break;
    case LT: for (size_t i = 0; i < n; ++i) r[i] = a[i] < b[i]; break;
    case LE: for (size_t i = 0; i < n; ++i) r[i] = a[i] <= b[i]; break;
    case EQ: for (size_t i = 0; i < n; ++i) r[i] = a[i] == b[i]; break;
    default: for (size_t i = 0; i < n; ++i) r[i] = a[i] != b[i]; break;
  }
  return r;
}
static std::vector<uint8_t> and_k(const std::vector<uint8_t>& a, const std::vector<uint8_t>& b) {
  std::vector<uint8_t> r(a.size());
  for (size_t i = 0; i < a.size(); ++i) r[i] = a[i] & b[i];
  return r;
}
template <typename T>
static void filter_append(const T* src, const std::vector<uint8_t>& mask, std::vector<T>& out) {
  for (size_t i = 0; i < mask.size(); ++i)
    if (mask[i]) out.push_back(src[i]);
}
template <typename T>
static std::vector<T> filter_new(const T* src, const std::vector<uint8_t>& mask) {
  std::vector<T> out;
  filter_append(src, mask, out);
  return out;
}
template <typename T>
static std::vector<T> take(const std::vector<T>& v, const std::vector<uint64_t>& idx) {
  std::vector<T> out(idx.size());
  for (size_t i = 0; i < idx.size(); ++i) out[i] = v[idx[i]];
  return out;
}
static i128 sum128(const std::vector<i128>& v) {
  u128 s = 0;
  for (i128 x : v) s += (u128)x;
  return (i128)s;
}

extern "C" {

// ---------------------------------------------------------------------------------------------
// Q6: Scan(pushed-down 5-term predicate) -> NoGroupingAggregate[SUM(price*discount)]
// ---------------------------------------------------------------------------------------------
void qcpu_q6(int64_t n, int64_t batch_rows, const int32_t* shipdate, const i128* discount, const i128* quantity,
             const i128* price, const char* date_lo, const char* date_hi, double disc_lo, double disc_hi, int64_t qty_lt,
             i128* out_sum, int64_t* out_rows) {
  bool have = false;
  u128 acc = 0;
  int64_t kept = 0;
  for (int64_t b0 = 0; b0 < n; b0 += batch_rows) {
    const size_t m = (size_t)std::min<int64_t>(batch_rows, n - b0);
    // predicate: ((((ship >= d0) AND (ship < d1)) AND (disc >= c0)) AND (disc <= c1)) AND (qty < c2)
    auto d0 = date_literal_array(date_lo, m), d1 = date_literal_array(date_hi, m);
    // CAST(Float64 AS Decimal128(15,2)): round(v * 100) per row of an m-row literal array
    std::vector<double> f0(m, disc_lo), f1(m, disc_hi);
    std::vector<i128> c0(m), c1(m);
    for (size_t i = 0; i < m; ++i) {
      c0[i] = (i128)__builtin_round(f0[i] * 100.0);
      c1[i] = (i128)__builtin_round(f1[i] * 100.0);
    }
    std::vector<int64_t> q0(m, qty_lt);
    std::vector<i128> c2(m);
    for (size_t i = 0; i < m; ++i) c2[i] = (i128)q0[i] * 100;
    auto m0 = cmp(shipdate + b0, d0, m, GE), m1 = cmp(shipdate + b0, d1, m, LT);
    auto m2 = cmp(discount + b0, c0, m, GE), m3 = cmp(discount + b0, c1, m, LE), m4 = cmp(quantity + b0, c2, m, LT);
    auto mask = and_k(and_k(and_k(and_k(m0, m1), m2), m3), m4);
    // filter_record_batch over the referenced columns
    auto fs = filter_new(shipdate + b0, mask);
    auto fd = filter_new(discount + b0, mask);
    auto fq = filter_new(quantity + b0, mask);
    auto fp = filter_new(price + b0, mask);
    kept += (int64_t)fp.size();
    // NoGroupingAggregate: evaluate price * discount on the filtered batch, SumAccumulator::accumluate
    std::vector<i128> prod(fp.size());
    for (size_t i = 0; i < fp.size(); ++i) prod[i] = (i128)((u128)fp[i] * (u128)fd[i]);
    if (!prod.empty()) {
      acc += (u128)sum128(prod);
      have = true;
    }
    (void)fs; (void)fq;
  }
  *out_sum = have ? (i128)acc : 0;
  *out_rows = kept;
}

// ---------------------------------------------------------------------------------------------
// Q1.  Outputs at most `max_groups` groups in first-occurrence order:
//   key bytes (first byte of each Utf8 key -- the generator's keys are 1 char), sums[4], avgs[3], count.
// returns the number of groups.
// ---------------------------------------------------------------------------------------------
int64_t qcpu_q1(int64_t n, int64_t batch_rows, const int32_t* shipdate, const int32_t* rf_off, const char* rf_data,
                const int32_t* ls_off, const char* ls_data, const i128* quantity, const i128* price, const i128* discount,
                const i128* tax, const char* date_le, int64_t max_groups, char* out_rf, char* out_ls, i128* out_sums /*[g][4]*/,
                i128* out_avgs /*[g][3]*/, int64_t* out_count) {
  // ---- Scan with pushed-down filter, one output batch per input batch; then concat_batches (hash.rs:150)
  std::vector<int32_t> c_off_rf{0}, c_off_ls{0};
  std::string c_rf, c_ls;
  std::vector<i128> c_qty, c_price, c_disc, c_tax;
  std::vector<int32_t> c_ship;
  for (int64_t b0 = 0; b0 < n; b0 += batch_rows) {
    const size_t m = (size_t)std::min<int64_t>(batch_rows, n - b0);
    auto d = date_literal_array(date_le, m);
    auto mask = cmp(shipdate + b0, d, m, LE);
    // filter_record_batch -> per-batch arrays
    auto f_ship = filter_new(shipdate + b0, mask);
    auto f_qty = filter_new(quantity + b0, mask), f_price = filter_new(price + b0, mask);
    auto f_disc = filter_new(discount + b0, mask), f_tax = filter_new(tax + b0, mask);
    std::vector<int32_t> f_off_rf{0}, f_off_ls{0};
    std::string f_rf, f_ls;
    for (size_t i = 0; i < m; ++i)
      if (mask[i]) {
        f_rf.append(rf_data + rf_off[b0 + i], (size_t)(rf_off[b0 + i + 1] - rf_off[b0 + i]));
        f_off_rf.push_back((int32_t)f_rf.size());
        f_ls.append(ls_data + ls_off[b0 + i], (size_t)(ls_off[b0 + i + 1] - ls_off[b0 + i]));
        f_off_ls.push_back((int32_t)f_ls.size());
      }
    // concat_batches: second copy
    c_ship.insert(c_ship.end(), f_ship.begin(), f_ship.end());
    c_qty.insert(c_qty.end(), f_qty.begin(), f_qty.end());
    c_price.insert(c_price.end(), f_price.begin(), f_price.end());
    c_disc.insert(c_disc.end(), f_disc.begin(), f_disc.end());
    c_tax.insert(c_tax.end(), f_tax.begin(), f_tax.end());
    for (size_t i = 1; i < f_off_rf.size(); ++i) c_off_rf.push_back((int32_t)c_rf.size() + f_off_rf[i]);
    c_rf += f_rf;
    for (size_t i = 1; i < f_off_ls.size(); ++i) c_off_ls.push_back((int32_t)c_ls.size() + f_off_ls[i]);
    c_ls += f_ls;
  }
  const size_t rows = c_qty.size();
  if (rows == 0) return 0;
  // ---- aggregate argument expressions on the one big batch (hash.rs:152-162), one fresh array per operator
  std::vector<int64_t> lit1(rows, 1);
  std::vector<i128> one20(rows);
  for (size_t i = 0; i < rows; ++i) one20[i] = (i128)lit1[i];  // CAST(Int64(1) AS Decimal128(20,0))
  std::vector<i128> one_minus(rows), disc_price(rows), one_plus(rows), charge(rows);
  for (size_t i = 0; i < rows; ++i) one_minus[i] = one20[i] * 100 - c_disc[i];          // sub: rescale to scale 2
  for (size_t i = 0; i < rows; ++i) disc_price[i] = (i128)((u128)c_price[i] * (u128)one_minus[i]);
  std::vector<i128> one20b(rows);
  for (size_t i = 0; i < rows; ++i) one20b[i] = (i128)lit1[i];
  for (size_t i = 0; i < rows; ++i) one_plus[i] = one20b[i] * 100 + c_tax[i];
  for (size_t i = 0; i < rows; ++i) charge[i] = (i128)((u128)disc_price[i] * (u128)one_plus[i]);
  std::vector<int64_t> count_arg(rows, 1);  // COUNT(Int64(1)): an N-row literal array
  const std::vector<i128>* args[7] = {&c_qty, &c_price, &disc_price, &charge, &c_qty, &c_price, &c_disc};
  // ---- GroupAccumulator::update: hashers (72 B each), create_hashes, HashMap<u64,usize>, per-group row ids
  std::vector<Sip13> hashers(rows);
  for (size_t i = 0; i < rows; ++i) {
    hashers[i].write((const uint8_t*)c_rf.data() + c_off_rf[i], (size_t)(c_off_rf[i + 1] - c_off_rf[i]));
    const uint8_t ff = 0xff;
    hashers[i].write(&ff, 1);  // impl Hash for str appends 0xff
  }
  for (size_t i = 0; i < rows; ++i) {
    hashers[i].write((const uint8_t*)c_ls.data() + c_off_ls[i], (size_t)(c_off_ls[i + 1] - c_off_ls[i]));
    const uint8_t ff = 0xff;
    hashers[i].write(&ff, 1);
  }
  std::vector<uint64_t> hashes(rows);
  for (size_t i = 0; i < rows; ++i) hashes[i] = hashers[i].finish();
  std::unordered_map<uint64_t, size_t, SipU64Hash> map;
  std::unordered_map<uint64_t, std::vector<uint64_t>, SipU64Hash> accs_indices;
  std::vector<size_t> group_first;  // first-occurrence order (the reference's order is unspecified)
  for (size_t row = 0; row < rows; ++row) {
    auto it = map.find(hashes[row]);
    if (it != map.end()) {
      accs_indices[it->second].push_back(row);
    } else {
      map.emplace(hashes[row], row);
      accs_indices[row] = std::vector<uint64_t>{row};
      group_first.push_back(row);
    }
  }
  // ---- per group x aggregate: clone the index vector, take, accumulate
  int64_t g = 0;
  for (size_t first : group_first) {
    if (g >= max_groups) break;
    const std::vector<uint64_t>& idx = accs_indices[first];
    i128 sums[7];
    for (int a = 0; a < 7; ++a) {
      std::vector<uint64_t> indices(idx);
      std::vector<i128> taken = take(*args[a], indices);
      sums[a] = sum128(taken);
    }
    std::vector<uint64_t> indices(idx);
    std::vector<int64_t> taken = take(count_arg, indices);
    const int64_t count = (int64_t)taken.size();
    out_rf[g] = c_rf[(size_t)c_off_rf[first]];
    out_ls[g] = c_ls[(size_t)c_off_ls[first]];
    for (int a = 0; a < 4; ++a) out_sums[g * 4 + a] = sums[a];
    // DecimalAvgAccumulator (avg.rs:89-116): sum * 10^(target_scale - scale) / count, truncating
    for (int a = 0; a < 3; ++a) out_avgs[g * 3 + a] = (sums[4 + a] * 10000) / (i128)count;
    out_count[g] = count;
    ++g;
  }
  return g;
}

// ---------------------------------------------------------------------------------------------
// Synthetic group-by (BASELINE.json configs[3]): HashAggregate[k; SUM(v) COUNT(v) MIN(v) MAX(v) AVG(f)] over a scan
// without predicate.  Outputs one row per group in first-occurrence order; returns the number of groups.
// hash.rs:45-107,138-170 + sum.rs/count.rs/min.rs/max.rs/avg.rs
// ---------------------------------------------------------------------------------------------
int64_t qcpu_groupby(int64_t n, int64_t batch_rows, const int64_t* k, const int64_t* v, const double* f, int64_t max_groups,
                     int64_t* out_k, int64_t* out_sum, int64_t* out_cnt, int64_t* out_min, int64_t* out_max, double* out_avg) {
  // MemoryTable::scan without filter hands the batches through; concat_batches copies them into one (hash.rs:150)
  std::vector<int64_t> ck, cv;
  std::vector<double> cf;
  for (int64_t b0 = 0; b0 < n; b0 += batch_rows) {
    const size_t m = (size_t)std::min<int64_t>(batch_rows, n - b0);
    ck.insert(ck.end(), k + b0, k + b0 + m);
    cv.insert(cv.end(), v + b0, v + b0 + m);
    cf.insert(cf.end(), f + b0, f + b0 + m);
  }
  const size_t rows = ck.size();
  if (rows == 0) return 0;
  std::vector<Sip13> hashers(rows);
  for (size_t i = 0; i < rows; ++i) hashers[i].write((const uint8_t*)&ck[i], 8);
  std::vector<uint64_t> hashes(rows);
  for (size_t i = 0; i < rows; ++i) hashes[i] = hashers[i].finish();
  std::unordered_map<uint64_t, size_t, SipU64Hash> map;
  std::unordered_map<uint64_t, std::vector<uint64_t>, SipU64Hash> accs_indices;
  std::vector<size_t> group_first;
  for (size_t row = 0; row < rows; ++row) {
    auto it = map.find(hashes[row]);
    if (it != map.end()) {
      accs_indices[it->second].push_back(row);
    } else {
      map.emplace(hashes[row], row);
      accs_indices[row] = std::vector<uint64_t>{row};
      group_first.push_back(row);
    }
  }
  int64_t g = 0;
  for (size_t first : group_first) {
    if (g >= max_groups) break;
    const std::vector<uint64_t>& idx = accs_indices[first];
    int64_t sum = 0, mn = INT64_MAX, mx = INT64_MIN, cnt = 0;
    {
      std::vector<uint64_t> indices(idx);
      std::vector<int64_t> t = take(cv, indices);
      uint64_t acc = 0;
      for (int64_t x : t) acc += (uint64_t)x;  // add_wrapping
      sum = (int64_t)acc;
    }
    {
      std::vector<uint64_t> indices(idx);
      std::vector<int64_t> t = take(cv, indices);
      cnt = (int64_t)t.size();
    }
    {
      std::vector<uint64_t> indices(idx);
      std::vector<int64_t> t = take(cv, indices);
      for (int64_t x : t) mn = std::min(mn, x);
    }
    {
      std::vector<uint64_t> indices(idx);
      std::vector<int64_t> t = take(cv, indices);
      for (int64_t x : t) mx = std::max(mx, x);
    }
    double fs = 0;
    int64_t fc = 0;
    {
      std::vector<uint64_t> indices(idx);
      std::vector<double> t = take(cf, indices);
      for (double x : t) fs += x;
      fc = (int64_t)t.size();
    }
    out_k[g] = ck[first];
    out_sum[g] = sum;
    out_cnt[g] = cnt;
    out_min[g] = mn;
    out_max[g] = mx;
    out_avg[g] = fs / (double)fc;
    ++g;
  }
  return g;
}

// ---------------------------------------------------------------------------------------------
// Q3: customer(BUILDING) JOIN orders(o_orderdate < d) JOIN lineitem(l_shipdate > d)
//   -> HashAggregate[l_orderkey, o_orderdate, o_shippriority; SUM(l_extendedprice * (1 - l_discount))]
// hash_join.rs:148-216,354-385 (build = left, chained JoinHashMap, probe per right batch, key equality re-check,
// take of the carried columns), hash.rs as above.  Outputs groups in first-occurrence order; returns their number.
// ---------------------------------------------------------------------------------------------
struct ChainMap {  // JoinHashMap (hash_join.rs:40-107): hash -> last inserted row + 1, next[row] = previous head
  std::unordered_map<uint64_t, uint64_t, SipU64Hash> map;
  std::vector<uint64_t> next;
  void build(const std::vector<uint64_t>& hashes) {
    next.assign(hashes.size(), 0);
    for (size_t r = hashes.size(); r-- > 0;) {  // inserted in reverse so that the chains ascend
      auto it = map.find(hashes[r]);
      if (it == map.end()) {
        map.emplace(hashes[r], (uint64_t)r + 1);
      } else {
        next[r] = it->second;
        it->second = (uint64_t)r + 1;
      }
    }
  }
};
static std::vector<uint64_t> sip_i64(const int64_t* key, size_t n) {
  std::vector<Sip13> hs(n);
  for (size_t i = 0; i < n; ++i) hs[i].write((const uint8_t*)&key[i], 8);
  std::vector<uint64_t> out(n);
  for (size_t i = 0; i < n; ++i) out[i] = hs[i].finish();
  return out;
}

int64_t qcpu_q3(int64_t n_c, const int64_t* c_custkey, const int32_t* seg_off, const char* seg_data, int64_t n_o,
                const int64_t* o_orderkey, const int64_t* o_custkey, const int32_t* o_orderdate, const int64_t* o_shippriority,
                int64_t n_l, const int64_t* l_orderkey, const int32_t* l_shipdate, const i128* l_price, const i128* l_discount,
                int64_t batch_rows, const char* segment, const char* date, int64_t max_groups, int64_t* out_orderkey,
                i128* out_revenue, int32_t* out_orderdate, int64_t* out_shippriority) {
  const std::string seg(segment);
  // ---- customer scan: c_mktsegment = 'BUILDING' (literal array of N strings per batch, string compare) --------
  std::vector<int64_t> b1_custkey;
  for (int64_t b0 = 0; b0 < n_c; b0 += batch_rows) {
    const size_t m = (size_t)std::min<int64_t>(batch_rows, n_c - b0);
    std::vector<std::string> lit(m, seg);
    std::vector<uint8_t> mask(m);
    for (size_t i = 0; i < m; ++i) {
      const int32_t a = seg_off[b0 + i], e = seg_off[b0 + i + 1];
      mask[i] = (size_t)(e - a) == lit[i].size() && memcmp(seg_data + a, lit[i].data(), lit[i].size()) == 0;
    }
    auto fk = filter_new(c_custkey + b0, mask);
    b1_custkey.insert(b1_custkey.end(), fk.begin(), fk.end());  // concat of the build side (hash_join.rs:154)
  }
  ChainMap m1;
  m1.build(sip_i64(b1_custkey.data(), b1_custkey.size()));
  // ---- J1 probe: orders batches (filter o_orderdate < date), output columns o_orderkey, o_orderdate, o_shippriority
  std::vector<int64_t> j1_orderkey, j1_prio;
  std::vector<int32_t> j1_date;
  for (int64_t b0 = 0; b0 < n_o; b0 += batch_rows) {
    const size_t m = (size_t)std::min<int64_t>(batch_rows, n_o - b0);
    auto d = date_literal_array(date, m);
    auto mask = cmp(o_orderdate + b0, d, m, LT);
    auto f_ok = filter_new(o_orderkey + b0, mask), f_ck = filter_new(o_custkey + b0, mask), f_pr = filter_new(o_shippriority + b0, mask);
    auto f_dt = filter_new(o_orderdate + b0, mask);
    if (f_ok.empty()) continue;
    auto ph = sip_i64(f_ck.data(), f_ck.size());
    std::vector<uint64_t> bi, pi;  // candidate (build, probe) pairs by hash, ascending build rows per probe row
    for (size_t r = 0; r < ph.size(); ++r) {
      auto it = m1.map.find(ph[r]);
      if (it == m1.map.end()) continue;
      for (uint64_t cur = it->second; cur != 0; cur = m1.next[cur - 1]) {
        bi.push_back(cur - 1);
        pi.push_back(r);
      }
    }
    // true key equality: take both key columns, eq, filter the index pairs (hash_join.rs:177-216)
    std::vector<int64_t> bk = take(b1_custkey, bi), pk = take(f_ck, pi);
    std::vector<uint64_t> pi2;
    for (size_t i = 0; i < bk.size(); ++i)
      if (bk[i] == pk[i]) pi2.push_back(pi[i]);
    // build_batch_from_indices: take of the carried columns
    auto t_ok = take(f_ok, pi2);
    auto t_pr = take(f_pr, pi2);
    auto t_dt = take(f_dt, pi2);
    j1_orderkey.insert(j1_orderkey.end(), t_ok.begin(), t_ok.end());
    j1_prio.insert(j1_prio.end(), t_pr.begin(), t_pr.end());
    j1_date.insert(j1_date.end(), t_dt.begin(), t_dt.end());
  }
  ChainMap m2;
  m2.build(sip_i64(j1_orderkey.data(), j1_orderkey.size()));
  // ---- J2 probe: lineitem batches (filter l_shipdate > date) -----------------------------------------------------
  std::vector<int64_t> j2_orderkey, j2_prio;
  std::vector<int32_t> j2_date;
  std::vector<i128> j2_price, j2_disc;
  for (int64_t b0 = 0; b0 < n_l; b0 += batch_rows) {
    const size_t m = (size_t)std::min<int64_t>(batch_rows, n_l - b0);
    auto d = date_literal_array(date, m);
    auto mask = cmp(l_shipdate + b0, d, m, GT);
    auto f_ok = filter_new(l_orderkey + b0, mask);
    auto f_price = filter_new(l_price + b0, mask), f_disc = filter_new(l_discount + b0, mask);
    if (f_ok.empty()) continue;
    auto ph = sip_i64(f_ok.data(), f_ok.size());
    std::vector<uint64_t> bi, pi;
    for (size_t r = 0; r < ph.size(); ++r) {
      auto it = m2.map.find(ph[r]);
      if (it == m2.map.end()) continue;
      for (uint64_t cur = it->second; cur != 0; cur = m2.next[cur - 1]) {
        bi.push_back(cur - 1);
        pi.push_back(r);
      }
    }
    std::vector<int64_t> bk = take(j1_orderkey, bi), pk = take(f_ok, pi);
    std::vector<uint64_t> bi2, pi2;
    for (size_t i = 0; i < bk.size(); ++i)
      if (bk[i] == pk[i]) {
        bi2.push_back(bi[i]);
        pi2.push_back(pi[i]);
      }
    auto t_ok = take(f_ok, pi2);
    auto t_price = take(f_price, pi2), t_disc = take(f_disc, pi2);
    auto t_date = take(j1_date, bi2);
    auto t_prio = take(j1_prio, bi2);
    j2_orderkey.insert(j2_orderkey.end(), t_ok.begin(), t_ok.end());
    j2_price.insert(j2_price.end(), t_price.begin(), t_price.end());
    j2_disc.insert(j2_disc.end(), t_disc.begin(), t_disc.end());
    j2_date.insert(j2_date.end(), t_date.begin(), t_date.end());
    j2_prio.insert(j2_prio.end(), t_prio.begin(), t_prio.end());
  }
  const size_t rows = j2_orderkey.size();
  if (rows == 0) return 0;
  // ---- HashAggregate: revenue = l_extendedprice * (CAST(1 AS Decimal(20,0)) - l_discount) ---------------------------
  std::vector<int64_t> lit1(rows, 1);
  std::vector<i128> one20(rows), one_minus(rows), revenue(rows);
  for (size_t i = 0; i < rows; ++i) one20[i] = (i128)lit1[i];
  for (size_t i = 0; i < rows; ++i) one_minus[i] = one20[i] * 100 - j2_disc[i];
  for (size_t i = 0; i < rows; ++i) revenue[i] = (i128)((u128)j2_price[i] * (u128)one_minus[i]);
  std::vector<Sip13> hashers(rows);
  for (size_t i = 0; i < rows; ++i) hashers[i].write((const uint8_t*)&j2_orderkey[i], 8);
  for (size_t i = 0; i < rows; ++i) hashers[i].write((const uint8_t*)&j2_date[i], 4);
  for (size_t i = 0; i < rows; ++i) hashers[i].write((const uint8_t*)&j2_prio[i], 8);
  std::vector<uint64_t> hashes(rows);
  for (size_t i = 0; i < rows; ++i) hashes[i] = hashers[i].finish();
  std::unordered_map<uint64_t, size_t, SipU64Hash> map;
  std::unordered_map<uint64_t, std::vector<uint64_t>, SipU64Hash> accs_indices;
  std::vector<size_t> group_first;
  for (size_t row = 0; row < rows; ++row) {
    auto it = map.find(hashes[row]);
    if (it != map.end()) {
      accs_indices[it->second].push_back(row);
    } else {
      map.emplace(hashes[row], row);
      accs_indices[row] = std::vector<uint64_t>{row};
      group_first.push_back(row);
    }
  }
  int64_t g = 0;
  for (size_t first : group_first) {
    if (g >= max_groups) break;
    std::vector<uint64_t> indices(accs_indices[first]);
    std::vector<i128> taken = take(revenue, indices);
    out_orderkey[g] = j2_orderkey[first];
    out_revenue[g] = sum128(taken);
    out_orderdate[g] = j2_date[first];
    out_shippriority[g] = j2_prio[first];
    ++g;
  }
  return g;
}

}  // extern "C"
