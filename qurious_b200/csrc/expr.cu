// IR parser, type binder, constant folder and program compiler for physical expressions.
//
// Typing rules restate arrow-rs (the crate the reference calls into):
//   * comparison operands must have identical DataType (decimals identical (p,s));
//   * integer/float arithmetic requires identical operand types, wraps for ints;
//   * Decimal128 add/sub: s = max(s1,s2), p = min(38, max(p1-s1,p2-s2)+s+1); mul: s = s1+s2,
//     p = min(38, p1+p2+1); pinned by binary.rs:197-251 (Decimal(15,2)*(..-..) -> Decimal(32,4));
//   * Div with a decimal side is evaluated in Float64 (binary.rs:54-67).
#include <cstring>

#include "expr.h"
#include "plan.h"

namespace qgpu {

std::string DType::str() const {
  static const char* names[] = {"Null", "Boolean", "Int8", "Int16", "Int32", "Int64", "UInt8", "UInt16",
                                "UInt32", "UInt64", "Float32", "Float64", "Utf8", "Date32", "Date64"};
  if (id == QGPU_T_DECIMAL128) return "Decimal128(" + std::to_string(precision) + ", " + std::to_string(scale) + ")";
  if (id <= QGPU_T_DATE64) return names[id];
  if (id == QGPU_T_TIME32) return scale == 1 ? "Time32(Millisecond)" : "Time32(Second)";
  if (id == QGPU_T_TIME64) return scale == 3 ? "Time64(Nanosecond)" : "Time64(Microsecond)";
  return "Unknown(" + std::to_string(id) + ")";
}

int arrow_width(const DType& t) {
  switch (t.id) {
    case QGPU_T_INT8: case QGPU_T_UINT8: return 1;
    case QGPU_T_INT16: case QGPU_T_UINT16: return 2;
    case QGPU_T_INT32: case QGPU_T_UINT32: case QGPU_T_FLOAT32: case QGPU_T_DATE32: case QGPU_T_TIME32: return 4;
    case QGPU_T_INT64: case QGPU_T_UINT64: case QGPU_T_FLOAT64: case QGPU_T_DATE64: case QGPU_T_TIME64: return 8;
    case QGPU_T_DECIMAL128: return 16;
    default: return 0;
  }
}

VClass class_of(const DType& t) {
  if (t.id == QGPU_T_BOOL) return VC_BOOL;
  if (t.is_signed_int() || t.is_date() || t.is_time()) return VC_INT;
  if (t.is_unsigned_int()) return VC_UINT;
  if (t.is_decimal()) return VC_DEC;
  if (t.is_float()) return VC_FLT;
  if (t.id == QGPU_T_UTF8) return VC_STR;
  return VC_NULLT;
}

static int int_bits(const DType& t) {
  switch (t.id) {
    case QGPU_T_INT8: case QGPU_T_UINT8: return 8;
    case QGPU_T_INT16: case QGPU_T_UINT16: return 16;
    case QGPU_T_INT32: case QGPU_T_UINT32: case QGPU_T_DATE32: case QGPU_T_TIME32: return 32;
    default: return 64;
  }
}

// ------------------------------------------------------------------------------------------------
// parser
// ------------------------------------------------------------------------------------------------
namespace {
struct Reader {
  const uint8_t* p;
  const uint8_t* end;
  void need(size_t n) {
    if ((size_t)(end - p) < n) throw_internal("malformed expression IR (truncated)");
  }
  uint8_t u8() {
    need(1);
    return *p++;
  }
  template <typename T>
  T rd() {
    need(sizeof(T));
    T v;
    memcpy(&v, p, sizeof(T));
    p += sizeof(T);
    return v;
  }
  DType type() {
    DType t;
    t.id = u8();
    t.precision = u8();
    t.scale = (int8_t)u8();
    if (t.id > QGPU_T_DECIMAL128) throw_internal("malformed expression IR (bad type id)");
    return t;
  }
};
}  // namespace

std::unique_ptr<ExprNode> parse_ir(const uint8_t* ir, size_t len) {
  Reader r{ir, ir + len};
  std::vector<std::unique_ptr<ExprNode>> st;
  auto pop = [&]() {
    if (st.empty()) throw_internal("malformed expression IR (stack underflow)");
    auto n = std::move(st.back());
    st.pop_back();
    return n;
  };
  while (r.p < r.end) {
    auto n = std::make_unique<ExprNode>();
    n->kind = r.u8();
    switch (n->kind) {
      case QGPU_IR_COLUMN: n->col_index = (int)r.rd<uint32_t>(); break;
      case QGPU_IR_LITERAL: {
        n->lit_type = r.type();
        n->lit_null = r.u8() != 0;
        if (n->lit_type.id == QGPU_T_NULL) n->lit_null = true;
        if (!n->lit_null) {
          if (n->lit_type.id == QGPU_T_DECIMAL128) {
            n->lit_lo = r.rd<uint64_t>();
            n->lit_hi = r.rd<uint64_t>();
          } else if (n->lit_type.id == QGPU_T_UTF8) {
            uint32_t l = r.rd<uint32_t>();
            r.need(l);
            n->lit_str.assign((const char*)r.p, l);
            r.p += l;
          } else {
            n->lit_lo = r.rd<uint64_t>();
          }
        }
        break;
      }
      case QGPU_IR_BINARY: {
        n->op = r.u8();
        if (n->op > 12) throw_internal("malformed expression IR (bad operator)");
        auto rr = pop();
        auto ll = pop();
        n->children.push_back(std::move(ll));
        n->children.push_back(std::move(rr));
        break;
      }
      case QGPU_IR_CAST: {
        n->cast_type = r.type();
        n->children.push_back(pop());
        break;
      }
      case QGPU_IR_CASE: {
        n->n_when = (int)r.rd<uint32_t>();
        size_t need = 2 * (size_t)n->n_when + 1;
        if (st.size() < need) throw_internal("malformed expression IR (CASE operands)");
        size_t base = st.size() - need;
        for (size_t i = base; i < st.size(); ++i) n->children.push_back(std::move(st[i]));
        st.resize(base);
        break;
      }
      case QGPU_IR_IS_NULL:
      case QGPU_IR_IS_NOT_NULL:
      case QGPU_IR_NEGATIVE: n->children.push_back(pop()); break;
      case QGPU_IR_LIKE: {
        n->op = r.u8();  // negated
        auto pat = pop();
        auto val = pop();
        n->children.push_back(std::move(val));
        n->children.push_back(std::move(pat));
        break;
      }
      case QGPU_IR_EXTRACT: {
        n->op = r.u8();  // 0 year, 1 month, 2 day
        if (n->op > 2) throw_internal("Date part not supported");
        n->children.push_back(pop());
        break;
      }
      case QGPU_IR_SUBQUERY: {
        const qgpu_plan* sp = (const qgpu_plan*)(uintptr_t)r.rd<uint64_t>();
        if (!sp || !sp->node) throw_internal("SubQuery: null plan");
        n->subplan = sp->node;
        break;
      }
      default: throw_internal("malformed expression IR (unknown opcode " + std::to_string(n->kind) + ")");
    }
    st.push_back(std::move(n));
  }
  if (st.size() != 1) throw_internal("malformed expression IR (expected exactly one root)");
  return std::move(st[0]);
}

// ------------------------------------------------------------------------------------------------
// compiler
// ------------------------------------------------------------------------------------------------
DType decimal_result_type(int op, const DType& l, const DType& r) {
  int p1 = l.precision, s1 = l.scale, p2 = r.precision, s2 = r.scale;
  if (op == 8 || op == 9) {
    int s = std::max(s1, s2);
    int p = std::min(38, std::max(p1 - s1, p2 - s2) + s + 1);
    return mk_type(QGPU_T_DECIMAL128, p, s);
  }
  if (op == 10) {
    int s = s1 + s2;
    if (s > 38)
      throw_arrow("Invalid argument error: Output scale of " + l.str() + " * " + r.str() + " would exceed max scale of 38");
    return mk_type(QGPU_T_DECIMAL128, std::min(38, p1 + p2 + 1), s);
  }
  int s = std::max(s1, s2);
  int p = std::min(38, std::min(p1 - s1, p2 - s2) + s);
  return mk_type(QGPU_T_DECIMAL128, std::max(p, 1), s);
}

[[noreturn]] void throw_eval_error(int code) {
  switch (code) {
    case EE_DIV_ZERO: throw_arrow("Divide by zero error");
    case EE_CAST: throw_arrow("Cast error: value out of range for the target type");
    case EE_OVERFLOW: throw_arrow("Compute error: Overflow happened");
    case EE_DEC_PRECISION: throw_arrow("Invalid argument error: value is too large to store in the target Decimal128 precision");
    case EE_PARSE: throw_arrow("Cast error: Cannot cast string to value of the target type");
    default: throw_internal("expression evaluation failed");
  }
}

namespace {
const char* kOpNames[13] = {"=", "!=", ">", ">=", "<", "<=", "AND", "OR", "+", "-", "*", "/", "%"};

struct Compiler {
  const Schema& schema;
  Compiled& out;
  Program& P;
  explicit Compiler(const Schema& s, Compiled& c) : schema(s), out(c), P(c.prog) {}

  struct R {
    DType type;
    bool is_const;
    std::string disp;
  };

  int add_const(const Val& v, bool is_str) {
    if (P.n_consts >= QGPU_MAX_CONSTS) throw_internal("expression too large (constants)");
    P.consts[P.n_consts] = v;
    out.const_is_str.push_back(is_str ? 1 : 0);
    return P.n_consts++;
  }
  Op& add_op(uint8_t code) {
    if (P.n_ops >= QGPU_MAX_OPS) throw_internal("expression too large (ops)");
    Op& o = P.ops[P.n_ops++];
    memset(&o, 0, sizeof(Op));
    o.code = code;
    return o;
  }
  int slot_of(int col) {
    for (size_t i = 0; i < out.col_slots.size(); ++i)
      if (out.col_slots[i] == col) return (int)i;
    if (out.col_slots.size() >= QGPU_MAX_COLS) throw_internal("expression references too many columns");
    out.col_slots.push_back(col);
    return (int)out.col_slots.size() - 1;
  }

  void check_cast(const DType& from, const DType& to) {
    VClass fc = class_of(from);
    bool ok = false;
    bool to_int = to.is_int();
    switch (fc) {
      case VC_NULLT: ok = true; break;
      case VC_BOOL: ok = to_int || to.id == QGPU_T_BOOL; break;
      case VC_INT:
      case VC_UINT:
        ok = to_int || to.is_float() || to.is_decimal() || (to.is_date() && from.is_int()) || ((from.is_date() || from.is_time()) && to == from);
        break;
      case VC_FLT: ok = to_int || to.is_float() || to.is_decimal(); break;
      case VC_DEC: ok = to_int || to.is_float() || to.is_decimal(); break;
      case VC_STR: ok = to.id == QGPU_T_DATE32 || to_int || to.id == QGPU_T_UTF8; break;
    }
    if (!ok) throw_arrow("Cast error: Casting from " + from.str() + " to " + to.str() + " not supported");
  }

  void emit_cast(const DType& from, const DType& to) {
    Op& o = add_op(OP_CAST);
    o.vclass = class_of(from);
    o.from_id = from.id;
    o.from_scale = from.scale;
    o.to_id = to.id;
    o.to_prec = to.precision;
    o.to_scale = to.scale;
  }

  // Evaluate ops [begin, n_ops) on the host (they reference no column) and replace them by a constant.
  void fold(int begin, const DType& type) {
    Program tmp;
    memset(&tmp, 0, sizeof(int32_t) * 4);
    tmp.n_ops = P.n_ops - begin;
    memcpy(tmp.ops, P.ops + begin, sizeof(Op) * tmp.n_ops);
    tmp.n_consts = P.n_consts;
    for (int i = 0; i < P.n_consts; ++i) {
      tmp.consts[i] = P.consts[i];
      if (out.const_is_str[i]) tmp.consts[i].lo = (uint64_t)(out.blob.data() + P.consts[i].lo);
    }
    tmp.n_cols = 0;
    int err = 0;
    Val v = eval_row(tmp, 0, &err);
    if (err) {
      if (out.deferred_err < err) out.deferred_err = err;
      v.valid = 0;
      v.lo = v.hi = 0;
    }
    bool is_str = class_of(type) == VC_STR;
    if (is_str && v.valid) v.lo = (uint64_t)((const char*)v.lo - out.blob.data());
    if (is_str && !v.valid) v.lo = v.hi = 0;
    P.n_ops = begin;
    Op& o = add_op(OP_CONST);
    o.arg = add_const(v, is_str);
  }

  R emit(const ExprNode& n) {
    const int begin = P.n_ops;
    switch (n.kind) {
      case QGPU_IR_COLUMN: {
        if (n.col_index < 0 || n.col_index >= (int)schema.fields.size())
          throw_internal("PhysicalExpr Column references column at index " + std::to_string(n.col_index) +
                         " (zero-based) but input schema only has " + std::to_string(schema.fields.size()) + " columns");
        Op& o = add_op(OP_COL);
        o.arg = slot_of(n.col_index);
        const Field& f = schema.fields[n.col_index];
        return {f.type, false, f.name + "(" + std::to_string(n.col_index) + ")"};
      }
      case QGPU_IR_SUBQUERY: {  // subquery.rs:15-20: batches[0].column(0)
        if (!n.subplan || n.subplan->schema.fields.empty()) throw_internal("SubQuery plan has no columns");
        out.subqueries.push_back(n.subplan);
        Op& o = add_op(OP_COL);
        o.arg = slot_of(-(1000 + (int)out.subqueries.size() - 1));
        return {n.subplan->schema.fields[0].type, false, "SubQuery"};
      }
      case QGPU_IR_LITERAL: {
        Val v;
        memset(&v, 0, sizeof(v));
        v.valid = n.lit_null ? 0 : 1;
        bool is_str = n.lit_type.id == QGPU_T_UTF8;
        if (!n.lit_null) {
          if (is_str) {
            v.lo = out.blob.size();
            v.hi = n.lit_str.size();
            out.blob.insert(out.blob.end(), n.lit_str.begin(), n.lit_str.end());
            out.blob.push_back(0);
          } else if (n.lit_type.id == QGPU_T_FLOAT32) {
            union { uint64_t u; double d; } c;
            c.u = n.lit_lo;
            c.d = (double)(float)c.d;
            v.lo = c.u;
          } else {
            v.lo = n.lit_lo;
            v.hi = n.lit_hi;
          }
        }
        Op& o = add_op(OP_CONST);
        o.arg = add_const(v, is_str);
        return {n.lit_type, true, n.lit_null ? "NULL" : (is_str ? n.lit_str : "lit")};
      }
      case QGPU_IR_BINARY: {
        const ExprNode& ln = *n.children[0];
        const ExprNode& rn = *n.children[1];
        const int op = n.op;
        R l = emit(ln);
        const bool dec_div_probe = (op == 11);
        // decimal division: both sides are cast to Float64 first (binary.rs:54-67).  We need the right
        // type before deciding, so emit right, then patch: emitting is postfix so the left cast must be
        // placed before the right operand's ops.  Do it by emitting the right side into a scratch tail.
        int mid = P.n_ops;
        R r = emit(rn);
        DType lt = l.type, rt = r.type;
        std::string disp = l.disp + " " + kOpNames[op] + " " + r.disp;
        bool is_const = l.is_const && r.is_const;
        DType res;
        if (op <= 5) {
          if (lt != rt) throw_arrow("Invalid comparison operation: " + lt.str() + " " + kOpNames[op] + " " + rt.str());
          Op& o = add_op(OP_CMP);
          o.sub = (uint8_t)op;
          o.vclass = class_of(lt);
          res = mk_type(QGPU_T_BOOL);
        } else if (op == 6 || op == 7) {
          if (lt.id != QGPU_T_BOOL || rt.id != QGPU_T_BOOL)
            throw_internal(std::string("boolean operands required for ") + kOpNames[op] + ", got " + lt.str() + " and " + rt.str());
          add_op(op == 6 ? OP_AND : OP_OR);
          res = mk_type(QGPU_T_BOOL);
        } else if (lt.is_decimal() || rt.is_decimal()) {
          if (dec_div_probe) {
            // insert CAST(left -> f64) at `mid` by shifting the right operand's ops
            DType f64 = mk_type(QGPU_T_FLOAT64);
            check_cast(lt, f64);
            check_cast(rt, f64);
            int n_right = P.n_ops - mid;
            if (P.n_ops + 2 > QGPU_MAX_OPS) throw_internal("expression too large (ops)");
            memmove(&P.ops[mid + 1], &P.ops[mid], sizeof(Op) * n_right);
            P.n_ops += 1;
            Op o;
            memset(&o, 0, sizeof(Op));
            o.code = OP_CAST;
            o.vclass = class_of(lt);
            o.from_id = lt.id;
            o.from_scale = lt.scale;
            o.to_id = QGPU_T_FLOAT64;
            P.ops[mid] = o;
            emit_cast(rt, f64);
            Op& a = add_op(OP_ARITH);
            a.sub = 11;
            a.vclass = VC_FLT;
            a.wbits = 64;
            res = f64;
          } else {
            if (!(lt.is_decimal() && rt.is_decimal()))
              throw_arrow("Invalid arithmetic operation: " + lt.str() + " " + kOpNames[op] + " " + rt.str());
            res = decimal_result_type(op, lt, rt);
            int cslot = 0;
            if (op != 10) {
              Val lm, rm;
              memset(&lm, 0, sizeof(lm));
              memset(&rm, 0, sizeof(rm));
              lm.valid = rm.valid = 1;
              set_i128(lm, pow10_i128(res.scale - lt.scale));
              set_i128(rm, pow10_i128(res.scale - rt.scale));
              cslot = add_const(lm, false);
              add_const(rm, false);
            }
            Op& a = add_op(OP_ARITH);
            a.sub = (uint8_t)op;
            a.vclass = VC_DEC;
            a.arg = cslot;
          }
        } else {
          if (lt != rt || !(lt.is_int() || lt.is_float()))
            throw_arrow("Invalid arithmetic operation: " + lt.str() + " " + kOpNames[op] + " " + rt.str());
          Op& a = add_op(OP_ARITH);
          a.sub = (uint8_t)op;
          a.vclass = class_of(lt);
          a.wbits = lt.is_float() ? (lt.id == QGPU_T_FLOAT32 ? 32 : 64) : (uint8_t)int_bits(lt);
          res = lt;
        }
        if (is_const) fold(begin, res);
        return {res, is_const, disp};
      }
      case QGPU_IR_CAST: {
        R c = emit(*n.children[0]);
        if (c.type != n.cast_type) {
          check_cast(c.type, n.cast_type);
          emit_cast(c.type, n.cast_type);
        }
        if (c.is_const) fold(begin, n.cast_type);
        return {n.cast_type, c.is_const, "CAST(" + c.disp + " AS " + n.cast_type.str() + ")"};
      }
      case QGPU_IR_CASE: {
        bool is_const = true;
        DType res;
        std::string disp = "CASE";
        for (int i = 0; i < n.n_when; ++i) {
          R w = emit(*n.children[2 * i]);
          if (w.type.id != QGPU_T_BOOL) throw_internal("CASE WHEN must be boolean");
          R t = emit(*n.children[2 * i + 1]);
          if (i == 0) res = t.type;
          else if (t.type != res) throw_arrow("Invalid argument error: arguments need to have the same data type");
          is_const = is_const && w.is_const && t.is_const;
          disp += " WHEN " + w.disp + " THEN " + t.disp;
        }
        R e = emit(*n.children[2 * n.n_when]);
        if (n.n_when == 0) res = e.type;
        else if (e.type != res) throw_arrow("Invalid argument error: arguments need to have the same data type");
        is_const = is_const && e.is_const;
        Op& o = add_op(OP_CASE);
        o.arg = n.n_when;
        if (2 * n.n_when + 1 > QGPU_STACK - 2) throw_internal("CASE expression too large");
        if (is_const) fold(begin, res);
        return {res, is_const, disp + " ELSE " + e.disp + " END"};
      }
      case QGPU_IR_IS_NULL:
      case QGPU_IR_IS_NOT_NULL: {
        R c = emit(*n.children[0]);
        add_op(n.kind == QGPU_IR_IS_NULL ? OP_ISNULL : OP_ISNOTNULL);
        DType res = mk_type(QGPU_T_BOOL);
        if (c.is_const) fold(begin, res);
        return {res, c.is_const, std::string(n.kind == QGPU_IR_IS_NULL ? "IsNull(" : "IsNotNull(") + c.disp + ")"};
      }
      case QGPU_IR_NEGATIVE: {
        R c = emit(*n.children[0]);
        if (!(c.type.is_signed_int() || c.type.is_decimal() || c.type.is_float()))
          throw_arrow("Invalid arithmetic operation: -" + c.type.str());
        Op& o = add_op(OP_NEG);
        o.vclass = class_of(c.type);
        o.wbits = (uint8_t)int_bits(c.type);
        if (c.is_const) fold(begin, c.type);
        return {c.type, c.is_const, "- " + c.disp};
      }
      case QGPU_IR_LIKE: {
        R v = emit(*n.children[0]);
        R pt = emit(*n.children[1]);
        if (v.type.id != QGPU_T_UTF8 || pt.type.id != QGPU_T_UTF8)
          throw_arrow("Invalid argument error: Invalid string operation: " + v.type.str() + " LIKE " + pt.type.str());
        Op& o = add_op(OP_LIKE);
        o.sub = (uint8_t)(n.op ? 1 : 0);
        DType res = mk_type(QGPU_T_BOOL);
        if (v.is_const && pt.is_const) fold(begin, res);
        return {res, v.is_const && pt.is_const, v.disp + (n.op ? " NOT LIKE " : " LIKE ") + pt.disp};
      }
      case QGPU_IR_EXTRACT: {
        R c = emit(*n.children[0]);
        if (c.type.id != QGPU_T_DATE32 && c.type.id != QGPU_T_DATE64)
          throw_arrow("Compute error: EXTRACT does not support " + c.type.str());
        Op& o = add_op(OP_EXTRACT);
        o.sub = (uint8_t)n.op;
        o.from_id = (uint8_t)c.type.id;
        DType res = mk_type(QGPU_T_INT64);
        if (c.is_const) fold(begin, res);
        return {res, c.is_const, "EXTRACT"};
      }
      default: throw_internal("unsupported physical expression kind " + std::to_string(n.kind));
    }
  }
};
}  // namespace

std::shared_ptr<Compiled> compile_expr(const ExprNode& root, const Schema& input) {
  auto c = std::make_shared<Compiled>();
  memset(&c->prog, 0, sizeof(Program));
  Compiler comp(input, *c);
  Compiler::R r = comp.emit(root);
  c->result_type = r.type;
  c->prog.result_class = class_of(r.type);
  c->display = r.disp;
  if (root.kind == QGPU_IR_COLUMN) {
    c->is_column_ref = true;
    c->column_ref = root.col_index;
  }
  if (r.is_const) {
    c->is_const = true;
    // after folding the program is a single OP_CONST
    c->const_val = c->prog.consts[c->prog.ops[c->prog.n_ops - 1].arg];
  }
  c->prog.n_cols = (int)c->col_slots.size();
  return c;
}

Program bind_program(Ctx* ctx, Compiled& c, const View& v) {
  Program P = c.prog;
  if (!c.blob.empty() && !c.dev_blob) {
    c.dev_blob = ctx->alloc(c.blob.size());
    ctx->h2d(c.dev_blob->ptr, c.blob.data(), c.blob.size());
  }
  for (int i = 0; i < P.n_consts; ++i)
    if (c.const_is_str[i]) P.consts[i].lo = (uint64_t)((const char*)c.dev_blob->ptr + c.prog.consts[i].lo);
  c.bound_subquery_cols.clear();
  for (size_t s = 0; s < c.col_slots.size(); ++s) {
    int ci = c.col_slots[s];
    LazyCol sub;
    if (ci <= -1000) {  // SubQuery operand: execute the sub-plan now
      PlanNode& sp = *c.subqueries[(size_t)(-ci - 1000)];
      View sv = sp.execute();
      sv.resolve();
      if (sv.num_batches == 0 || sv.cols.empty()) throw_internal("SubQuery returned no batch");  // batches[0] panics in the reference
      if (sv.num_rows != v.num_rows)
        throw_arrow("Invalid argument error: Cannot perform a binary operation on arrays of different length (SubQuery returned " +
                    std::to_string(sv.num_rows) + " rows for " + std::to_string(v.num_rows) + " input rows)");
      sub = sv.cols[0];
      c.bound_subquery_cols.push_back(sub);
    } else if (ci < 0 || ci >= (int)v.cols.size()) {
      throw_internal("expression column index out of range");
    }
    const LazyCol& lc = ci <= -1000 ? sub : v.cols[ci];
    ColRef& r = P.cols[s];
    memset(&r, 0, sizeof(ColRef));
    if (!lc.base) {
      if (v.num_rows == 0) {  // a table with zero batches has no buffers at all
        r.phys = PH_NULL;
        continue;
      }
      throw_internal("column '" + (ci >= 0 ? v.schema.fields[ci].name : std::string("SubQuery")) + "' is referenced but was not uploaded to the GPU table");
    }
    const DCol& d = *lc.base;
    r.phys = d.phys;
    r.data = d.data ? d.data->ptr : nullptr;
    r.offsets = d.offsets ? (const int32_t*)d.offsets->ptr : nullptr;
    r.validity = d.validity ? (const uint32_t*)d.validity->ptr : nullptr;
    r.idx = lc.idx ? lc.idx->ptr() : nullptr;
  }
  return P;
}

}  // namespace qgpu
