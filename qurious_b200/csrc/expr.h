// Parsed / compiled physical expressions.
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "interp.cuh"
#include "qgpu_internal.h"

namespace qgpu {

struct PlanNode;

struct ExprNode {
  std::shared_ptr<PlanNode> subplan;  // QGPU_IR_SUBQUERY
  int kind = 0;  // qgpu_ir_op
  int col_index = -1;
  DType lit_type;
  bool lit_null = false;
  uint64_t lit_lo = 0, lit_hi = 0;
  std::string lit_str;
  int op = 0;  // Operator code for BINARY
  DType cast_type;
  int n_when = 0;
  std::vector<std::unique_ptr<ExprNode>> children;
};

VClass class_of(const DType& t);

// An expression compiled against one input schema.
struct Compiled {
  Program prog;                 // cols[] are filled by bind_program()
  std::vector<int> col_slots;   // program column slot -> input column index
  DType result_type;
  bool is_column_ref = false;   // a bare Column: can be aliased without evaluation
  int column_ref = -1;
  bool is_const = false;        // whole expression folded to a constant
  Val const_val;                // valid when is_const (string constants: lo = offset into blob)
  std::vector<char> blob;       // bytes of string literals
  std::vector<uint8_t> const_is_str;
  DBufP dev_blob;               // blob uploaded to the device (lazily)
  int deferred_err = 0;         // EvalErr raised while folding constants; raised if rows > 0
  // SubQuery operands (subquery.rs): column slot -(1000 + k) is the first column of subqueries[k]'s result, executed by
  // bind_program; bound_subquery_cols keeps the columns alive while the kernels read them
  std::vector<std::shared_ptr<PlanNode>> subqueries;
  std::vector<LazyCol> bound_subquery_cols;
  std::string display;
};

std::unique_ptr<ExprNode> parse_ir(const uint8_t* ir, size_t len);
std::shared_ptr<Compiled> compile_expr(const ExprNode& root, const Schema& input);
// Fill column references from the view and fix up string-constant pointers for the device.
Program bind_program(Ctx* ctx, Compiled& c, const View& v);
[[noreturn]] void throw_eval_error(int code);
// decimal type rules of arrow-rs numeric kernels (see oracle/qref.py decimal_result_type)
DType decimal_result_type(int op, const DType& l, const DType& r);

}  // namespace qgpu

struct qgpu_expr {
  qgpu::Ctx* ctx = nullptr;
  std::shared_ptr<qgpu::ExprNode> root;
};
