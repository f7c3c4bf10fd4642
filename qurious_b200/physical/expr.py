"""Physical expressions: the host-side mirror of qurious/src/physical/expr/*.

The reference's expression structs have private fields and no visitor, so the
GPU path cannot introspect an `Arc<dyn PhysicalExpr>`: a planner-side sibling must build the GPU
expression from the LogicalExpr tree.  These classes are that sibling's output.  They carry the
same constructor arguments as the reference structs and serialise to the postfix IR that
`qgpu_expr_parse` (include/qgpu.h) consumes.  Evaluation happens on the GPU only; there is no
host evaluation code here.

  Column        qurious/src/physical/expr/column.rs:7-20
  Literal       qurious/src/physical/expr/literal.rs:8-17
  BinaryExpr    qurious/src/physical/expr/binary.rs:17-28
  CastExpr      qurious/src/physical/expr/cast.rs:20-30
  CaseExpr      qurious/src/physical/expr/case.rs:15-28
  IsNull        qurious/src/physical/expr/is_null.rs:14-23
  IsNotNull     qurious/src/physical/expr/is_not_null.rs:14-23
  Negative      qurious/src/physical/expr/negative.rs:11-19
  Sum/Min/Max/Avg/Count AggregateExpr   qurious/src/physical/expr/aggregate/*.rs
"""
from __future__ import annotations

import struct
from typing import List, Sequence, Tuple

import pyarrow as pa

from ..datatypes import AggregateOperator, Operator, ScalarValue, encode_type

# IR opcodes (include/qgpu.h: enum qgpu_ir_op)
IR_COLUMN, IR_LITERAL, IR_BINARY, IR_CAST, IR_CASE, IR_IS_NULL, IR_IS_NOT_NULL, IR_NEGATIVE, IR_LIKE, IR_EXTRACT, IR_SUBQUERY = (
    1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11,
)


class PhysicalExpr:
    """`trait PhysicalExpr` (qurious/src/physical/expr/mod.rs:33-35)."""

    def to_ir(self) -> bytes:  # postfix byte stream
        raise NotImplementedError


class Column(PhysicalExpr):
    def __init__(self, name: str, index: int):
        self.name = name
        self.index = index

    def to_ir(self) -> bytes:
        return struct.pack("<BI", IR_COLUMN, self.index)

    def __str__(self) -> str:  # column.rs:36-40
        return f"{self.name}({self.index})"


class Literal(PhysicalExpr):
    def __init__(self, value: ScalarValue):
        self.value = value

    def to_ir(self) -> bytes:
        return struct.pack("<B", IR_LITERAL) + self.value.encode()

    def __str__(self) -> str:
        return str(self.value)


class BinaryExpr(PhysicalExpr):
    def __init__(self, left: PhysicalExpr, op: Operator, right: PhysicalExpr):
        self.left = left
        self.op = Operator(op)
        self.right = right

    def to_ir(self) -> bytes:
        return self.left.to_ir() + self.right.to_ir() + struct.pack("<BB", IR_BINARY, int(self.op))

    def __str__(self) -> str:  # binary.rs:83-87
        return f"{self.left} {self.op} {self.right}"


class CastExpr(PhysicalExpr):
    def __init__(self, expr: PhysicalExpr, data_type: pa.DataType):
        self.expr = expr
        self.data_type = data_type

    def to_ir(self) -> bytes:
        return self.expr.to_ir() + struct.pack("<B", IR_CAST) + encode_type(self.data_type)

    def __str__(self) -> str:
        return f"CAST({self.expr} AS {self.data_type})"


class CaseExpr(PhysicalExpr):
    def __init__(self, when_then: Sequence[Tuple[PhysicalExpr, PhysicalExpr]], else_expr: PhysicalExpr):
        self.when_then = list(when_then)
        self.else_expr = else_expr

    def to_ir(self) -> bytes:
        out = b""
        for w, t in self.when_then:
            out += w.to_ir() + t.to_ir()
        out += self.else_expr.to_ir()
        return out + struct.pack("<BI", IR_CASE, len(self.when_then))

    def __str__(self) -> str:
        s = "CASE"
        for w, t in self.when_then:
            s += f" WHEN {w} THEN {t}"
        return s + f" ELSE {self.else_expr} END"


class IsNull(PhysicalExpr):
    def __init__(self, expr: PhysicalExpr):
        self.expr = expr

    def to_ir(self) -> bytes:
        return self.expr.to_ir() + struct.pack("<B", IR_IS_NULL)

    def __str__(self) -> str:
        return f"IsNull({self.expr})"


class IsNotNull(PhysicalExpr):
    def __init__(self, expr: PhysicalExpr):
        self.expr = expr

    def to_ir(self) -> bytes:
        return self.expr.to_ir() + struct.pack("<B", IR_IS_NOT_NULL)

    def __str__(self) -> str:
        return f"IsNotNull({self.expr})"


class Negative(PhysicalExpr):
    def __init__(self, expr: PhysicalExpr):
        self.expr = expr

    def to_ir(self) -> bytes:
        return self.expr.to_ir() + struct.pack("<B", IR_NEGATIVE)

    def __str__(self) -> str:
        return f"- {self.expr}"


class Like(PhysicalExpr):
    """`Like::new(negated, expr, pattern)` (physical/expr/like.rs:20-24): arrow `like` / `nlike`."""

    def __init__(self, negated: bool, expr: PhysicalExpr, pattern: PhysicalExpr):
        self.negated, self.expr, self.pattern = bool(negated), expr, pattern

    def to_ir(self) -> bytes:
        return self.expr.to_ir() + self.pattern.to_ir() + struct.pack("<BB", IR_LIKE, 1 if self.negated else 0)

    def __str__(self) -> str:  # like.rs:44-52
        return f"{self.expr} {'NOT LIKE' if self.negated else 'LIKE'} {self.pattern}"


class DatetimeExtract:
    """`functions::datetime::extract::DatetimeExtract` (functions/datetime/extract.rs:17-27): EXTRACT(part FROM date) -> Int64."""

    PARTS = {"year": 0, "month": 1, "day": 2}

    def name(self) -> str:
        return "EXTRACT"

    def return_type(self) -> pa.DataType:
        return pa.int64()


class Function(PhysicalExpr):
    """`Function::new(func, args)` (physical/expr/function.rs:11-20).  The one function on the path is EXTRACT: args =
    [Literal(Utf8 part), date expression] (planner/sql.rs EXTRACT -> function call); the part is folded into the IR."""

    def __init__(self, func, args: Sequence[PhysicalExpr]):
        self.func, self.args = func, list(args)

    def to_ir(self) -> bytes:
        if not isinstance(self.func, DatetimeExtract):
            raise QuriousErrorLazy(f"InternalError: function {self.func.name()} is not supported on the GPU path")
        if len(self.args) != 2:
            raise QuriousErrorLazy("InvalidArgumentError: EXTRACT requires 2 arguments")
        unit = self.args[0]
        if not isinstance(unit, Literal) or unit.value.value is None or not pa.types.is_string(unit.value.data_type):
            raise QuriousErrorLazy("InvalidArgumentError: First argument of `EXTRACT` must be non-null scalar Utf8")
        part = str(unit.value.value).lower()          # extract.rs:63: case-insensitive
        if part not in DatetimeExtract.PARTS:
            raise QuriousErrorLazy(f"InternalError: Date part '{part}' not supported")
        return self.args[1].to_ir() + struct.pack("<BB", IR_EXTRACT, DatetimeExtract.PARTS[part])

    def __str__(self) -> str:  # function.rs:35-39
        return self.func.name()


class SubQuery(PhysicalExpr):
    """`SubQuery { plan }` (physical/expr/subquery.rs:11-20): `evaluate` executes the sub-plan and returns the FIRST column of
    its first batch, whatever the input batch is -- usable where that array has as many rows as the input (a scalar subquery
    over a one-row relation; arrow's kernels reject operands of different lengths).  The planner leaves these in SELECT
    lists; scalar subqueries in predicates are rewritten into joins (scalar_subquery_to_join.rs)."""

    def __init__(self, plan):
        self.plan = plan

    def to_ir(self) -> bytes:
        from .. import _lib
        ctx = _lib.current_parse_context()
        if ctx is None:
            raise QuriousErrorLazy("InternalError: SubQuery expressions are serialised by Context.parse_expr")
        _, h, _ = self.plan._native_cached(ctx)       # the sub-plan's native handles live as long as the plan object
        return struct.pack("<BQ", IR_SUBQUERY, h.value)

    def __str__(self) -> str:  # subquery.rs:29-33
        return "SubQuery"


def QuriousErrorLazy(msg: str):
    from .._lib import QuriousError
    return QuriousError(1, msg)


# ---------------------------------------------------------------------------------------------
# aggregate expressions (qurious/src/physical/expr/aggregate/mod.rs:16-19 `trait AggregateExpr`)
# ---------------------------------------------------------------------------------------------
class AggregateExpr:
    op: AggregateOperator

    def expression(self) -> PhysicalExpr:
        return self.expr  # type: ignore[attr-defined]


class SumAggregateExpr(AggregateExpr):
    """sum.rs:14-24.  Result types UInt64/Int64/Float64/Decimal128 only (sum.rs:37-50)."""

    op = AggregateOperator.Sum

    def __init__(self, expr: PhysicalExpr, return_type: pa.DataType):
        self.expr = expr
        self.return_type = return_type

    def __str__(self) -> str:
        return f"SUM({self.expr})"


class MinAggregateExpr(AggregateExpr):
    """min.rs:30-40."""

    op = AggregateOperator.Min

    def __init__(self, expr: PhysicalExpr, return_type: pa.DataType):
        self.expr = expr
        self.return_type = return_type

    def __str__(self) -> str:
        return f"MIN({self.expr})"


class MaxAggregateExpr(AggregateExpr):
    """max.rs:30-40."""

    op = AggregateOperator.Max

    def __init__(self, expr: PhysicalExpr, return_type: pa.DataType):
        self.expr = expr
        self.return_type = return_type

    def __str__(self) -> str:
        return f"MAX({self.expr})"


class CountAggregateExpr(AggregateExpr):
    """count.rs:9-18.  Result is always Int64 (logical/expr/aggregate.rs:67)."""

    op = AggregateOperator.Count

    def __init__(self, expr: PhysicalExpr):
        self.expr = expr
        self.return_type = pa.int64()

    def __str__(self) -> str:
        return f"COUNT({self.expr})"


class AvgAggregateExpr(AggregateExpr):
    """avg.rs:16-30: (expr, expr_data_type, return_type)."""

    op = AggregateOperator.Avg

    def __init__(self, expr: PhysicalExpr, expr_data_type: pa.DataType, return_type: pa.DataType):
        self.expr = expr
        self.expr_data_type = expr_data_type
        self.return_type = return_type

    def __str__(self) -> str:
        return f"AVG({self.expr})"


def avg_return_type(expr_data_type: pa.DataType) -> pa.DataType:
    """logical/expr/aggregate.rs:73-90."""
    if pa.types.is_decimal128(expr_data_type):
        return pa.decimal128(min(38, expr_data_type.precision + 4), min(38, expr_data_type.scale + 4))
    if pa.types.is_integer(expr_data_type) or pa.types.is_floating(expr_data_type):
        return pa.float64()
    raise TypeError(f"InternalError: avg does not support {expr_data_type}")


__all__: List[str] = [
    "PhysicalExpr", "Column", "Literal", "BinaryExpr", "CastExpr", "CaseExpr", "IsNull", "IsNotNull",
    "Negative", "AggregateExpr", "SumAggregateExpr", "MinAggregateExpr", "MaxAggregateExpr",
    "CountAggregateExpr", "AvgAggregateExpr", "avg_return_type",
]
