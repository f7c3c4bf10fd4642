"""Scalar currency and operator enums exchanged by the physical operators.

Mirrors (names, variants, meaning) the reference's
  * `Operator`     -- qurious/src/datatypes/operator.rs:4-20   (13 binary operators)
  * `ScalarValue`  -- qurious/src/datatypes/scalar.rs:85-107   (typed nullable scalar)
  * `JoinType`     -- qurious/src/common/join_type.rs:4-11     (6 join types)
  * `JoinSide`     -- qurious/src/physical/plan/join/nest_loop_join.rs:23-27

Arrow `DataType`s are represented by `pyarrow.DataType` objects (the reference uses
arrow-rs `DataType`; both describe the same Arrow logical types).
"""
from __future__ import annotations

import enum
import struct
from dataclasses import dataclass
from typing import Any, Optional

import pyarrow as pa


class Operator(enum.IntEnum):
    """qurious/src/datatypes/operator.rs:4-20 (same order => same wire code)."""

    Eq = 0
    NotEq = 1
    Gt = 2
    GtEq = 3
    Lt = 4
    LtEq = 5
    And = 6
    Or = 7
    Add = 8
    Sub = 9
    Mul = 10
    Div = 11
    Mod = 12

    def __str__(self) -> str:  # operator.rs:22-39 Display
        return {
            0: "=", 1: "!=", 2: ">", 3: ">=", 4: "<", 5: "<=", 6: "AND", 7: "OR",
            8: "+", 9: "-", 10: "*", 11: "/", 12: "%",
        }[int(self)]


class JoinType(enum.IntEnum):
    """qurious/src/common/join_type.rs:4-11."""

    Left = 0
    Right = 1
    Inner = 2
    Full = 3
    LeftSemi = 4
    LeftAnti = 5


class JoinSide(enum.IntEnum):
    """qurious/src/physical/plan/join/nest_loop_join.rs:23-27."""

    Left = 0
    Right = 1


class AggregateOperator(enum.IntEnum):
    """qurious/src/logical/expr/aggregate.rs:56-62."""

    Sum = 0
    Min = 1
    Max = 2
    Avg = 3
    Count = 4


# ---- wire type ids shared with include/qgpu.h (enum qgpu_type_id) ----------------------------
T_NULL, T_BOOL, T_INT8, T_INT16, T_INT32, T_INT64 = 0, 1, 2, 3, 4, 5
T_UINT8, T_UINT16, T_UINT32, T_UINT64 = 6, 7, 8, 9
T_FLOAT32, T_FLOAT64, T_UTF8, T_DATE32, T_DATE64, T_DECIMAL128 = 10, 11, 12, 13, 14, 15
T_TIME32, T_TIME64 = 16, 17      # unit in the `scale` slot: 0 s, 1 ms, 2 us, 3 ns

_SIMPLE = {
    pa.null(): T_NULL, pa.bool_(): T_BOOL, pa.int8(): T_INT8, pa.int16(): T_INT16,
    pa.int32(): T_INT32, pa.int64(): T_INT64, pa.uint8(): T_UINT8, pa.uint16(): T_UINT16,
    pa.uint32(): T_UINT32, pa.uint64(): T_UINT64, pa.float32(): T_FLOAT32,
    pa.float64(): T_FLOAT64, pa.string(): T_UTF8, pa.date32(): T_DATE32, pa.date64(): T_DATE64,
}


def type_triple(dt: pa.DataType) -> tuple[int, int, int]:
    """(type_id, precision, scale) as carried in the expression IR and `qgpu_agg_desc`."""
    if pa.types.is_decimal128(dt):
        return (T_DECIMAL128, dt.precision, dt.scale)
    if dt in _SIMPLE:
        return (_SIMPLE[dt], 0, 0)
    if pa.types.is_time32(dt):
        return (T_TIME32, 0, {"s": 0, "ms": 1}[dt.unit])
    if pa.types.is_time64(dt):
        return (T_TIME64, 0, {"us": 2, "ns": 3}[dt.unit])
    raise TypeError(f"InternalError: data type {dt} is not supported by the GPU operators")


def encode_type(dt: pa.DataType) -> bytes:
    t, p, s = type_triple(dt)
    return struct.pack("<BBb", t, p, s)


@dataclass(frozen=True)
class ScalarValue:
    """Typed nullable scalar (qurious/src/datatypes/scalar.rs:85-107).

    `value` is a Python bool/int/float/str, `None` for a typed NULL.  Decimal128 values are
    the raw unscaled i128 (as in `ScalarValue::Decimal128(Option<i128>, u8, i8)`).
    """

    data_type: pa.DataType
    value: Optional[Any]

    # constructors named after the enum variants
    @staticmethod
    def Null() -> "ScalarValue":
        return ScalarValue(pa.null(), None)

    @staticmethod
    def Boolean(v: Optional[bool]) -> "ScalarValue":
        return ScalarValue(pa.bool_(), v)

    @staticmethod
    def Int64(v: Optional[int]) -> "ScalarValue":
        return ScalarValue(pa.int64(), v)

    @staticmethod
    def Int32(v: Optional[int]) -> "ScalarValue":
        return ScalarValue(pa.int32(), v)

    @staticmethod
    def Int16(v: Optional[int]) -> "ScalarValue":
        return ScalarValue(pa.int16(), v)

    @staticmethod
    def Int8(v: Optional[int]) -> "ScalarValue":
        return ScalarValue(pa.int8(), v)

    @staticmethod
    def UInt64(v: Optional[int]) -> "ScalarValue":
        return ScalarValue(pa.uint64(), v)

    @staticmethod
    def UInt32(v: Optional[int]) -> "ScalarValue":
        return ScalarValue(pa.uint32(), v)

    @staticmethod
    def UInt16(v: Optional[int]) -> "ScalarValue":
        return ScalarValue(pa.uint16(), v)

    @staticmethod
    def UInt8(v: Optional[int]) -> "ScalarValue":
        return ScalarValue(pa.uint8(), v)

    @staticmethod
    def Float64(v: Optional[float]) -> "ScalarValue":
        return ScalarValue(pa.float64(), v)

    @staticmethod
    def Float32(v: Optional[float]) -> "ScalarValue":
        return ScalarValue(pa.float32(), v)

    @staticmethod
    def Decimal128(v: Optional[int], precision: int, scale: int) -> "ScalarValue":
        return ScalarValue(pa.decimal128(precision, scale), v)

    @staticmethod
    def Utf8(v: Optional[str]) -> "ScalarValue":
        return ScalarValue(pa.string(), v)

    def __str__(self) -> str:
        return "NULL" if self.value is None else str(self.value)

    def encode(self) -> bytes:
        """Wire form used inside the expression IR (see include/qgpu.h, QGPU_IR_LITERAL)."""
        t, p, s = type_triple(self.data_type)
        head = struct.pack("<BBbB", t, p, s, 1 if self.value is None else 0)
        v = self.value
        if t == T_NULL or v is None:
            return head
        if t == T_BOOL:
            return head + struct.pack("<q", 1 if v else 0)
        if t in (T_INT8, T_INT16, T_INT32, T_INT64, T_DATE32, T_DATE64):
            return head + struct.pack("<q", int(v))
        if t in (T_UINT8, T_UINT16, T_UINT32, T_UINT64):
            return head + struct.pack("<Q", int(v))
        if t in (T_FLOAT32, T_FLOAT64):
            return head + struct.pack("<d", float(v))
        if t == T_DECIMAL128:
            return head + (int(v) & ((1 << 128) - 1)).to_bytes(16, "little")
        if t == T_UTF8:
            b = v.encode("utf-8") if isinstance(v, str) else bytes(v)
            return head + struct.pack("<I", len(b)) + b
        raise TypeError(f"unsupported scalar type {self.data_type}")
