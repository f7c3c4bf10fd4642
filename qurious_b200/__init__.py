"""qurious_b200 -- B200-native physical operators for holicc/qurious (hot path only).

The package holds what the path needs and nothing else:
  csrc/            hand-written CUDA kernels for sm_100a + the C ABI (include/qgpu.h) -> libqgpu.so
  physical/        host-side mirror of qurious/src/physical/{expr,plan} (constructors + IR serialisation)
  datatypes.py     ScalarValue / Operator / JoinType (qurious/src/datatypes, common/join_type.rs)
  _lib.py          ctypes binding of the C ABI (stand-in for the Rust extern "C" shim)
  tpch.py          synthetic TPC-H-shaped tables + the physical plans of Q1/Q6/Q3
"""
from .datatypes import AggregateOperator, JoinSide, JoinType, Operator, ScalarValue  # noqa: F401
from ._lib import Context, QuriousError, default_context  # noqa: F401

__all__ = ["AggregateOperator", "JoinSide", "JoinType", "Operator", "ScalarValue", "Context", "QuriousError",
           "default_context"]
