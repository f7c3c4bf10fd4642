"""The C-ABI boundary: libqgpu.so loads without a GPU and exports every entry point include/qgpu.h declares; the Python
binding lists exactly those symbols; nothing in the product package imports the oracle."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "qgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)            # comments mention entry points too
    return sorted(set(re.findall(r"\b(qgpu_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(os.path.join(ROOT, "qurious_b200", "libqgpu.so"))
    names = declared_symbols()
    assert len(names) > 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_python_binding_covers_the_header():
    from qurious_b200 import _lib
    assert sorted(_lib.SYMBOLS) == declared_symbols()


def test_product_package_never_touches_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "qurious_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "from oracle" not in text and "import oracle" not in text, f
