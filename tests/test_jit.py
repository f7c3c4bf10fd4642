"""Run-time specialisation of the fused scan-aggregate kernel (csrc/fused_jit.cu), host side: qgpu_jit_compile instantiates
the hand-written kernel source for a shape signature through NVRTC and returns the sm_100a CUBIN -- no GPU needed, so the
"does it build" check covers the run-time compiled path too.  (GPU side: tests/test_gpu_fused.py.)"""
import ctypes

import pytest

from qurious_b200 import _lib

WC_I8, WC_U8, WC_I16, WC_U16, WC_I32, WC_U32, WC_I64 = range(7)
FK_SUM, FK_MIN, FK_MAX, FK_SUMF = range(4)


def sig_factor(wc, plain):
    return wc | ((1 if plain else 0) << 3)


def sig_acc(kind, chain, unit, nf, *f):
    v = kind | (int(chain) << 2) | (int(unit) << 3) | (nf << 4)
    for i, x in enumerate(f):
        v |= x << (6 + 4 * i)
    return v


def sig_head(preds, keys, n_accs):
    v = len(preds) | (len(keys) << 28) | (n_accs << 43)
    for i, w in enumerate(preds):
        v |= w << (4 + 3 * i)
    for i, w in enumerate(keys):
        v |= w << (31 + 3 * i)
    return v


def accs(*a):
    s = [0, 0, 0]
    for k, x in enumerate(a):
        s[k // 3] |= x << (18 * (k % 3))
    return s


def compile_sig(sig, pack):
    lib = _lib.load_library()
    log = ctypes.create_string_buffer(1 << 16)
    n = lib.qgpu_jit_compile((ctypes.c_uint64 * 4)(*sig), pack, log, len(log))
    return n, log.value.decode(errors="replace")


def nvrtc_present():
    try:
        ctypes.CDLL("libnvrtc.so.12")
        return True
    except OSError:
        return False


@pytest.mark.skipif(not nvrtc_present(), reason="libnvrtc.so.12 is not installed on this host")
@pytest.mark.parametrize("name,sig,pack", [
    # TPC-H Q1's shape (fused.cu SIG_Q1), packed and unpacked
    ("q1/packed", [sig_head([WC_I32], [WC_U8, WC_U8], 5)] + accs(
        sig_acc(FK_SUM, False, True, 1, sig_factor(WC_I64, True)), sig_acc(FK_SUM, False, True, 1, sig_factor(WC_I64, True)),
        sig_acc(FK_SUM, True, False, 1, sig_factor(WC_I64, False)), sig_acc(FK_SUM, True, False, 1, sig_factor(WC_I64, False)),
        sig_acc(FK_SUM, False, True, 1, sig_factor(WC_I64, True))), 0x11),
    # MIN / MAX / Float64 SUM over three keys and two predicates: a shape with no ahead-of-time kernel
    ("min-max-sumf/3 keys", [sig_head([WC_I32, WC_I64], [WC_I32, WC_U8, WC_I16], 3)] + accs(
        sig_acc(FK_MIN, False, True, 1, sig_factor(WC_I64, True)), sig_acc(FK_MAX, False, True, 1, sig_factor(WC_I32, True)),
        sig_acc(FK_SUMF, False, True, 1, sig_factor(WC_I64, True))), 0),
    # ungrouped product of three factors, eight predicates
    ("no keys/8 preds", [sig_head([WC_I32] * 8, [], 1)] + accs(
        sig_acc(FK_SUM, False, False, 3, sig_factor(WC_I64, True), sig_factor(WC_I32, False), sig_factor(WC_U8, True))), 0),
])
def test_nvrtc_instantiates_the_kernel_for_a_shape(name, sig, pack):
    n, log = compile_sig(sig, pack)
    assert n > 50_000, f"{name}: no CUBIN ({n}); compiler log:\n{log[:2000]}"
    assert "error" not in log.lower()
