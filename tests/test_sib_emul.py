"""Sibling formats (Q4_2 / Q5_0 / Q5_1 / Q8_0): the PRODUCT's register arithmetic (ggmlsharp_b200/csrc/ggb_sib_math.cuh, the
header the CUDA codecs and the GEMV include) compiled for the host (tests/emul/sib_emul.cpp supplies the CUDA intrinsics) and
checked against the oracle on the CPU: codecs bit-exact, GEMV unit / group dots to float reassociation."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emul", "sib_emul.cpp")
HDR = os.path.join(HERE, "..", "ggmlsharp_b200", "csrc", "ggb_sib_math.cuh")
SO = os.path.join(HERE, "emul", "_build", "libsib_emul.so")
SIB = [orc.Q4_2, orc.Q5_0, orc.Q5_1, orc.Q8_0]
UNIT_K = {orc.Q4_2: 128, orc.Q5_0: 128, orc.Q5_1: 64, orc.Q8_0: 128}


@pytest.fixture(scope="module")
def emul():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-Wall",
                               "-Wno-unknown-pragmas", "-Wno-unused-function", "-o", SO, SRC])
    lib = C.CDLL(SO)
    lib.emul_quantize.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_long]
    lib.emul_dequantize.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_long]
    lib.emul_gemv.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_long, C.c_long, C.c_void_p, C.c_void_p]
    return lib


def _edge_rows():
    rng = np.random.default_rng(5)
    rows = [rng.standard_normal(32) * s for s in (1e-3, 0.02, 1.0, 50.0, 1e20, 1e-30, 3e-41, 1e-44)]
    rows += [np.zeros(32), np.full(32, 2.5), np.arange(32) - 16.0, (np.arange(32) - 16.0) * 0.5 + 0.25]
    rows += [np.array([3.0, -3.0] + [1.0] * 30), np.array([-3.0, 3.0] + [1.0] * 30), np.array([0.0] * 31 + [-7.75])]
    a = rng.standard_normal(32); a[3] = np.nan; rows.append(a)
    a = rng.standard_normal(32); a[7] = np.inf; rows.append(a)
    a = rng.standard_normal(32); a[9] = -np.inf; a[20] = np.nan; rows.append(a)
    rows.append(np.array([0.0, -0.0] * 16)); rows.append(np.array([-0.0, 0.0] * 16))
    rows.append(np.full(32, 65520.0 * 16))      # d overflows fp16 -> inf
    return np.array(rows, dtype=np.float32)


@pytest.mark.parametrize("t", SIB)
def test_emulated_device_quantizer_bit_exact(emul, t):
    rng = np.random.default_rng(t)
    with np.errstate(all="ignore"):
        x = np.concatenate([_edge_rows(), (rng.standard_normal((4096, 32)) * rng.choice([1e-3, 0.02, 1, 30], (4096, 1))).astype(np.float32),
                            rng.uniform(-1, 1, (2048, 32)).astype(np.float32)])
    want = orc.quantize_rows(t, x)
    got = np.zeros_like(want)
    assert emul.emul_quantize(t, x.ctypes.data, got.ctypes.data, x.shape[0]) == 0
    bad = np.nonzero((got != want).any(axis=1))[0]
    assert bad.size == 0, (bad[:5], got[bad[0]], want[bad[0]], x[bad[0]])


@pytest.mark.parametrize("t", SIB)
def test_emulated_device_dequantizer_bit_exact(emul, t):
    rng = np.random.default_rng(100 + t)
    # every byte pattern is a legal block except NaN scales (compared as bit patterns below, so they may stay)
    q = rng.integers(0, 256, (8192, orc.row_bytes(t, 32)), dtype=np.uint8)
    q = np.concatenate([q, orc.quantize_rows(t, _edge_rows())])
    with np.errstate(all="ignore"):
        want = orc.dequantize_rows(t, q, 32)
    got = np.zeros_like(want)
    assert emul.emul_dequantize(t, q.ctypes.data, got.ctypes.data, q.shape[0]) == 0
    wn, gn = np.isnan(want), np.isnan(got)
    np.testing.assert_array_equal(wn, gn)
    np.testing.assert_array_equal(got.view(np.uint32)[~wn], want.view(np.uint32)[~wn])


@pytest.mark.parametrize("t", SIB)
@pytest.mark.parametrize("units", [1, 0])
@pytest.mark.parametrize("K", [128, 256, 4096, 11008])
def test_emulated_gemv_dot_matches_oracle(emul, t, units, K):
    rng = np.random.default_rng(K + t)
    M = 7
    W = (rng.standard_normal((M, K)) * 0.02).astype(np.float32)
    W[0, :64] = 0                                # all-zero groups
    x = rng.standard_normal((1, K)).astype(np.float32)
    x[0, 32:64] = 0
    wq = orc.quantize_rows(t, W)
    xq = orc.quantize_rows(orc.Q8_0, x)
    want = orc.mul_mat_2d(t, wq, M, K, x)[0]
    got = np.zeros(M, dtype=np.float32)
    assert emul.emul_gemv(t, units, wq.ctypes.data, M, K, xq.ctypes.data, got.ctypes.data) == 0
    rel = np.linalg.norm(got - want) / np.linalg.norm(want)
    assert rel < 3e-6, rel
    # a single unit is one lane's sequential chain -- the reference's own summation order -> bit-exact
    # (except Q5_1, whose m * (s0 + s1) is evaluated as m * (d * sum): ggb_sib_math.cuh)
    if units and t != orc.Q5_1:
        w1 = orc.quantize_rows(t, W[:, :UNIT_K[t]])
        x1 = np.ascontiguousarray(x[:, :UNIT_K[t]])
        xq1 = orc.quantize_rows(orc.Q8_0, x1)
        g1 = np.zeros(M, dtype=np.float32)
        emul.emul_gemv(t, units, w1.ctypes.data, M, UNIT_K[t], xq1.ctypes.data, g1.ctypes.data)
        np.testing.assert_array_equal(g1, orc.mul_mat_2d(t, w1, M, UNIT_K[t], x1)[0])
