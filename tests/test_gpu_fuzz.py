"""Randomised shapes through ggb_dev_mul_mat_batch against the oracle: every weight type, GEMV and tensor-core regimes, ragged
row counts, K that is / is not a whole number of units, K-chunked rows, several nodes per call (so the batch executor's grouping,
activation sharing and per-type launch sequences are exercised together).  Seeds are fixed: a failure reproduces."""
import ctypes as C

import numpy as np
import pytest

from gpu_util import rel_l2
from ggmlsharp_b200 import native as N
from oracle import pyoracle as orc
from test_gpu_parity import Dev, weights

pytestmark = pytest.mark.gpu
TYPES = [N.F32, N.F16, N.Q4_0, N.Q4_1, N.Q4_2, N.Q5_0, N.Q5_1, N.Q8_0]
KS = [32, 64, 96, 128, 160, 256, 288, 384, 512, 1024, 1056, 2048, 4096, 4128, 5120, 6144, 11008]
NS = [1, 1, 1, 2, 3, 5, 8, 9, 15, 16, 17, 31, 64, 130]


@pytest.mark.parametrize("seed", range(40))
def test_random_batches_vs_oracle(seed):
    rng = np.random.default_rng(7000 + seed)
    n_nodes = int(rng.integers(1, 6))
    share_x = bool(rng.integers(0, 2))
    K = int(rng.choice(KS))
    Nn = int(rng.choice(NS))
    d = Dev()
    try:
        mms = (N.ggb_dev_mm * n_nodes)()
        cases = []
        Xs = rng.standard_normal((Nn, K)).astype(np.float32)
        pXs = d.put(Xs)
        for i in range(n_nodes):
            t = int(rng.choice(TYPES))
            Ki = K if share_x else int(rng.choice(KS))
            M = int(rng.integers(1, 700)) if Ki <= 4128 else int(rng.integers(1, 200))
            X = Xs if (share_x or Ki == K and rng.integers(0, 2)) else (rng.uniform(-1, 1, (Nn, Ki)).astype(np.float32))
            W = weights(rng, M, Ki, "weights" if rng.integers(0, 2) else "uniform")
            wb = orc.encode_weights(t, W)
            mm = mms[i]
            mm.type, mm.M, mm.K, mm.N = t, M, Ki, Nn
            mm.W, mm.nb01 = d.put(wb), wb.shape[1]
            mm.X, mm.ldx_bytes = (pXs if X is Xs else d.put(X)), 4 * Ki
            mm.Y, mm.ldy_bytes = d.empty(4 * M * Nn), 4 * M
            cases.append((t, M, Ki, wb, X))
        wsb = N.lib().ggb_dev_workspace_bytes(mms, n_nodes)
        ws = d.empty(wsb)
        N.check(N.lib().ggb_dev_mul_mat_batch(mms, n_nodes, ws, wsb, None))
        N.check(N.lib().ggb_stream_sync(None))
        for i, (t, M, Ki, wb, X) in enumerate(cases):
            got = d.get(mms[i].Y, (Nn, M))
            want = orc.mul_mat_2d(t, wb, M, Ki, X, nth=8)
            err = rel_l2(got, want)
            if t == N.F32:
                tol = 1e-5                                        # the contract for F32 weights (never on tensor cores)
            elif Nn >= 16:
                tol = 1e-3                                        # fp16 operands on the tensor-core path
            else:
                tol = 6e-6                                        # only the summation order differs
            assert err <= tol, (seed, i, t, M, Ki, Nn, share_x, err)
    finally:
        d.close()


@pytest.mark.parametrize("seed", range(32))
def test_random_strides_and_offsets_vs_oracle(seed):
    """The same comparison with what a VIEW hands the kernels: weight rows padded (nb01 > row bytes, 16-byte multiples or not), a
    weight base that is not 16-byte aligned, padded activation and output rows.  Aligned cases take the TMA / bulk-copy kernels,
    the others the plain-load GEMV and the fp16-expansion fallback -- all must agree with the oracle on the dense operands."""
    rng = np.random.default_rng(9000 + seed)
    t = int(rng.choice(TYPES))
    K = int(rng.choice([32, 64, 128, 160, 256, 512, 1024, 2048, 4096, 4128]))
    M = int(rng.integers(1, 400))
    Nn = int(rng.choice([1, 1, 2, 7, 16, 24, 33]))
    W = weights(rng, M, K)
    X = rng.standard_normal((Nn, K)).astype(np.float32)
    wb = orc.encode_weights(t, W)                                  # [M, rb]
    rb = wb.shape[1]
    align = {N.F32: 4, N.F16: 2, N.Q4_0: 4, N.Q4_1: 4, N.Q4_2: 2, N.Q5_0: 2, N.Q5_1: 2, N.Q8_0: 4}[t]
    pad = int(rng.choice([0, 16, 48, align, 3 * align]))
    base = int(rng.choice([0, 0, 16, align]))
    xpad = int(rng.choice([0, 16, 4]))
    ypad = int(rng.choice([0, 64, 4]))
    wbuf = np.zeros(base + M * (rb + pad) + 64, dtype=np.uint8)
    wbuf[base:base + M * (rb + pad)].reshape(M, rb + pad)[:, :rb] = wb
    xbuf = np.zeros((Nn, K + xpad // 4), dtype=np.float32)
    xbuf[:, :K] = X
    d = Dev()
    try:
        mm = N.ggb_dev_mm()
        mm.type, mm.M, mm.K, mm.N = t, M, K, Nn
        mm.W, mm.nb01 = d.put(wbuf) + base, rb + pad
        mm.X, mm.ldx_bytes = d.put(xbuf), 4 * K + xpad
        ldy = 4 * M + ypad
        py = d.put(np.full(Nn * ldy // 4 + 16, np.float32(-77.0)))
        mm.Y, mm.ldy_bytes = py, ldy
        wsb = N.lib().ggb_dev_workspace_bytes(C.byref(mm), 1)
        ws = d.empty(wsb)
        N.check(N.lib().ggb_dev_mul_mat_batch(C.byref(mm), 1, ws, wsb, None))
        N.check(N.lib().ggb_stream_sync(None))
        raw = d.get(py, (Nn * ldy // 4 + 16,))
        got = raw[:Nn * ldy // 4].reshape(Nn, ldy // 4)
        want = orc.mul_mat_2d(t, wb, M, K, X, nth=4)
        tol = 1e-5 if t == N.F32 else 1e-3 if Nn >= 16 else 6e-6
        assert rel_l2(got[:, :M], want) <= tol, (seed, t, M, K, Nn, pad, base, xpad, ypad, rel_l2(got[:, :M], want))
        assert (got[:, M:] == -77.0).all() and (raw[Nn * ldy // 4:] == -77.0).all()      # nothing written outside dst
    finally:
        d.close()
