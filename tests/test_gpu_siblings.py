"""GPU parity of the sibling weight formats: Q4_2, Q5_0, Q5_1 and Q8_0 as a weight type -- codecs bit-exact,
mul_mat (GEMV and tensor-core paths), F32 -> quantized CPY and add_q_f32, all through the C ABI / the host mirror, against the
CPU oracle (oracle/ggb_oracle.c; its sibling section is pinned by tests/golden/sibling_small.json)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from gpu_util import rel_l2
from ggmlsharp_b200 import ggml, native as N
from oracle import pyoracle as orc
from test_gpu_parity import dev_mul_mat, weights, _nasty, Dev

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
SIB = [N.Q4_2, N.Q5_0, N.Q5_1, N.Q8_0]
NAME = {N.Q4_2: "q4_2", N.Q5_0: "q5_0", N.Q5_1: "q5_1", N.Q8_0: "q8_0"}
TOL = 1e-3                       # the contract for quantized weights (BASELINE.json north_star)
TIGHT_GEMV = 5e-6                # only the summation order differs on the single-token path
TIGHT_GEMM = 7e-4                # fp16 operands on the tensor-core path (Q4_0 / Q4_1 measure 3.3e-4 / < 6e-4)


def _f32(h):
    return np.frombuffer(bytes.fromhex(h), dtype=np.float32)


@pytest.fixture(scope="module")
def sib():
    with open(os.path.join(G, "sibling_small.json")) as f:
        return json.load(f)


# ---------------------------------------------------------------- codecs: bit-exact

@pytest.mark.parametrize("t", SIB)
def test_sibling_codecs_golden_kats(sib, t):
    for b in sib["blocks"]:
        assert ggml.quantize_rows(t, _f32(b["x"])).tobytes().hex() == b[NAME[t]], (b["name"], NAME[t])
        q = np.frombuffer(bytes.fromhex(b[NAME[t]]), dtype=np.uint8)
        assert ggml.dequantize_rows(t, q, 32).tobytes().hex() == b["de" + NAME[t]], (b["name"], NAME[t])


@pytest.mark.parametrize("t", SIB)
def test_sibling_quantize_bit_exact_random_and_edge_blocks(t):
    rng = np.random.default_rng(12)
    for nrows, k in ((16, 32), (12, 96), (33, 4096), (7, 11008)):
        x = _nasty(rng, nrows, k) if nrows >= 11 else rng.standard_normal((nrows, k)).astype(np.float32)
        if nrows >= 12 and k >= 64:
            x[11, :32] = np.where(np.arange(32) == 5, np.nan, x[11, :32])       # NaN / inf: .NET cast semantics
            x[11, 32:64] = np.where(np.arange(32) == 9, np.inf, x[11, 32:64])
        for scale in (1.0, 0.02, 1e-4):
            with np.errstate(all="ignore"):
                xs = (x * np.float32(scale)).astype(np.float32)
            got, want = ggml.quantize_rows(t, xs), orc.quantize_rows(t, xs)
            assert np.array_equal(got, want), (NAME[t], nrows, k, scale, int(np.argmax((got != want).any(1))))


@pytest.mark.parametrize("t", SIB)
def test_sibling_dequantize_bit_exact_and_roundtrip_bound(t):
    rng = np.random.default_rng(13)
    W = weights(rng, 512, 4096)
    q = orc.quantize_rows(t, W)
    back = ggml.dequantize_rows(t, q, 4096)
    assert np.array_equal(back, orc.dequantize_rows(t, q, 4096))
    # size-independent property: |dequantize(quantize(W)) - W| stays within one quantization step per block
    nb = 16 if t == N.Q4_2 else 32
    blk, bk = W.reshape(-1, nb), back.reshape(-1, nb)
    step = {N.Q4_2: np.abs(blk).max(1) / 8, N.Q5_0: np.abs(blk).max(1) / 16, N.Q5_1: (blk.max(1) - blk.min(1)) / 31,
            N.Q8_0: np.abs(blk).max(1) / 127}[t]
    assert (np.abs(bk - blk).max(1) <= step * 1.01 + 1e-6).all()
    # arbitrary byte patterns decode identically too (every pattern except NaN scales is a legal block)
    raw = rng.integers(0, 256, (64, orc.row_bytes(t, 256)), dtype=np.uint8)
    with np.errstate(all="ignore"):
        want = orc.dequantize_rows(t, raw, 256)
    got = ggml.dequantize_rows(t, raw, 256)
    ok = ~np.isnan(want)
    np.testing.assert_array_equal(got.view(np.uint32)[ok], want.view(np.uint32)[ok])


def test_sibling_codec_edge_cases():
    assert ggml.quantize_rows(N.Q5_0, np.zeros((0, 32), np.float32)).shape == (0, 22)
    L = N.lib()
    x = np.zeros((1, 48), np.float32)
    out = np.zeros(64, np.uint8)
    assert L.ggb_quantize_rows(N.Q4_2, x.ctypes.data, out.ctypes.data, 1, 48) == N.E_INVALID       # whole 32-element groups only
    assert L.ggb_quantize_rows(5, x.ctypes.data, out.ctypes.data, 1, 32) == N.E_UNSUPPORTED        # Q4_3: `default` table entry (Ggml.cs:247)
    f = np.zeros((1, 32), np.float32)
    assert L.ggb_dequantize_rows(N.Q8_1, out.ctypes.data, f.ctypes.data, 1, 32) == N.E_UNSUPPORTED # dequantize_row_q = null (Ggml.cs:278)


# ---------------------------------------------------------------- mul_mat

@pytest.mark.parametrize("t", SIB)
def test_sibling_mul_mat_small_golden(sib, t):
    M, K, Nn = sib["M"], sib["K"], sib["N"]
    wb = np.frombuffer(bytes.fromhex(sib["W_" + NAME[t]]), dtype=np.uint8)
    got = dev_mul_mat(t, wb, M, K, _f32(sib["X"]).reshape(Nn, K))
    assert rel_l2(got, _f32(sib[NAME[t]]).reshape(Nn, M)) <= TIGHT_GEMV


SHAPES = [
    (4096, 4096, 1),      # cfg 1 shape: one K-row per copy, activations in registers (Q5_1, Q8_0: shared-memory activations / K-chunks)
    (11008, 4096, 1),     # cfg 2: w1/w3
    (4096, 11008, 1),     # cfg 2: w2 (rows chunked)
    (1, 32, 1), (3, 64, 1), (130, 96, 1), (77, 160, 1),     # rows that are not whole units -> plain-load staging
    (257, 4128, 1),
    (512, 2048, 2), (300, 1024, 3), (64, 4096, 7), (96, 512, 8), (40, 256, 13), (128, 1024, 15), (48, 11008, 15),
]


@pytest.mark.parametrize("t", SIB)
@pytest.mark.parametrize("M,K,Nn", SHAPES)
def test_sibling_mul_mat_vs_oracle(t, M, K, Nn):
    rng = np.random.default_rng(2000 + M + K + Nn)
    for kind, xs in (("weights", "normal"), ("uniform", "uniform")):
        W = weights(rng, M, K, kind)
        X = rng.standard_normal((Nn, K)).astype(np.float32) if xs == "normal" else rng.uniform(-1, 1, (Nn, K)).astype(np.float32)
        wb = orc.encode_weights(t, W)
        got = dev_mul_mat(t, wb, M, K, X)
        want = orc.mul_mat_2d(t, wb, M, K, X, nth=8)
        err = rel_l2(got, want)
        assert err <= TOL, (NAME[t], M, K, Nn, err)
        assert err <= TIGHT_GEMV, (NAME[t], M, K, Nn, err)


@pytest.mark.parametrize("t", SIB)
def test_sibling_mul_mat_padded_rows_and_zero_sizes(t):
    rng = np.random.default_rng(78)
    M, K = 50, 256
    wb = orc.encode_weights(t, weights(rng, M, K))
    rb = wb.shape[1]
    padded = np.zeros((M, rb + 48), dtype=np.uint8)     # nb01 > row bytes
    padded[:, :rb] = wb
    X = rng.standard_normal((2, K)).astype(np.float32)
    assert rel_l2(dev_mul_mat(t, padded, M, K, X, nb01=rb + 48), orc.mul_mat_2d(t, wb, M, K, X)) <= TIGHT_GEMV
    assert dev_mul_mat(t, wb[:0], 0, K, X).shape == (2, 0)
    assert dev_mul_mat(t, wb, M, K, X[:0]).shape == (0, M)


@pytest.mark.parametrize("t", SIB)
def test_sibling_linearity_and_determinism_full_size(t):
    # exact power-of-two linearity at the cfg 1 size: quantize_row_q8_0 scales d by 4 and keeps the quants
    rng = np.random.default_rng(4)
    M = K = 4096
    wb = orc.quantize_rows(t, weights(rng, M, K))
    x = rng.standard_normal((1, K)).astype(np.float32)
    y1 = dev_mul_mat(t, wb, M, K, x)
    np.testing.assert_array_equal(dev_mul_mat(t, wb, M, K, x * np.float32(4)), y1 * np.float32(4))
    assert np.array_equal(dev_mul_mat(t, wb, M, K, x), y1)


@pytest.mark.parametrize("t", SIB)
@pytest.mark.parametrize("M,K,Nn", [(256, 512, 16), (4096, 4096, 512), (1000, 1024, 100), (384, 11008, 33), (200, 160, 24)])
def test_sibling_batched_tensor_core_path(t, M, K, Nn):
    """N >= 16 on tcgen05: Q4_2 / Q5_1 dequantized in the kernel like Q4_0 / Q4_1 (K % 128 == 0); Q5_0 / Q8_0 -- and K = 160 for
    every type -- through the fp16 expansion of the weights and the F16 kernel."""
    rng = np.random.default_rng(3000 + M + Nn)
    W = weights(rng, M, K)
    X = rng.standard_normal((Nn, K)).astype(np.float32)
    wb = orc.encode_weights(t, W)
    N.lib().ggb_reset_stats()
    got = dev_mul_mat(t, wb, M, K, X)
    # the oracle on a row sample (the full 4096 x 512 reference dot takes minutes on the CPU)
    rows = np.unique(np.concatenate([np.arange(min(M, 8)), rng.integers(0, M, 56), [M - 1]]))
    want = orc.mul_mat_2d(t, wb[rows], len(rows), K, X, nth=8)
    err = rel_l2(got[:, rows], want)
    assert err <= TOL, (NAME[t], M, K, Nn, err)
    assert err <= TIGHT_GEMM, (NAME[t], M, K, Nn, err)
    # every output is written and finite, and the sample is representative of the whole
    assert np.isfinite(got).all()
    wd = orc.dequantize_rows(t, wb, K).astype(np.float64)
    assert rel_l2(got, X.astype(np.float64) @ wd.T) < 2e-2          # unquantized activations: Q8 noise only


def test_sibling_batch_of_mixed_nodes():
    """One ggb_dev_mul_mat_batch call with GEMV and tensor-core nodes of every sibling type next to Q4_0 / F16."""
    rng = np.random.default_rng(41)
    specs = [(N.Q4_2, 256, 512, 1), (N.Q5_0, 192, 1024, 1), (N.Q5_1, 128, 256, 4), (N.Q8_0, 320, 512, 1), (N.Q4_0, 256, 512, 1),
             (N.Q5_0, 256, 512, 32), (N.Q8_0, 128, 512, 48), (N.Q4_2, 384, 256, 16), (N.Q5_1, 256, 1024, 64), (N.F16, 256, 512, 32)]
    d = Dev()
    try:
        mms = (N.ggb_dev_mm * len(specs))()
        keep = []
        for i, (t, M, K, Nn) in enumerate(specs):
            wb = orc.encode_weights(t, weights(rng, M, K))
            X = rng.standard_normal((Nn, K)).astype(np.float32)
            keep.append((wb, X))
            mm = mms[i]
            mm.type, mm.M, mm.K, mm.N = t, M, K, Nn
            mm.W, mm.nb01 = d.put(wb), wb.shape[1]
            mm.X, mm.ldx_bytes = d.put(X), 4 * K
            mm.Y, mm.ldy_bytes = d.empty(4 * M * Nn), 4 * M
        wsb = N.lib().ggb_dev_workspace_bytes(mms, len(specs))
        ws = d.empty(wsb)
        N.check(N.lib().ggb_dev_mul_mat_batch(mms, len(specs), ws, wsb, None))
        N.check(N.lib().ggb_stream_sync(None))
        for i, (t, M, K, Nn) in enumerate(specs):
            got = d.get(mms[i].Y, (Nn, M))
            want = orc.mul_mat_2d(t, keep[i][0], M, K, keep[i][1], nth=8)
            assert rel_l2(got, want) <= (TIGHT_GEMM if Nn >= 16 else TIGHT_GEMV), (i, specs[i], rel_l2(got, want))
    finally:
        d.close()


# ---------------------------------------------------------------- through the reference-shaped host API

@pytest.mark.parametrize("t", SIB)
def test_sibling_graph_cpy_then_mul_mat(t):
    """ggml_cpy(F32 -> sibling type) is the public route to quantize_row_q (Ggml.cs:4339-4363); the quantized tensor then feeds
    ggml_mul_mat in the same graph, so the weights never leave the device between the two nodes."""
    rng = np.random.default_rng(51)
    M, K = 192, 512
    W = weights(rng, M, K)
    x = rng.standard_normal((1, K)).astype(np.float32)
    with ggml.Context(32 << 20) as c:
        wf = c.tensor_from(N.F32, K, M, data=W)
        wq = c.new_tensor(t, K, M)
        cp = c.cpy(wf, wq)
        b = c.tensor_from(N.F32, K, data=x)
        y = c.mul_mat(cp, b)
        g = c.build_forward(y)
        c.graph_compute(g)
        qbytes = ggml.tensor_bytes(wq).reshape(M, -1).copy()
        got = ggml.tensor_f32(y).reshape(1, M).copy()
    want_q = orc.quantize_rows(t, W)
    assert np.array_equal(qbytes, want_q)
    assert rel_l2(got, orc.mul_mat_2d(t, want_q, M, K, x)) <= TIGHT_GEMV


@pytest.mark.parametrize("t", SIB)
@pytest.mark.parametrize("inplace", [False, True])
def test_sibling_add_q_f32(t, inplace):
    rng = np.random.default_rng(61)
    R, K = 37, 256
    q = orc.quantize_rows(t, weights(rng, R, K))
    x = (rng.standard_normal((R, K)) * 0.01).astype(np.float32)
    x[0, :32] = 0
    d = Dev()
    try:
        pq, px = d.put(q), d.put(x)
        pd = pq if inplace else d.empty(q.nbytes)
        N.check(N.lib().ggb_dev_add_q(t, pq, px, pd, R, K, None))
        N.check(N.lib().ggb_stream_sync(None))
        got = d.get(pd, q.shape, np.uint8)
    finally:
        d.close()
    assert np.array_equal(got, orc.add_q_f32(t, q, x))


@pytest.mark.parametrize("t", SIB)
def test_sibling_quantize_into_a_2_byte_aligned_destination(t):
    """10- and 22-byte blocks make 2-byte aligned destinations legal (a row slice of a Q5_0 tensor); those take the
    one-thread-per-group kernel instead of the tiled one, and must give the same bytes.  Also a ragged tile count."""
    rng = np.random.default_rng(71)
    R, K = 45, 96                                    # 135 groups: four full tiles + 7
    x = _nasty(rng, R, K)
    want = orc.quantize_rows(t, x)
    d = Dev()
    try:
        px = d.put(x)
        pd = d.empty(want.nbytes + 16)
        for off in (0, 2):
            N.check(N.lib().ggb_dev_quantize_rows(t, px, pd + off, R, K, None))
            N.check(N.lib().ggb_stream_sync(None))
            got = d.get(pd, (want.nbytes + 16,), np.uint8)[off:off + want.nbytes].reshape(want.shape)
            assert np.array_equal(got, want), (NAME[t], off)
    finally:
        d.close()


@pytest.mark.parametrize("t", [N.Q4_0, N.Q4_2, N.Q5_0])
def test_symmetric_formats_requantize_to_themselves_full_size(t):
    """Size-independent property at the cfg 1 size, entirely on the device: quantize(dequantize(q)) == q for the symmetric formats
    (the dequantized block's largest magnitude is exactly -8d / -16d)."""
    rng = np.random.default_rng(23 + t)
    W = weights(rng, 4096, 4096)
    q = ggml.quantize_rows(t, W)
    again = ggml.quantize_rows(t, ggml.dequantize_rows(t, q, 4096))
    assert np.array_equal(q, again)
