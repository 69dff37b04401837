"""GPU parity: the CUDA path, called through the C ABI (libggb200.so) and the host mirror of the reference API,
against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): quantized blocks bit-exact; mul_mat rel-L2 <= 1e-3 for Q4_0/Q4_1/F16
weights and <= 1e-5 for F32 weights.  The kernels are in fact much closer (only the summation order
differs), so each test also asserts a tighter regression bound.
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

from gpu_util import rel_l2
from ggmlsharp_b200 import ggml, native as N
from oracle import pyoracle as orc

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")

TOL = {N.F32: 1e-5, N.F16: 1e-3, N.Q4_0: 1e-3, N.Q4_1: 1e-3}          # the contract
TIGHT = {N.F32: 2e-6, N.F16: 2e-6, N.Q4_0: 2e-6, N.Q4_1: 5e-6}        # what the GEMV path actually achieves
NAMES = {N.F32: "f32", N.F16: "f16", N.Q4_0: "q4_0", N.Q4_1: "q4_1"}


def _f32(h):
    return np.frombuffer(bytes.fromhex(h), dtype=np.float32)


def weights(rng, M, K, kind="weights"):
    if kind == "weights":
        return (rng.standard_normal((M, K)) * 0.02).astype(np.float32)
    return rng.uniform(-1, 1, (M, K)).astype(np.float32)


# ---------------------------------------------------------------- codecs: bit-exact

@pytest.mark.parametrize("t,key", [(N.Q4_0, "q4_0"), (N.Q4_1, "q4_1"), (N.Q8_0, "q8_0"), (N.Q8_1, "q8_1")])
def test_quantize_golden_kats(t, key):
    with open(os.path.join(G, "quant_kat.json")) as f:
        for b in json.load(f)["blocks"]:
            got = ggml.quantize_rows(t, _f32(b["x"])).tobytes().hex()
            assert got == b[key], (b["name"], key)


@pytest.mark.parametrize("t,key,src", [(N.Q4_0, "deq4_0", "q4_0"), (N.Q4_1, "deq4_1", "q4_1")])
def test_dequantize_golden_kats(t, key, src):
    with open(os.path.join(G, "quant_kat.json")) as f:
        for b in json.load(f)["blocks"]:
            q = np.frombuffer(bytes.fromhex(b[src]), dtype=np.uint8)
            assert ggml.dequantize_rows(t, q, 32).tobytes().hex() == b[key], b["name"]


def _nasty(rng, nrows, k):
    """Inputs that exercise every branch that decides a bit: ties, zero blocks, +-0, equal magnitudes, tiny/huge."""
    x = rng.standard_normal((nrows, k)).astype(np.float32)
    x[0, :32] = 0.0                                        # all-zero block -> d = -0.0f
    x[1, :32] = -0.0
    x[2, :32] = np.where(np.arange(32) % 2 == 0, 1.5, -1.5)            # |max| tie: first element wins
    x[3, :32] = np.where(np.arange(32) % 2 == 0, -1.5, 1.5)
    x[4, :32] = np.arange(32) - 16                                       # ties-to-even on .5
    x[5, :32] = (np.arange(32) - 16) * 0.5
    x[6, :32] = 1e-38 * rng.standard_normal(32)                          # subnormal scales
    x[7, :32] = 1e30 * rng.standard_normal(32)
    x[8, :32] = 3.25
    x[9, :32] = np.array([0.0, -0.0] * 16)                               # min is +0 or -0 by order
    x[10, :32] = np.array([-0.0, 0.0] * 16)
    return x


@pytest.mark.parametrize("t", [N.Q4_0, N.Q4_1, N.Q8_0, N.Q8_1])
def test_quantize_bit_exact_random_and_edge_blocks(t):
    rng = np.random.default_rng(11)
    for nrows, k in ((16, 32), (12, 96), (33, 4096), (7, 11008)):
        x = _nasty(rng, nrows, k) if nrows >= 11 else rng.standard_normal((nrows, k)).astype(np.float32)
        for scale in (1.0, 0.02):
            xs = (x * np.float32(scale)).astype(np.float32)
            got, want = ggml.quantize_rows(t, xs), orc.quantize_rows(t, xs)
            assert np.array_equal(got, want), (t, nrows, k, scale, int(np.argmax((got != want).any(1))))


@pytest.mark.parametrize("t", [N.Q4_0, N.Q4_1])
def test_quantize_bit_exact_cfg1_full_matrix(t):
    # BASELINE.json configs[1]: quantize_row_q4_0 over the 4096x4096 F32 source, bit-exact
    rng = np.random.default_rng(1001)
    W = weights(rng, 4096, 4096)
    got, want = ggml.quantize_rows(t, W), orc.quantize_rows(t, W)
    assert np.array_equal(got, want)
    # size-independent property: dequantize(quantize(W)) stays within half a step of W (one step at the Q4_0 clamp)
    back = ggml.dequantize_rows(t, got, 4096)
    assert np.array_equal(back, orc.dequantize_rows(t, want, 4096))
    blk = W.reshape(-1, 32)
    step = np.abs(blk).max(1) / 8 if t == N.Q4_0 else (blk.max(1) - blk.min(1)) / 30
    assert (np.abs(back.reshape(-1, 32) - blk).max(1) <= step * 1.001 + 1e-12).all()


def test_f16_cast_matches_half_rne():
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.standard_normal(4096).astype(np.float32) * s for s in (1e-7, 1e-4, 1, 3e4, 1e5)]).reshape(5, 4096)
    got = ggml.quantize_rows(N.F16, x).view(np.uint16)
    np.testing.assert_array_equal(got, orc.f32_to_f16(x))


def test_codec_edge_cases():
    assert ggml.quantize_rows(N.Q4_0, np.zeros((0, 32), np.float32)).shape == (0, 20)        # empty input
    L = N.lib()
    x = np.zeros((1, 48), np.float32)
    out = np.zeros(64, np.uint8)
    assert L.ggb_quantize_rows(N.Q4_0, x.ctypes.data, out.ctypes.data, 1, 48) == N.E_INVALID  # k % 32 (Ggml.cs:336)
    assert L.ggb_quantize_rows(5, x.ctypes.data, out.ctypes.data, 1, 32) == N.E_UNSUPPORTED    # Q4_3: removed upstream, `default` table entry (Ggml.cs:247)


# ---------------------------------------------------------------- mul_mat through the device-level C ABI

class Dev:
    """Tiny RAII helper over ggb_dev_alloc/upload/download.

    compute-sanitizer is closed on this GPU pool, so every buffer a test hands to the library is fenced instead: GUARD bytes of a known
    pattern in front of and behind it, verified when the helper is closed.  A kernel that writes outside dst, outside its workspace
    slice or past the end of an operand it should only read turns into a test failure (reads out of bounds stay undetected)."""
    GUARD = 4096
    PATTERN = 0xA5

    def __init__(self):
        self.ptrs = []          # (base pointer, payload bytes)

    def _alloc(self, nbytes):
        p = C.c_void_p()
        N.check(N.lib().ggb_dev_alloc(max(nbytes, 1) + 2 * self.GUARD, C.byref(p)))
        fence = np.full(self.GUARD, self.PATTERN, dtype=np.uint8)
        N.check(N.lib().ggb_dev_upload(p.value, fence.ctypes.data, self.GUARD))
        N.check(N.lib().ggb_dev_upload(p.value + self.GUARD + max(nbytes, 1), fence.ctypes.data, self.GUARD))
        self.ptrs.append((p, max(nbytes, 1)))
        return p.value + self.GUARD

    def put(self, arr):
        arr = np.ascontiguousarray(arr)
        q = self._alloc(arr.nbytes)
        if arr.nbytes:
            N.check(N.lib().ggb_dev_upload(q, arr.ctypes.data, arr.nbytes))
        return q

    def empty(self, nbytes):
        return self._alloc(nbytes)

    def get(self, ptr, shape, dtype=np.float32):
        out = np.zeros(shape, dtype=dtype)
        if out.nbytes:
            N.check(N.lib().ggb_dev_download(out.ctypes.data, ptr, out.nbytes))
        return out

    def check_fences(self):
        N.check(N.lib().ggb_stream_sync(None))
        got = np.zeros(self.GUARD, dtype=np.uint8)
        for p, nbytes in self.ptrs:
            for off, where in ((0, "in front of"), (self.GUARD + nbytes, "behind")):
                N.check(N.lib().ggb_dev_download(got.ctypes.data, p.value + off, self.GUARD))
                bad = np.flatnonzero(got != self.PATTERN)
                assert bad.size == 0, "out-of-bounds write %s a %d-byte buffer: %d bytes, first at offset %d" % (where, nbytes, bad.size, int(bad[0]) - (self.GUARD if off == 0 else 0))

    def close(self):
        try:
            if self.ptrs:
                self.check_fences()
        finally:
            for p, _ in self.ptrs:
                N.lib().ggb_dev_free(p)
            self.ptrs = []


def dev_mul_mat(t, wbytes, M, K, X, nb01=None):
    X = np.ascontiguousarray(X, dtype=np.float32)
    Nn = X.shape[0]
    d = Dev()
    try:
        mm = N.ggb_dev_mm()
        mm.type, mm.M, mm.K, mm.N = t, M, K, Nn
        mm.W, mm.nb01 = d.put(wbytes), nb01 or (N.TYPE_SIZE[t] * (K // N.BLCK_SIZE[t]))
        mm.X, mm.ldx_bytes = d.put(X), 4 * K
        mm.Y, mm.ldy_bytes = d.empty(4 * M * Nn), 4 * M
        wsb = N.lib().ggb_dev_workspace_bytes(C.byref(mm), 1)
        ws = d.empty(wsb)
        N.check(N.lib().ggb_dev_mul_mat_batch(C.byref(mm), 1, ws, wsb, None))
        N.check(N.lib().ggb_stream_sync(None))
        return d.get(mm.Y, (Nn, M))
    finally:
        d.close()


@pytest.mark.parametrize("t,key,wkey", [(N.F32, "f32", "W"), (N.F16, "f16", "W_f16"), (N.Q4_0, "q4_0", "W_q4_0"), (N.Q4_1, "q4_1", "W_q4_1")])
def test_mul_mat_small_golden(t, key, wkey):
    with open(os.path.join(G, "mul_mat_small.json")) as f:
        g = json.load(f)
    M, K, Nn = g["M"], g["K"], g["N"]
    wb = np.frombuffer(bytes.fromhex(g[wkey]), dtype=np.uint8)
    got = dev_mul_mat(t, wb, M, K, _f32(g["X"]).reshape(Nn, K))
    want = _f32(g[key]).reshape(Nn, M)
    assert rel_l2(got, want) <= TIGHT[t]


SHAPES = [
    (4096, 4096, 1),      # cfg 1 / cfg 0
    (11008, 4096, 1),     # cfg 2: w1/w3
    (4096, 11008, 1),     # cfg 2: w2 (K = 11008, rows chunked)
    (1, 32, 1), (3, 64, 1), (130, 96, 1),          # tiny / ragged: rows shorter than one copy
    (257, 4128, 1),       # 129 blocks per row: Q4_0 rows are not 16-byte multiples -> plain-load staging
    (512, 2048, 2), (300, 1024, 3), (64, 4096, 7), (96, 512, 8), (40, 256, 13), (128, 1024, 15),
    (64, 11008, 9), (48, 11008, 15),     # long rows x several tokens: eight activation columns no longer fit in shared memory -> narrower passes
]


@pytest.mark.parametrize("t", [N.F32, N.F16, N.Q4_0, N.Q4_1])
@pytest.mark.parametrize("M,K,Nn", SHAPES)
def test_mul_mat_vs_oracle(t, M, K, Nn):
    rng = np.random.default_rng(1000 + M + K + Nn)
    for kind, xs in (("weights", "normal"), ("uniform", "uniform")):
        W = weights(rng, M, K, kind)
        X = rng.standard_normal((Nn, K)).astype(np.float32) if xs == "normal" else rng.uniform(-1, 1, (Nn, K)).astype(np.float32)
        wb = orc.encode_weights(t, W)
        got = dev_mul_mat(t, wb, M, K, X)
        want = orc.mul_mat_2d(t, wb, M, K, X, nth=8)
        err = rel_l2(got, want)
        assert err <= TOL[t], (NAMES[t], M, K, Nn, err)
        assert err <= TIGHT[t], (NAMES[t], M, K, Nn, err)


@pytest.mark.parametrize("t", [N.F32, N.F16, N.Q4_0, N.Q4_1])
def test_mul_mat_padded_rows_and_zero_sizes(t):
    rng = np.random.default_rng(77)
    M, K = 50, 256
    W = weights(rng, M, K)
    wb = orc.encode_weights(t, W)                       # [M, rb]
    rb = wb.shape[1]
    padded = np.zeros((M, rb + 48), dtype=np.uint8)     # a view with nb01 > row bytes
    padded[:, :rb] = wb
    X = rng.standard_normal((2, K)).astype(np.float32)
    got = dev_mul_mat(t, padded, M, K, X, nb01=rb + 48)
    assert rel_l2(got, orc.mul_mat_2d(t, wb, M, K, X)) <= TIGHT[t]
    assert dev_mul_mat(t, wb[:0], 0, K, X).shape == (2, 0)          # empty weight
    assert dev_mul_mat(t, wb, M, K, X[:0]).shape == (0, M)          # no activations


def test_linearity_property_full_size_q4_0():
    # size-independent property at BASELINE's full size: W.(a*x) == a*(W.x) exactly for a power of two
    # (quantize_row_q8_0 scales d by a and keeps the same quants), and y(x1)+y(x2) ~ y(x1+x2) within Q8 noise
    rng = np.random.default_rng(3)
    M = K = 4096
    wb = orc.quantize_rows(orc.Q4_0, weights(rng, M, K))
    x = rng.standard_normal((1, K)).astype(np.float32)
    y1 = dev_mul_mat(N.Q4_0, wb, M, K, x)
    y4 = dev_mul_mat(N.Q4_0, wb, M, K, x * np.float32(4))
    np.testing.assert_array_equal(y4, y1 * np.float32(4))
    assert np.array_equal(dev_mul_mat(N.Q4_0, wb, M, K, x), y1)     # deterministic run to run


# ---------------------------------------------------------------- through the reference-shaped host API

def host_mul_mat(t, W, X, n_dims_x=None):
    M, K = W.shape
    X = np.atleast_2d(X)
    wb = orc.encode_weights(t, W)
    with ggml.Context(64 << 20 if wb.nbytes < (16 << 20) else wb.nbytes * 2 + (64 << 20)) as c:
        a = c.tensor_from(t, K, M, data=wb)
        b = c.tensor_from(N.F32, K, data=X) if (X.shape[0] == 1 and n_dims_x == 1) else c.tensor_from(N.F32, K, X.shape[0], data=X)
        y = c.mul_mat(a, b)
        g = c.build_forward(y)
        c.graph_compute(g)
        out = ggml.tensor_f32(y).reshape(X.shape[0], M).copy()
        assert y.contents.perf_runs == 1 and g.perf_runs == 1
    return out, wb


@pytest.mark.parametrize("t", [N.F32, N.F16, N.Q4_0, N.Q4_1])
def test_graph_compute_single_node(t):
    rng = np.random.default_rng(21)
    W, X = weights(rng, 320, 1024), rng.standard_normal((1, 1024)).astype(np.float32)
    got, wb = host_mul_mat(t, W, X, n_dims_x=1)
    assert rel_l2(got, orc.mul_mat_2d(t, wb, 320, 1024, X)) <= TIGHT[t]
    got5, wb = host_mul_mat(t, W, rng.standard_normal((5, 1024)).astype(np.float32))
    assert got5.shape == (5, 320)


def test_test3_style_f32_mul_mat():
    # Test3/Program.cs:22-57: F[NP=4096 rows][NF=256] from the LCG, x* = (+1 x128, -1 x128) is what L-BFGS must
    # converge to, so F.x* ~ l = (+1 x2048, -1 x2048).  Pins orientation, strides and the F32 dot through the API.
    NP, NF = 1 << 12, 1 << 8
    nxt, r = 0, np.zeros(NP * NF, dtype=np.float32)
    for n in range(NP * NF):                                        # xrand(), Test3/Program.cs:98-102, xsrand(0)
        nxt = (nxt * 214013 + 2531011) & 0xFFFFFFFFFFFFFFFF
        r[n] = (nxt >> 16) & 0x7FFF
    l = np.where(np.arange(NP) < NP // 2, 1.0, -1.0).astype(np.float32)
    i = np.arange(NF)[None, :]
    ind = np.where(((l[:, None] > 0) & (i < NF // 2)) | ((l[:, None] < 0) & (i >= NF // 2)), 1.0, 0.0).astype(np.float32)
    noise = (r.reshape(NP, NF) / np.float32(32767) - np.float32(0.5)) * np.float32(0.1)
    F = ((ind + noise) / np.float32(0.5 * NF)).astype(np.float32)
    xstar = np.where(np.arange(NF) < NF // 2, 1.0, -1.0).astype(np.float32)[None, :]
    got, wb = host_mul_mat(N.F32, F, xstar, n_dims_x=1)
    want = orc.mul_mat_2d(orc.F32, wb, NP, NF, xstar, nth=8)
    assert rel_l2(got, want) <= 1e-6
    assert np.abs(got[0] - l).max() < 2e-2            # Test3/Program.cs:82-88 tolerance 1e-2 on x, same order on F.x


def test_graph_levels_chain_and_fanout():
    # y1 = W1.x ; y2 = W2.y1 ; y3 = W3.y1  -> two dependency levels, y1 stays on the device
    rng = np.random.default_rng(31)
    K, M1, M2 = 512, 256, 96
    W1, W2, W3 = weights(rng, M1, K), weights(rng, M2, M1, "uniform"), weights(rng, M2, M1, "uniform")
    x = rng.standard_normal((1, K)).astype(np.float32)
    w1b, w2b, w3b = orc.encode_weights(N.Q4_0, W1), orc.encode_weights(N.F16, W2), orc.encode_weights(N.Q4_1, W3)
    with ggml.Context(32 << 20) as c:
        a1, a2, a3 = c.tensor_from(N.Q4_0, K, M1, data=w1b), c.tensor_from(N.F16, M1, M2, data=w2b), c.tensor_from(N.Q4_1, M1, M2, data=w3b)
        b = c.tensor_from(N.F32, K, data=x)
        y1 = c.mul_mat(a1, b)
        y2, y3 = c.mul_mat(a2, y1), c.mul_mat(a3, y1)
        g = c.build_forward(y2)
        N.host().ggml_build_forward_expand(C.byref(g), y3)
        assert g.n_nodes == 3
        c.graph_compute(g)
        g1, g2, g3 = (ggml.tensor_f32(t).reshape(1, -1).copy() for t in (y1, y2, y3))
    r1 = orc.mul_mat_2d(orc.Q4_0, w1b, M1, K, x)
    assert rel_l2(g1, r1) <= TIGHT[N.Q4_0]
    # feed the ORACLE's y1 to the oracle, the DEVICE's y1 to the device: compare like with like
    assert rel_l2(g2, orc.mul_mat_2d(orc.F16, w2b, M2, M1, g1)) <= 1e-5
    assert rel_l2(g3, orc.mul_mat_2d(orc.Q4_1, w3b, M2, M1, g1)) <= 1e-5


def test_batched_dims_ne02_ne03():
    # src0 [K, M, 2, 3] x src1 [K, N, 2, 3]: every (i02, i03) slice is an independent mul_mat (Ggml.cs:6142-6162)
    rng = np.random.default_rng(41)
    K, M, Nn = 128, 24, 2
    W = weights(rng, 6 * M, K).reshape(3, 2, M, K)
    X = rng.standard_normal((3, 2, Nn, K)).astype(np.float32)
    for t in (N.F32, N.Q4_0):
        wb = orc.encode_weights(t, W.reshape(-1, K))
        with ggml.Context(16 << 20) as c:
            a = c.tensor_from(t, K, M, 2, 3, data=wb)
            b = c.tensor_from(N.F32, K, Nn, 2, 3, data=X)
            y = c.mul_mat(a, b)
            g = c.build_forward(y)
            c.graph_compute(g)
            got = ggml.tensor_f32(y).copy()                     # [3, 2, Nn, M]
        rb = wb.shape[1]
        for i3 in range(3):
            for i2 in range(2):
                ws = wb.reshape(3, 2, M, rb)[i3, i2]
                assert rel_l2(got[i3, i2], orc.mul_mat_2d(t, ws, M, K, X[i3, i2])) <= TIGHT[t]


def test_weights_are_reread_by_default_like_the_reference():
    # The reference reads src0->data on every ggml_graph_compute (Ggml.cs:6139-6164); user code rewrites tensor->data directly.
    # Default (no opt-in): every compute uploads the weights again and sees the rewrite.
    rng = np.random.default_rng(51)
    M, K = 64, 256
    x = rng.standard_normal((1, K)).astype(np.float32)
    W1, W2 = weights(rng, M, K), weights(rng, M, K)
    with ggml.Context(8 << 20) as c:
        a = c.tensor_from(N.F32, K, M, data=W1)
        b = c.tensor_from(N.F32, K, data=x)
        y = c.mul_mat(a, b)
        g = c.build_forward(y)
        N.lib().ggb_reset_stats()
        c.graph_compute(g)
        c.graph_compute(g)
        s = N.stats()
        assert s.weight_uploads == 0 and s.weight_cache_hits == 0
        r1 = ggml.tensor_f32(y).reshape(1, M).copy()
        ggml.tensor_bytes(a)[:] = W2.view(np.uint8).ravel()     # user rewrites the weight through tensor->data
        c.graph_compute(g)
        r2 = ggml.tensor_f32(y).reshape(1, M).copy()
    assert rel_l2(r1, orc.mul_mat_2d(orc.F32, W1.view(np.uint8).reshape(M, -1), M, K, x)) <= TIGHT[N.F32]
    assert rel_l2(r2, orc.mul_mat_2d(orc.F32, W2.view(np.uint8).reshape(M, -1), M, K, x)) <= TIGHT[N.F32]


def test_weight_cache_is_opt_in_and_invalidated_by_range():
    rng = np.random.default_rng(52)
    M, K = 64, 256
    x = rng.standard_normal((1, K)).astype(np.float32)
    W1, W2, W3 = weights(rng, M, K), weights(rng, M, K), weights(rng, M, K)
    want = lambda W: orc.mul_mat_2d(orc.F32, W.view(np.uint8).reshape(M, -1), M, K, x)
    with ggml.Context(8 << 20) as c:
        N.check(N.host().ggml_host_set_weight_cache(c.ctx, 1))
        a = c.tensor_from(N.F32, K, M, data=W1)
        b = c.tensor_from(N.F32, K, data=x)
        y = c.mul_mat(a, b)
        g = c.build_forward(y)
        N.lib().ggb_reset_stats()
        c.graph_compute(g)
        c.graph_compute(g)
        s = N.stats()
        assert s.weight_uploads == 1 and s.weight_cache_hits == 1
        # behind the API's back: the resident copy is what the contract of the opt-in says -- until it is invalidated
        ggml.tensor_bytes(a)[:] = W2.view(np.uint8).ravel()
        pool = N.host().ggml_host_pool_of(c.ctx)
        view = N.host().ggml_view_tensor(c.ctx, a)                 # invalidating through a VIEW of the cached leaf drops it too (byte ranges)
        view.contents.data = a.contents.data + 4 * K * 8
        view.contents.ne[1] = 4
        N.check(N.lib().ggb_tensor_invalidate(pool, view))
        c.graph_compute(g)
        assert rel_l2(ggml.tensor_f32(y).reshape(1, M), want(W2)) <= TIGHT[N.F32]
        # a setter of the API invalidates by itself
        N.host().ggml_set_f32(a, 0.5)
        c.graph_compute(g)
        got = ggml.tensor_f32(y).reshape(1, M).copy()
        assert rel_l2(got, want(np.full((M, K), 0.5, np.float32))) <= TIGHT[N.F32]
        # parameters are never resident: ggml_opt rewrites them in place between computes (Ggml.cs:1734-1760)
        a.contents.is_param = 1
        ggml.tensor_bytes(a)[:] = W3.view(np.uint8).ravel()
        N.lib().ggb_reset_stats()
        c.graph_compute(g)
        ggml.tensor_bytes(a)[:] = W1.view(np.uint8).ravel()
        c.graph_compute(g)
        assert N.stats().weight_uploads == 0
        assert rel_l2(ggml.tensor_f32(y).reshape(1, M), want(W1)) <= TIGHT[N.F32]
        # switching the cache off drops the mirrors
        a.contents.is_param = 0
        N.check(N.host().ggml_host_set_weight_cache(c.ctx, 0))
        ggml.tensor_bytes(a)[:] = W2.view(np.uint8).ravel()
        c.graph_compute(g)
        assert rel_l2(ggml.tensor_f32(y).reshape(1, M), want(W2)) <= TIGHT[N.F32]


def test_cpu_nodes_interleaved_with_device_nodes_keep_node_order():
    """seam B runs the nodes it takes before the caller's loop runs the rest.  x2 = sqr_inplace(x) stays on the
    host (not an op of this path) and rewrites the leaf x; y = W . x names x itself.  The reference runs sqr first, so y = W . x^2."""
    rng = np.random.default_rng(53)
    M, K, Nn = 48, 128, 3
    W, X = weights(rng, M, K), rng.standard_normal((Nn, K)).astype(np.float32)
    with ggml.Context(8 << 20) as c:
        a = c.tensor_from(N.F32, K, M, data=W)
        b = c.tensor_from(N.F32, K, Nn, data=X)
        x2 = c.op("sqr_inplace", b)
        y = c.mul_mat(a, b)                     # reads b AFTER the in-place square in node order
        z = c.op("silu", y)
        g = N.ggml_cgraph()
        g.n_threads = 4
        N.host().ggml_build_forward_expand(C.byref(g), x2)
        N.host().ggml_build_forward_expand(C.byref(g), z)
        ggml.tensor_f32(y)[...] = 0.0
        # this mixed graph cannot run end to end here: the MUL_MAT is (correctly) left to the caller's loop, which this C++ stand-in
        # of the C# host does not implement -- what matters is that the device did NOT compute W . x with the unsquared x
        N.host().ggml_graph_compute(c.ctx, C.byref(g))
        assert N.host().ggml_host_last_status() == N.E_UNSUPPORTED
        np.testing.assert_array_equal(ggml.tensor_f32(b).reshape(Nn, K), X * X)      # the host loop ran the square
        assert not ggml.tensor_f32(y).any(), "the device ran the MUL_MAT ahead of the CPU node it depends on through memory"
    # the other order is independent work and the device takes the MUL_MAT: y = W . x, then the host squares x
    with ggml.Context(8 << 20) as c:
        a = c.tensor_from(N.F32, K, M, data=W)
        b = c.tensor_from(N.F32, K, Nn, data=X)
        y = c.mul_mat(a, b)
        x2 = c.op("sqr_inplace", b)
        g = N.ggml_cgraph()
        g.n_threads = 4
        N.host().ggml_build_forward_expand(C.byref(g), y)
        N.host().ggml_build_forward_expand(C.byref(g), x2)
        c.graph_compute(g)
        got = ggml.tensor_f32(y).reshape(Nn, M).copy()
        np.testing.assert_array_equal(ggml.tensor_f32(b).reshape(Nn, K), X * X)
    assert rel_l2(got, orc.mul_mat_2d(orc.F32, W.view(np.uint8).reshape(M, -1), M, K, X)) <= TIGHT[N.F32]
    # a host node that CONSUMES a device result gets it on the host: s = sqr(W . x)
    with ggml.Context(8 << 20) as c:
        a = c.tensor_from(N.F32, K, M, data=W)
        b = c.tensor_from(N.F32, K, Nn, data=X)
        y = c.mul_mat(a, b)
        sq = c.op("sqr", y)
        g = c.build_forward(sq)
        c.graph_compute(g)
        yy = ggml.tensor_f32(y).reshape(Nn, M).copy()
        np.testing.assert_array_equal(ggml.tensor_f32(sq).reshape(Nn, M), yy * yy)


def test_cpy_into_an_offset_view_of_a_cached_leaf():
    """a CPY into view(cache, offset > 0) must (a) drop the cached mirror of the whole cache -- by byte range, not by start
    pointer -- and (b) never let a MUL_MAT over the whole cache multiply a mix of stale host rows and the fresh device rows."""
    rng = np.random.default_rng(54)
    K, R, Nn = 64, 16, 2
    cache0 = weights(rng, R, K).astype(np.float16)
    fresh = weights(rng, 8, K)
    X = rng.standard_normal((Nn, K)).astype(np.float32)
    with ggml.Context(8 << 20) as c:
        N.check(N.host().ggml_host_set_weight_cache(c.ctx, 1))
        cache = c.tensor_from(N.F16, K, R, data=cache0)
        b = c.tensor_from(N.F32, K, Nn, data=X)
        # 1st compute: the whole cache is multiplied and becomes resident
        y0 = c.mul_mat(cache, b)
        c.graph_compute(c.build_forward(y0))
        assert rel_l2(ggml.tensor_f32(y0).reshape(Nn, R), orc.mul_mat_2d(orc.F16, cache0.view(np.uint8).reshape(R, -1), R, K, X)) <= TIGHT[N.F16]
        # 2nd compute: CPY (F32 -> F16) into rows [8, 16)
        src = c.tensor_from(N.F32, K, 8, data=fresh)
        v = N.host().ggml_view_tensor(c.ctx, cache)
        v.contents.ne[1] = 8
        v.contents.nb[2] = v.contents.nb[1] * 8
        v.contents.nb[3] = v.contents.nb[2]
        v.contents.data = cache.contents.data + 8 * K * 2
        cp = c.cpy(src, v)
        c.graph_compute(c.build_forward(cp))
        want_cache = cache0.copy()
        want_cache[8:] = fresh.astype(np.float16)
        np.testing.assert_array_equal(ggml.tensor_bytes(cache).view(np.float16).reshape(R, K), want_cache)
        # 3rd compute: the whole cache again -- the resident copy made in the 1st compute is stale and must be gone
        y1 = c.mul_mat(cache, b)
        c.graph_compute(c.build_forward(y1))
        assert rel_l2(ggml.tensor_f32(y1).reshape(Nn, R), orc.mul_mat_2d(orc.F16, want_cache.view(np.uint8).reshape(R, -1), R, K, X)) <= TIGHT[N.F16]


def test_cpy_quantizes_through_the_public_route():
    # ggml_cpy(f32 -> q4_0) is the only public route to quantize_row_q (Ggml.cs:4339-4363); then use the result as src0
    rng = np.random.default_rng(61)
    M, K = 48, 512
    W, x = weights(rng, M, K), rng.standard_normal((1, K)).astype(np.float32)
    for t in (N.Q4_0, N.Q4_1, N.F16):
        with ggml.Context(8 << 20) as c:
            src = c.tensor_from(N.F32, K, M, data=W)
            dstq = c.new_tensor(t, K, M)
            cp = c.cpy(src, dstq)
            b = c.tensor_from(N.F32, K, data=x)
            y = c.mul_mat(cp, b)
            g = c.build_forward(y)
            assert g.n_nodes == 2
            c.graph_compute(g)
            qbytes = ggml.tensor_bytes(dstq).copy()
            got = ggml.tensor_f32(y).reshape(1, M).copy()
        want_q = orc.encode_weights(t, W)
        assert np.array_equal(qbytes, want_q.ravel())           # bit-exact blocks in the user's tensor
        assert rel_l2(got, orc.mul_mat_2d(t, want_q, M, K, x)) <= TIGHT[t]


def test_error_behaviour_matches_reference_asserts():
    L = N.lib()
    buf = np.zeros(4 << 20, dtype=np.uint8)
    with ggml.Context(buf.nbytes, mem_buffer=buf) as c:
        pool = C.c_void_p()
        N.check(L.ggb_pool_adopt(buf.ctypes.data, buf.nbytes, C.byref(pool)))
        w = c.new_tensor(N.Q8_1, 64, 4)                          # vec_dot_q = null in quantize_fns[] (Ggml.cs:274-282): no dot exists
        x = c.new_tensor(N.F32, 64)
        y = c.mul_mat(w, x)
        assert L.ggb_mul_mat_node(pool, y) == N.E_UNSUPPORTED
        w2 = c.new_tensor(N.F32, 64, 4)
        y2 = c.mul_mat(w2, x)
        y2.contents.ne[0] = 5                                    # dst shape no longer matches (Ggml.cs:6031)
        assert L.ggb_mul_mat_node(pool, y2) == N.E_INVALID
        y2.contents.ne[0] = 4
        x.contents.nb[0] = 8                                     # permuted src1 (Ggml.cs:6023)
        assert L.ggb_mul_mat_node(pool, y2) == N.E_INVALID
        x.contents.nb[0] = 4
        assert L.ggb_mul_mat_node(pool, y2) == 0
        N.check(L.ggb_pool_free(pool))


def test_stats_count_launches():
    N.lib().ggb_reset_stats()
    rng = np.random.default_rng(71)
    W = weights(rng, 64, 128)
    dev_mul_mat(N.Q4_0, orc.encode_weights(N.Q4_0, W), 64, 128, rng.standard_normal((1, 128)).astype(np.float32))
    s = N.stats()
    assert s.kernel_launches == 1            # a small single-token level: the GEMV quantizes the activation row in its prologue


def test_gemv_that_quantizes_its_own_row_equals_the_staged_path():
    """Levels of up to 8 single-token nodes skip the staging launch: k_gemv_fast builds the Q8 blocks of src1 itself (ggb_act_q8.cuh,
    quantize_row_q8_0 Ggml.cs:1158-1196).  The split-phase calls always stage with k_act_batch, so phase 1 + phase 2 over the same
    nodes must leave the same bytes -- on inputs that exercise every branch that decides a bit -- and one call is one launch."""
    rng = np.random.default_rng(93)
    # (type, M, K, launches of the one-call form): rows that are not whole 16-byte-aligned units take the plain-load GEMV, which is staged
    specs = [(N.Q4_0, 4096, 4096, 1), (N.Q4_1, 300, 2048, 1), (N.Q4_2, 129, 1024, 1), (N.Q5_0, 48, 4096, 1), (N.Q5_1, 7, 64, 1), (N.Q8_0, 100, 512, 1),
             (N.Q4_0, 64, 11008, 1), (N.Q4_1, 40, 11008, 1), (N.Q8_0, 33, 8192, 1),      # rows of more than 32 units, staged in K-chunks
             (N.Q5_1, 7, 32, 2), (N.Q4_0, 33, 96, 2)]
    d = Dev()
    try:
        for t, M, K, launches in specs:
            wb = orc.encode_weights(t, weights(rng, M, K))
            xs = _nasty(rng, 11, K)
            xs[3, :] = 0.0
            mm = N.ggb_dev_mm()
            mm.type, mm.M, mm.K, mm.N = t, M, K, 1
            mm.W, mm.nb01 = d.put(wb), wb.shape[1]
            X = d.put(xs)
            mm.ldx_bytes = 4 * K
            mm.Y, mm.ldy_bytes = d.empty(4 * M), 4 * M
            wsb = N.lib().ggb_dev_workspace_bytes(C.byref(mm), 1)
            ws = d.empty(wsb)
            for r in range(xs.shape[0]):
                mm.X = X + 4 * K * r
                N.lib().ggb_reset_stats()
                N.check(N.lib().ggb_dev_mul_mat_batch(C.byref(mm), 1, ws, wsb, None))
                N.check(N.lib().ggb_stream_sync(None))
                assert N.stats().kernel_launches == launches, (t, M, K)
                fused = d.get(mm.Y, (1, M))
                N.check(N.lib().ggb_dev_mul_mat_batch_phase(C.byref(mm), 1, ws, wsb, None, 1))
                N.check(N.lib().ggb_dev_mul_mat_batch_phase(C.byref(mm), 1, ws, wsb, None, 2))
                N.check(N.lib().ggb_stream_sync(None))
                staged = d.get(mm.Y, (1, M))
                assert np.array_equal(fused.view(np.uint32), staged.view(np.uint32)), (t, M, K, r)
                want = orc.mul_mat_2d(t, wb, M, K, xs[r:r + 1])
                assert rel_l2(fused, want) <= 6e-6 or not np.isfinite(want).all(), (t, M, K, r, rel_l2(fused, want))
    finally:
        d.close()


def test_peer_exchange_degenerates_cleanly_on_one_rank():
    # world = 1: ggb_peer_push_barrier copies nothing and the flag barrier completes on its own flags (epoch from device memory)
    from ggmlsharp_b200 import rowsplit
    sym = rowsplit.SymmetricBuffer(4096, 0, 1, lambda h: [h])
    try:
        for _ in range(3):
            sym.push_barrier(None, 0, 1024, 2048, 2)
            sym.barrier(None)
        N.check(N.lib().ggb_stream_sync(None))
        flags = np.zeros(32, dtype=np.uint64)
        N.check(N.lib().ggb_dev_download(flags.ctypes.data, sym.base, 256))
        assert flags[0] == 3                     # the last writer was barrier() with host epoch 3; push_barrier used device epochs 1..3
    finally:
        sym.close()


def test_rows_of_any_length():
    """The reference loops over any ne00 (Ggml.cs:6139-6164, 6676-6699).  One activation row has to sit in shared memory next to the GEMV's
    weight stages; rows beyond that (24 576 F32 / 49 152 F16 / ~78 000 quantized elements) are multiplied in K segments whose partial sums
    are added in segment order (ggb_shim.cu: needs_k_segments)."""
    rng = np.random.default_rng(91)
    for t, K, Nn in ((N.F32, 24576, 1), (N.F16, 49152, 1), (N.Q4_0, 65536, 1),                      # the longest rows one pass takes
                     (N.F32, 24608, 1), (N.F32, 70000 - 70000 % 32, 3), (N.F16, 131072, 1), (N.Q4_0, 131072 + 4096, 2),
                     (N.Q4_1, 98304, 1), (N.Q5_0, 81920 + 32, 1), (N.Q8_0, 163840, 1)):
        M = 5
        W = weights(rng, M, K)
        x = rng.standard_normal((Nn, K)).astype(np.float32)
        wb = orc.encode_weights(t, W)
        err = rel_l2(dev_mul_mat(t, wb, M, K, x), orc.mul_mat_2d(t, wb, M, K, x))
        assert err <= (1e-5 if t == N.F32 else 6e-6), (t, K, Nn, err)
    # ... and through the public route, next to an ordinary node
    K = 100000 - 100000 % 32
    W, x = weights(rng, 7, K), rng.standard_normal((1, K)).astype(np.float32)
    wb = orc.encode_weights(N.Q4_0, W)
    with ggml.Context(64 << 20) as c:
        a = c.tensor_from(N.Q4_0, K, 7, data=wb)
        b = c.tensor_from(N.F32, K, data=x)
        y = c.mul_mat(a, b)
        c.graph_compute(c.build_forward(y))
        got = ggml.tensor_f32(y).reshape(1, 7).copy()
    assert rel_l2(got, orc.mul_mat_2d(N.Q4_0, wb, 7, K, x)) <= 6e-6


def test_split_phases_equal_one_call():
    """ggb_dev_mul_mat_batch_phase: staging (phase 1, or 3 = without waiting for the preceding kernel) and multiplying (phase 2) as two
    calls leave the bytes of the one-call form; prompt-sized nodes are refused (their staging and GEMM are one pipeline)."""
    rng = np.random.default_rng(92)
    specs = [(N.Q4_0, 300, 512), (N.Q4_1, 129, 512), (N.F16, 64, 256), (N.F32, 33, 128), (N.Q5_0, 48, 512), (N.Q4_0, 40, 4128)]
    d = Dev()
    try:
        mms = (N.ggb_dev_mm * len(specs))()
        for i, (t, M, K) in enumerate(specs):
            wb = orc.encode_weights(t, weights(rng, M, K))
            x = rng.standard_normal((2, K)).astype(np.float32)
            m = mms[i]
            m.type, m.M, m.K, m.N = t, M, K, 2
            m.W, m.nb01 = d.put(wb), wb.shape[1]
            m.X, m.ldx_bytes = d.put(x), 4 * K
            m.Y, m.ldy_bytes = d.empty(4 * M * 2), 4 * M
        wsb = N.lib().ggb_dev_workspace_bytes(mms, len(specs))
        ws = d.empty(wsb)
        N.check(N.lib().ggb_dev_mul_mat_batch(mms, len(specs), ws, wsb, None))
        N.check(N.lib().ggb_stream_sync(None))
        want = [d.get(mms[i].Y, (2, specs[i][1])) for i in range(len(specs))]
        zero = np.zeros(wsb, np.uint8)
        for stage_phase in (1, 3):
            N.check(N.lib().ggb_dev_upload(ws, zero.ctypes.data, wsb))
            for i, (t, M, K) in enumerate(specs):
                z = np.zeros((2, M), np.float32)
                N.check(N.lib().ggb_dev_upload(mms[i].Y, z.ctypes.data, z.nbytes))
            N.check(N.lib().ggb_dev_mul_mat_batch_phase(mms, len(specs), ws, wsb, None, stage_phase))
            N.check(N.lib().ggb_stream_sync(None))
            assert not d.get(mms[0].Y, (2, specs[0][1])).any()              # staging alone multiplies nothing
            N.check(N.lib().ggb_dev_mul_mat_batch_phase(mms, len(specs), ws, wsb, None, 2))
            N.check(N.lib().ggb_stream_sync(None))
            for i in range(len(specs)):
                assert np.array_equal(d.get(mms[i].Y, (2, specs[i][1])), want[i]), (stage_phase, i)
        mms[0].N = 16                                                       # a prompt-sized node
        assert N.lib().ggb_dev_mul_mat_batch_phase(mms, 1, ws, wsb, None, 1) == N.E_UNSUPPORTED
        assert N.lib().ggb_dev_mul_mat_batch_phase(mms, 1, ws, wsb, None, 7) == N.E_INVALID
    finally:
        d.close()
