"""Batched (N >= 16) mul_mat on the tcgen05 tensor-core path vs the CPU oracle (BASELINE.json configs[3]).

Bar: rel-L2 <= 1e-3 for Q4_0 / Q4_1 / F16 weights.  Expected error: activations enter as fp16(d1*q) and Q4 weights
as fp16(d0*(q-8)) (two ~2^-12 roundings), products exact, fp32 accumulation -> ~3e-4..5e-4 for Q4, ~1e-6 for F16
(whose activations the reference itself rounds to Half)."""
import ctypes as C

import numpy as np
import pytest

from gpu_util import rel_l2
from ggmlsharp_b200 import native as N
from oracle import pyoracle as orc
from test_gpu_parity import dev_mul_mat, weights

pytestmark = pytest.mark.gpu

GEMM_TOL = {N.Q4_0: 1e-3, N.Q4_1: 1e-3, N.F16: 1e-3}
GEMM_TIGHT = {N.Q4_0: 5e-4, N.Q4_1: 6e-4, N.F16: 5e-5}

SHAPES = [
    (128, 128, 16), (128, 256, 17), (256, 512, 64), (200, 384, 48),     # ragged M, N
    (384, 1024, 128), (1024, 4096, 130), (4096, 4096, 512),             # cfg 3
    (11008, 4096, 32), (4096, 11008, 16),                               # Llama FFN shapes, prompt batch
]


@pytest.mark.parametrize("t", [N.Q4_0, N.Q4_1, N.F16])
@pytest.mark.parametrize("M,K,Nn", SHAPES)
def test_gemm_vs_oracle(t, M, K, Nn):
    rng = np.random.default_rng(3000 + M + K + Nn)
    W = weights(rng, M, K)
    X = rng.standard_normal((Nn, K)).astype(np.float32)
    wb = orc.encode_weights(t, W)
    N.lib().ggb_reset_stats()
    got = dev_mul_mat(t, wb, M, K, X)
    # (row exponents of quantized weights, when the caller does not pass W_rowexp) + activation kernel + ONE tcgen05 GEMM launch, not GEMV passes
    assert N.stats().kernel_launches == (2 if t == N.F16 else 3)
    want = orc.mul_mat_2d(t, wb, M, K, X, nth=16)
    err = rel_l2(got, want)
    print('GEMM_ERR type=%d M=%d K=%d N=%d rel_l2=%.3e' % (t, M, K, Nn, err))
    assert err <= GEMM_TOL[t], (t, M, K, Nn, err)
    assert err <= GEMM_TIGHT[t], (t, M, K, Nn, err)


@pytest.mark.parametrize("t", [N.Q4_0, N.F16])
def test_gemm_uniform_inputs_and_determinism(t):
    rng = np.random.default_rng(9)
    M, K, Nn = 512, 2048, 96
    W = weights(rng, M, K, "uniform")
    X = rng.uniform(-1, 1, (Nn, K)).astype(np.float32)
    wb = orc.encode_weights(t, W)
    a = dev_mul_mat(t, wb, M, K, X)
    b = dev_mul_mat(t, wb, M, K, X)
    assert np.array_equal(a, b)
    assert rel_l2(a, orc.mul_mat_2d(t, wb, M, K, X, nth=16)) <= GEMM_TOL[t]


def test_gemm_columns_match_gemv_columns():
    # the same activation row must give (nearly) the same answer through the N=1 GEMV path and inside a batch
    rng = np.random.default_rng(10)
    M, K, Nn = 256, 1024, 32
    W = weights(rng, M, K)
    X = rng.standard_normal((Nn, K)).astype(np.float32)
    wb = orc.encode_weights(N.Q4_0, W)
    batch = dev_mul_mat(N.Q4_0, wb, M, K, X)
    single = dev_mul_mat(N.Q4_0, wb, M, K, X[5:6])
    assert rel_l2(batch[5:6], single) <= 1e-3


def test_f32_weights_batched_stay_on_ffma_path():
    # F32 weights must hold 1e-5, which fp16/tf32 tensor-core operands cannot: the batch runs as GEMV column passes
    rng = np.random.default_rng(12)
    M, K, Nn = 96, 256, 20
    W = weights(rng, M, K)
    X = rng.standard_normal((Nn, K)).astype(np.float32)
    wb = orc.encode_weights(N.F32, W)
    got = dev_mul_mat(N.F32, wb, M, K, X)
    assert rel_l2(got, orc.mul_mat_2d(N.F32, wb, M, K, X)) <= 2e-6


def test_graph_compute_routes_prompt_batches_to_the_tensor_core_path():
    # through the reference-shaped API: ggml_mul_mat with N = 48 columns, two dependent levels (y2 = W2 . y1)
    import ctypes as C
    from ggmlsharp_b200 import ggml
    rng = np.random.default_rng(77)
    K, M1, M2, Nn = 512, 384, 256, 48
    W1, W2 = weights(rng, M1, K), weights(rng, M2, M1, "uniform")
    X = rng.standard_normal((Nn, K)).astype(np.float32)
    w1b, w2b = orc.encode_weights(N.Q4_0, W1), orc.encode_weights(N.F16, W2)
    with ggml.Context(32 << 20) as c:
        a1, a2 = c.tensor_from(N.Q4_0, K, M1, data=w1b), c.tensor_from(N.F16, M1, M2, data=w2b)
        b = c.tensor_from(N.F32, K, Nn, data=X)
        y1 = c.mul_mat(a1, b)
        y2 = c.mul_mat(a2, y1)
        g = c.build_forward(y2)
        N.lib().ggb_reset_stats()
        c.graph_compute(g)
        assert N.stats().kernel_launches == 6          # Q4_0: row exponents + activation + GEMM; F16: activation + GEMM; + 1 result-copy kernel (y1 goes by memcpy)
        g1, g2 = ggml.tensor_f32(y1).reshape(Nn, M1).copy(), ggml.tensor_f32(y2).reshape(Nn, M2).copy()
    assert rel_l2(g1, orc.mul_mat_2d(orc.Q4_0, w1b, M1, K, X, nth=8)) <= 1e-3
    assert rel_l2(g2, orc.mul_mat_2d(orc.F16, w2b, M2, M1, g1, nth=8)) <= 1e-4


def test_kernel_timing_brackets_only_the_mul_mat_kernels():
    L = N.lib()
    rng = np.random.default_rng(78)
    W = weights(rng, 256, 512)
    wb = orc.encode_weights(N.Q4_0, W)
    L.ggb_reset_stats()
    N.check(L.ggb_set_kernel_timing(1))
    try:
        dev_mul_mat(N.Q4_0, wb, 256, 512, rng.standard_normal((1, 512)).astype(np.float32))      # GEMV
        dev_mul_mat(N.Q4_0, wb, 256, 512, rng.standard_normal((32, 512)).astype(np.float32))     # GEMM
        s = N.stats()
    finally:
        N.check(L.ggb_set_kernel_timing(0))
    assert s.kernel_launches == 4 and s.timed_kernel_launches == 2 and 0.0 < s.timed_kernel_ms < 5.0      # GEMV (stages its own row) | row exponents + activations + GEMM


def dev_mul_mat_batch(nodes):
    """nodes: [(type, wbytes, M, K, X)] submitted as ONE ggb_dev_mul_mat_batch call; returns the list of dst arrays."""
    import ctypes as C
    from test_gpu_parity import Dev
    d = Dev()
    try:
        mms = (N.ggb_dev_mm * len(nodes))()
        for i, (t, wb, M, K, X) in enumerate(nodes):
            X = np.ascontiguousarray(X, dtype=np.float32)
            m = mms[i]
            m.type, m.M, m.K, m.N = t, M, K, X.shape[0]
            m.W, m.nb01 = d.put(wb), N.TYPE_SIZE[t] * (K // N.BLCK_SIZE[t])
            m.X, m.ldx_bytes = d.put(X), 4 * K
            m.Y, m.ldy_bytes = d.empty(4 * M * X.shape[0]), 4 * M
        wsb = N.lib().ggb_dev_workspace_bytes(mms, len(nodes))
        ws = d.empty(wsb)
        N.check(N.lib().ggb_dev_mul_mat_batch(mms, len(nodes), ws, wsb, None))
        N.check(N.lib().ggb_stream_sync(None))
        return [d.get(mms[i].Y, (nodes[i][4].shape[0], nodes[i][2])) for i in range(len(nodes))]
    finally:
        d.close()


# (M, K, N): ragged M and N, a single K step (K = 128), odd K-step counts (384, 640), more tiles than CTA pairs, long K
GROUP_SHAPES = [(300, 128, 17), (128, 384, 16), (520, 640, 130), (4096, 4096, 512), (1376, 4096, 40), (512, 11008, 24), (256, 256, 64)]


@pytest.mark.parametrize("t", [N.Q4_0, N.Q4_1, N.F16])
def test_grouped_gemm_batch_vs_oracle(t):
    # one call with several batched nodes -> the persistent grouped kernel (one activation launch + one GEMM launch):
    # every node must match the oracle, and match what the same node gives when submitted alone (per-node kernel)
    rng = np.random.default_rng(4100 + t)
    nodes = []
    for (M, K, Nn) in GROUP_SHAPES:
        wb = orc.encode_weights(t, weights(rng, M, K))
        nodes.append((t, wb, M, K, rng.standard_normal((Nn, K)).astype(np.float32)))
    N.lib().ggb_reset_stats()
    got = dev_mul_mat_batch(nodes)
    assert N.stats().kernel_launches == (2 if t == N.F16 else 3)      # (one row-exponent launch for all nodes) + one activation launch + one GEMM launch
    for (tt, wb, M, K, X), y in zip(nodes, got):
        want = orc.mul_mat_2d(tt, wb, M, K, X, nth=16)
        err = rel_l2(y, want)
        assert err <= GEMM_TIGHT[t], (t, M, K, X.shape[0], err)
        alone = dev_mul_mat(tt, wb, M, K, X)
        # same fp16 operands; only the fp32 accumulation order may differ (256-column tiles keep ONE accumulator over all K steps,
        # the per-node kernel adds an even-K and an odd-K one)
        assert rel_l2(y, alone) <= 5e-6, (t, M, K, X.shape[0])


def test_grouped_gemm_mixed_batch_and_rerun():
    # Q4_0 + Q4_1 + F16 batched nodes and single-token nodes in one call; run twice into the same buffers (ring phases of a
    # second launch start from scratch) and require identical bits
    rng = np.random.default_rng(4200)
    nodes = []
    for t, (M, K, Nn) in [(N.Q4_0, (512, 1024, 32)), (N.F16, (384, 512, 48)), (N.Q4_1, (640, 768, 20)), (N.Q4_0, (256, 1024, 1)),
                          (N.Q4_0, (768, 2048, 96)), (N.F16, (256, 2048, 16)), (N.Q4_1, (256, 128, 33)), (N.F32, (64, 256, 1))]:
        wb = orc.encode_weights(t, weights(rng, M, K))
        nodes.append((t, wb, M, K, rng.standard_normal((Nn, K)).astype(np.float32)))
    a = dev_mul_mat_batch(nodes)
    b = dev_mul_mat_batch(nodes)
    for (t, wb, M, K, X), ya, yb in zip(nodes, a, b):
        assert np.array_equal(ya, yb)
        tol = GEMM_TOL[t] if X.shape[0] >= 16 else 5e-6
        assert rel_l2(ya, orc.mul_mat_2d(t, wb, M, K, X, nth=8)) <= tol, (t, M, K, X.shape[0])


@pytest.mark.parametrize("Nn", [1, 5, 32])
def test_nodes_sharing_one_activation_tensor(Nn):
    """wq / wk / wv (and w1 / w3) of a layer multiply the same src1: the batch stages those activations once and every node reads
    the shared copy.  Mixed weight types and one node with its own activations, each checked against the oracle."""
    import ctypes as C
    from test_gpu_parity import Dev, weights
    rng = np.random.default_rng(500 + Nn)
    K = 512
    Xs = rng.standard_normal((Nn, K)).astype(np.float32)
    Xo = rng.standard_normal((Nn, K)).astype(np.float32)
    specs = [(N.Q4_0, 256, True), (N.Q4_0, 384, True), (N.Q4_1, 256, True), (N.F16, 128, True), (N.Q4_0, 256, False),
             (N.Q5_0, 256, True), (N.Q4_1, 192, True), (N.Q8_0, 128, True), (N.F16, 256, True), (N.F32, 64, True), (N.F32, 96, True)]
    d = Dev()
    try:
        pXs, pXo = d.put(Xs), d.put(Xo)
        mms = (N.ggb_dev_mm * len(specs))()
        wbs = []
        for i, (t, M, shared) in enumerate(specs):
            wb = orc.encode_weights(t, weights(rng, M, K))
            wbs.append(wb)
            mm = mms[i]
            mm.type, mm.M, mm.K, mm.N = t, M, K, Nn
            mm.W, mm.nb01 = d.put(wb), wb.shape[1]
            mm.X, mm.ldx_bytes = (pXs if shared else pXo), 4 * K
            mm.Y, mm.ldy_bytes = d.empty(4 * M * Nn), 4 * M
        wsb = N.lib().ggb_dev_workspace_bytes(mms, len(specs))
        ws = d.empty(wsb)
        for _ in range(2):                               # twice: the second run must not depend on leftovers of the first
            N.check(N.lib().ggb_dev_mul_mat_batch(mms, len(specs), ws, wsb, None))
        N.check(N.lib().ggb_stream_sync(None))
        for i, (t, M, shared) in enumerate(specs):
            got = d.get(mms[i].Y, (Nn, M))
            want = orc.mul_mat_2d(t, wbs[i], M, K, Xs if shared else Xo, nth=4)
            tol = 1e-3 if (Nn >= 16 and t != N.F32) else 5e-6
            assert rel_l2(got, want) <= tol, (i, specs[i], rel_l2(got, want))
    finally:
        d.close()


@pytest.mark.parametrize("t", [N.Q4_0, N.Q4_1])
def test_odd_k_batches_reach_the_tensor_cores_through_the_expansion(t):
    """K = 160 is not a whole number of 128-wide K steps, so the TMA kernels cannot take it; the batch then expands the weights to
    fp16 once and runs the F16 kernel instead of 24 GEMV column passes.  The fp16 operand rounding (error well above the GEMV
    path's 2e-6, well below the 1e-3 contract) shows which path ran."""
    rng = np.random.default_rng(900 + t)
    M, K, Nn = 200, 160, 24
    W = weights(rng, M, K)
    X = rng.standard_normal((Nn, K)).astype(np.float32)
    wb = orc.encode_weights(t, W)
    got = dev_mul_mat(t, wb, M, K, X)
    err = rel_l2(got, orc.mul_mat_2d(t, wb, M, K, X, nth=4))
    assert 1e-5 < err <= 1e-3, err


def test_two_weight_types_consuming_a_fresh_intermediate():
    """Regression for a launch-ordering race (found by tests/test_gpu_graph_fuzz.py as a rare wrong result): a graph level whose
    batched nodes have DIFFERENT weight types launches one activation kernel per type; the later ones used to be launched behind a
    GEMM that releases its dependents at start-up, without waiting themselves, and could read the previous level's output while it
    was still being written.  The input is rescaled by a power of two before every run -- every result then scales with it (Q8
    blocks keep their quants; Half conversion is exact except for the few values that become fp16 subnormals) -- so a consumer
    that read the intermediate of the PREVIOUS run (same device address, another scale: an error of order 1) cannot pass."""
    from ggmlsharp_b200 import ggml
    rng = np.random.default_rng(77)
    Nn, K0, M0, M1 = 40, 4096, 8192, 256             # few tokens: small activation grids (every CTA resident at once), a long producer
    x = rng.standard_normal((Nn, K0)).astype(np.float32)
    w0 = orc.encode_weights(N.Q4_0, weights(rng, M0, K0))
    types = (N.F16, N.Q4_1, N.Q8_0, N.Q5_0)
    ws = [orc.encode_weights(t, weights(rng, M1, M0)) for t in types]
    base = None
    with ggml.Context(256 << 20) as c:
        tx = c.tensor_from(N.F32, K0, Nn, data=x)
        n0 = c.mul_mat(c.tensor_from(N.Q4_0, K0, M0, data=w0), tx)
        outs = [c.mul_mat(c.tensor_from(t, M0, M1, data=w), n0) for t, w in zip(types, ws)]
        g = c.build_forward(outs[0])
        for o in outs[1:]:
            N.host().ggml_build_forward_expand(C.byref(g), o)
        for rep in range(24):
            scale = np.float32(2.0 ** ((rep * 3) % 5 - 2))                 # 1/4 .. 4, a different one every run
            ggml.tensor_f32(tx).reshape(Nn, K0)[...] = x * scale
            c.graph_compute(g)
            got = [ggml.tensor_f32(o).reshape(Nn, M1).copy() / scale for o in outs]
            if base is None:
                base = got
                y0 = ggml.tensor_f32(n0).reshape(Nn, M0).copy()
                for t, w, y in zip(types, ws, got):
                    assert rel_l2(y * scale, orc.mul_mat_2d(t, w, M1, M0, y0, nth=8)) <= 1e-3      # like with like: the device's own n0
            else:
                for k, (a, b) in enumerate(zip(base, got)):
                    assert rel_l2(b, a) <= 1e-5, (rep, k, float(scale), rel_l2(b, a))


# ---- fp16 operand range of the tensor-core path ----------------------------------------------------------
#
# The reference keeps block scales in float32 and multiplies d0 * d1 * sumi in float32 (Ggml.cs:1158, 1190-1196), so its result is
# exact under any power-of-two rescaling of x or W.  The MMA operands are fp16; without range handling a block scale under 6e-8
# becomes 0, values under 6.1e-5 go subnormal and anything above 65504 becomes inf.  Every operand row is therefore pre-scaled by an
# exact power of two (per activation row: exponent of the row's largest |x|; per weight row: exponent of the largest value a block
# can dequantize to) and the epilogue multiplies both back.  These tests hold the 1e-3 contract far outside O(1) data.

def scale_weights(t, wb, M, K, e):
    """Multiply a quantized weight matrix by 2^e exactly, by rewriting the float32 / fp16 block scales (and offsets) in place."""
    wb = np.ascontiguousarray(wb).reshape(M, -1).copy()
    f = np.float32(2.0 ** e)
    if t in (N.Q4_0, N.Q8_0):
        bs = 20 if t == N.Q4_0 else 36
        blk = wb.reshape(M, -1, bs)
        d = blk[:, :, 0:4].copy().view(np.float32)
        blk[:, :, 0:4] = (d * f).astype(np.float32).view(np.uint8)
    elif t == N.Q4_1:
        blk = wb.reshape(M, -1, 24)
        dm = blk[:, :, 0:8].copy().view(np.float32)
        blk[:, :, 0:8] = (dm * f).astype(np.float32).view(np.uint8)
    elif t == N.Q5_1:
        blk = wb.reshape(M, -1, 24)
        dm = blk[:, :, 0:4].copy().view(np.float16).astype(np.float32) * f
        assert np.all(np.isfinite(dm.astype(np.float16)))
        blk[:, :, 0:4] = dm.astype(np.float16).view(np.uint8)
    else:
        raise AssertionError(t)
    return wb.reshape(M, -1)


RANGE_TYPES = [N.Q4_0, N.Q4_1, N.Q5_1, N.Q8_0]


@pytest.mark.parametrize("t", RANGE_TYPES)
@pytest.mark.parametrize("ex,ew", [(20, 0), (-20, 0), (0, 12), (0, -12), (20, -12), (-20, 12), (30, 30), (-30, -30)])
def test_gemm_operand_range_power_of_two_rescaling(t, ex, ew):
    rng = np.random.default_rng(7000 + t)
    M, K, Nn = 256, 1024, 48
    if t == N.Q5_1:
        ew = max(-4, min(4, ew))                       # its block scales are fp16 in the reference format itself; stay inside fp16
    W = weights(rng, M, K)
    X = rng.standard_normal((Nn, K)).astype(np.float32)
    wb0 = orc.encode_weights(t, W)
    wb = scale_weights(t, wb0, M, K, ew)
    Xs = (X * np.float32(2.0 ** ex)).astype(np.float32)
    got = dev_mul_mat(t, wb, M, K, Xs)
    want = orc.mul_mat_2d(t, wb, M, K, Xs, nth=8)
    assert np.all(np.isfinite(got)), "fp16 operand overflow"
    err = rel_l2(got, want)
    print('RANGE type=%d ex=%d ew=%d rel_l2=%.3e' % (t, ex, ew, err))
    assert err <= 1e-3, (t, ex, ew, err)
    # ... and it is the SAME answer as on the unscaled data, times 2^(ex + ew): the pre-scaling is exact
    base = dev_mul_mat(t, wb0, M, K, X)
    assert rel_l2(got / np.float64(2.0 ** (ex + ew)), base) <= 2e-6, (t, ex, ew)


@pytest.mark.parametrize("t", [N.Q4_0, N.Q4_1])
def test_gemm_outlier_rows_and_scales_near_fp16_limits(t):
    """One 1e5 activation row and one 3e-7 row among N(0, 1) rows; weight rows whose block scales sit near fp16's largest (6e4) and
    smallest (6e-8) values among N(0, 0.02) rows.  Per-ROW exponents keep every row at full precision."""
    rng = np.random.default_rng(7100 + t)
    M, K, Nn = 384, 512, 128
    W = weights(rng, M, K)
    sw, sx = np.ones(M), np.ones(Nn)
    sw[7] = 4.0e6          # block scales ~ 3e4: d itself still fits fp16, (q - 8) * d does not
    sw[100] = 3.0e9        # scales ~ 2e7: far above fp16
    sw[200] = 2.0e-5       # scales ~ 1.5e-7: fp16 subnormal
    sw[300] = 1.0e-9       # scales ~ 8e-12: zero in fp16
    sx[3], sx[9], sx[20] = 1.0e5, 3.0e-7, 1.0e-12
    W = (W * sw[:, None]).astype(np.float32)
    X = (rng.standard_normal((Nn, K)) * sx[:, None]).astype(np.float32)
    wb = orc.encode_weights(t, W)
    got = dev_mul_mat(t, wb, M, K, X)
    want = orc.mul_mat_2d(t, wb, M, K, X, nth=8)
    assert np.all(np.isfinite(got))
    # take the known row factors out again, so that every element weighs the same in the norm (unnormalised, one element --
    # the 1e5 row times the 3e9 weight row -- would be the whole rel-L2)
    norm = sx[:, None] * sw[None, :]
    g, w = got.astype(np.float64) / norm, want.astype(np.float64) / norm
    assert rel_l2(g, w) <= 1e-3, rel_l2(g, w)
    for n in (3, 9, 20, 0):
        assert rel_l2(g[n], w[n]) <= 1e-3, ("activation row", n, rel_l2(g[n], w[n]))
    for m in (7, 100, 200, 300, 0):
        assert rel_l2(g[:, m], w[:, m]) <= 1e-3, ("weight row", m, rel_l2(g[:, m], w[:, m]))


def test_gemm_with_precomputed_row_exponents_matches_the_in_call_ones():
    """ggb_dev_weight_rowexp once while the weights are resident, then ggb_dev_mm.W_rowexp: same bits as computing them in the call,
    one launch fewer."""
    from test_gpu_parity import Dev
    rng = np.random.default_rng(7200)
    M, K, Nn = 512, 1024, 64
    W = weights(rng, M, K)
    W[5] *= 1.0e6
    X = rng.standard_normal((Nn, K)).astype(np.float32)
    wb = orc.encode_weights(N.Q4_0, W)
    ref = dev_mul_mat(N.Q4_0, wb, M, K, X)
    d = Dev()
    try:
        mm = N.ggb_dev_mm()
        mm.type, mm.M, mm.K, mm.N = N.Q4_0, M, K, Nn
        mm.W, mm.nb01 = d.put(wb), 20 * (K // 32)
        mm.X, mm.ldx_bytes = d.put(X), 4 * K
        mm.Y, mm.ldy_bytes = d.empty(4 * M * Nn), 4 * M
        rowexp = d.empty(4 * M)
        N.check(N.lib().ggb_dev_weight_rowexp(N.Q4_0, mm.W, mm.nb01, M, K, rowexp, None))
        e = d.get(rowexp, (M,), dtype=np.int32)
        # ilogb(8 * max|d|) - 13, from the bytes
        dmax = np.abs(np.ascontiguousarray(wb).reshape(M, -1, 20)[:, :, 0:4].copy().view(np.float32)).max(axis=(1, 2))
        assert np.array_equal(e, np.floor(np.log2(8.0 * dmax.astype(np.float64))).astype(np.int32) - 13)
        mm.W_rowexp = rowexp
        wsb = N.lib().ggb_dev_workspace_bytes(C.byref(mm), 1)
        ws = d.empty(wsb)
        N.lib().ggb_reset_stats()
        N.check(N.lib().ggb_dev_mul_mat_batch(C.byref(mm), 1, ws, wsb, None))
        N.check(N.lib().ggb_stream_sync(None))
        assert N.stats().kernel_launches == 2
        assert np.array_equal(d.get(mm.Y, (Nn, M)), ref)
    finally:
        d.close()


def test_expansion_fallback_honours_the_range_too():
    # K = 160 takes the fp16-expansion path (k_expand_f16 + the F16 kernel): same pre-scaling there
    rng = np.random.default_rng(7300)
    M, K, Nn = 200, 160, 24
    W = weights(rng, M, K) * np.float32(2.0 ** 24)
    X = (rng.standard_normal((Nn, K)) * 2.0 ** -22).astype(np.float32)
    wb = orc.encode_weights(N.Q4_0, W)
    got = dev_mul_mat(N.Q4_0, wb, M, K, X)
    assert np.all(np.isfinite(got))
    assert rel_l2(got, orc.mul_mat_2d(N.Q4_0, wb, M, K, X, nth=4)) <= 1e-3


def test_f16_weights_keep_the_reference_half_conversion():
    # F16 weights: the reference itself rounds src1 to Half (Ggml.cs:6362-6379), overflow to inf included -- no rescaling there
    rng = np.random.default_rng(7400)
    M, K, Nn = 128, 256, 16
    W = weights(rng, M, K)
    X = rng.standard_normal((Nn, K)).astype(np.float32)
    X[2, 5] = 1.0e6                                       # (Half)1e6 = +inf in the reference
    wb = orc.encode_weights(N.F16, W)
    got = dev_mul_mat(N.F16, wb, M, K, X)
    want = orc.mul_mat_2d(N.F16, wb, M, K, X, nth=4)
    assert np.array_equal(np.isfinite(got), np.isfinite(want))
    ok = np.isfinite(want)
    assert rel_l2(got[ok], want[ok]) <= 1e-4
