"""The neighbours of mul_mat in a Llama layer on the device, through the C ABI, against the CPU oracle.

Bars: ADD / MUL / SCALE / SILU / REPEAT / CONT(transpose) and add_q_f32 (quantized blocks) bit-exact; RMS_NORM within 1e-6
relative (its double-precision sum is a block reduction, not sequential -- in practice the floats are identical); a whole
FFN block (rms_norm -> mul -> mul_mat x2 -> silu -> mul -> mul_mat -> add -> scale) through ggml_graph_compute with every
node executed by the CUDA executor and every intermediate equal to the oracle's composition within the mul_mat tolerance."""
import ctypes as C

import numpy as np
import pytest

from gpu_util import rel_l2
from ggmlsharp_b200 import ggml, native as N
from oracle import pyoracle as orc
from test_gpu_parity import Dev, weights

pytestmark = pytest.mark.gpu


def _run(d, fn, out_ptr, shape, dtype=np.float32):
    N.check(fn())
    N.check(N.lib().ggb_stream_sync(None))
    return d.get(out_ptr, shape, dtype)


@pytest.mark.parametrize("n", [1, 3, 4, 1023, 4096, 11008 * 7 + 5])
def test_add_mul_scale_bit_exact(n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n).astype(np.float32) * np.float32(3.7)
    y = rng.standard_normal(n).astype(np.float32)
    if n >= 4:
        x[:4] = [0.0, -0.0, 1e-42, np.inf]
        y[:4] = [-0.0, 0.0, 3.0, 1.0]
    L, d = N.lib(), Dev()
    try:
        px, py, pz = d.put(x), d.put(y), d.empty(4 * n)
        assert np.array_equal(_run(d, lambda: L.ggb_dev_binary(N.OP_ADD, px, py, pz, n, None), pz, (n,)).view(np.uint32), orc.add_f32(x, y).view(np.uint32))
        assert np.array_equal(_run(d, lambda: L.ggb_dev_binary(N.OP_MUL, px, py, pz, n, None), pz, (n,)).view(np.uint32), orc.mul_f32(x, y).view(np.uint32))
        if n > 8:                                              # unaligned operands take the scalar kernel
            got = _run(d, lambda: L.ggb_dev_binary(N.OP_ADD, px + 4, py + 4, pz + 4, n - 1, None), pz + 4, (n - 1,))
            assert np.array_equal(got.view(np.uint32), orc.add_f32(x[1:], y[1:]).view(np.uint32))
        v = np.float32(0.4125)
        assert np.array_equal(_run(d, lambda: L.ggb_dev_scale(px, v, n, None), px, (n,)).view(np.uint32), orc.scale_f32(x, v).view(np.uint32))
        assert L.ggb_dev_binary(N.OP_SILU, px, py, pz, n, None) == N.E_UNSUPPORTED
    finally:
        d.close()


def test_silu_every_fp16_input_and_random_floats():
    # the reference computes silu through a 64 K-entry fp16 table (GGML_SILU_FP16): every table entry must match
    halves = np.arange(1 << 16, dtype=np.uint16).view(np.float16).astype(np.float32)
    rng = np.random.default_rng(17)
    x = np.concatenate([halves, rng.standard_normal(100000).astype(np.float32) * 4, np.float32([70000.0, -70000.0, 1e-30, -1e-30])])
    L, d = N.lib(), Dev()
    try:
        px, py = d.put(x), d.empty(x.nbytes)
        got = _run(d, lambda: L.ggb_dev_silu(px, py, x.size, None), py, x.shape)
        want = orc.silu_f32(x)
        nan = np.isnan(want)
        assert np.array_equal(np.isnan(got), nan)
        assert np.array_equal(got[~nan].view(np.uint32), want[~nan].view(np.uint32))
    finally:
        d.close()


@pytest.mark.parametrize("nrows,k", [(1, 4096), (7, 4096), (3, 100), (512, 4096), (2, 11008), (5, 1)])
def test_rms_norm(nrows, k):
    rng = np.random.default_rng(nrows * 31 + k)
    x = (rng.standard_normal((nrows, k)) * rng.uniform(0.01, 30, (nrows, 1))).astype(np.float32)
    x[0, : min(k, 3)] = 0.0
    L, d = N.lib(), Dev()
    try:
        px, py = d.put(x), d.empty(x.nbytes)
        got = _run(d, lambda: L.ggb_dev_rms_norm(px, k, py, k, nrows, k, None), py, x.shape)
        want = orc.rms_norm_f32(x)
        assert rel_l2(got, want) <= 1e-6
        ulp = np.abs(got.view(np.int32).astype(np.int64) - want.view(np.int32).astype(np.int64))
        assert ulp.max() <= 1, int(ulp.max())
        # size-independent property: every non-zero output row has mean square 1 (up to eps and rounding)
        ms = (got.astype(np.float64) ** 2).mean(axis=1)
        nz = (x.astype(np.float64) ** 2).mean(axis=1) > 1e-3
        assert np.allclose(ms[nz], 1.0, atol=1e-3)
    finally:
        d.close()


@pytest.mark.parametrize("nr0,nc0,rr,rc", [(1, 4096, 8, 1), (3, 8, 2, 3), (1, 1, 5, 7), (4, 100, 1, 1)])
def test_repeat_bit_exact(nr0, nc0, rr, rc):
    x = np.random.default_rng(nr0 + nc0).standard_normal((nr0, nc0)).astype(np.float32)
    nr, nc = nr0 * rr, nc0 * rc
    L, d = N.lib(), Dev()
    try:
        px, py = d.put(x), d.empty(4 * nr * nc)
        got = _run(d, lambda: L.ggb_dev_repeat(px, nc0, nc0, nr0, py, nc, nc, nr, None), py, (nr, nc))
        assert np.array_equal(got, orc.repeat_f32(x, nr, nc))
        if nc0 > 1:
            assert L.ggb_dev_repeat(px, nc0, nc0, nr0, py, nc, nc + 1, nr, None) == N.E_INVALID              # ggml_can_repeat
    finally:
        d.close()


@pytest.mark.parametrize("rows,cols", [(3, 64), (100, 37), (4096, 256), (256, 4096), (33, 33), (1, 50)])
def test_cont_of_transposed_view_bit_exact(rows, cols):
    # ggml_cont(ggml_transpose(x)) for a contiguous x [rows][cols]: view ne = [rows, cols], nb = [cols*4, 4]
    x = np.random.default_rng(rows * 7 + cols).standard_normal((rows, cols)).astype(np.float32)
    L, d = N.lib(), Dev()
    try:
        px, py = d.put(x), d.empty(x.nbytes)
        ne = (C.c_int64 * 4)(rows, cols, 1, 1)
        nb = (C.c_uint64 * 4)(cols * 4, 4, rows * cols * 4, rows * cols * 4)
        got = _run(d, lambda: L.ggb_dev_cont(px, ne, nb, py, None), py, (cols, rows))
        assert np.array_equal(got, np.ascontiguousarray(x.T))
        assert np.array_equal(got.ravel(), orc.dup_f32_strided(x, (rows, cols), (cols * 4, 4)))
        # a permuted 3-D view takes the generic kernel: ne = [2, rows, cols/2], element (i0,i1,i2) = x[i1][2*i2 + i0]
        if cols % 2 == 0:
            ne3 = (C.c_int64 * 4)(2, rows, cols // 2, 1)
            nb3 = (C.c_uint64 * 4)(4, cols * 4, 8, rows * cols * 4)
            got3 = _run(d, lambda: L.ggb_dev_cont(px, ne3, nb3, py, None), py, (cols // 2, rows, 2))
            assert np.array_equal(got3.ravel(), orc.dup_f32_strided(x, (2, rows, cols // 2), (4, cols * 4, 8)))
    finally:
        d.close()


@pytest.mark.parametrize("t", [N.Q4_0, N.Q4_1])
@pytest.mark.parametrize("nrows,k", [(4, 64), (33, 4096), (5, 11008)])
def test_add_q_f32_bit_exact(t, nrows, k):
    rng = np.random.default_rng(nrows + k + t)
    w = weights(rng, nrows, k)
    x = (rng.standard_normal((nrows, k)) * 0.01).astype(np.float32)
    q = orc.quantize_rows(t, w)
    # blocks the fast quantizer must hand to the exact path: the sum is all zeros / has a +-max tie / is constant
    deq = orc.dequantize_rows(t, q, k)
    x[0, :32] = -deq[0, :32]
    x[1, :32] = np.where(np.arange(32) % 2 == 0, 1.5, -1.5) - deq[1, :32]
    x[2, :32] = 0.25 - deq[2, :32]
    want = orc.add_q_f32(t, q, x)
    L, d = N.lib(), Dev()
    try:
        pq, px, po = d.put(q), d.put(x), d.empty(q.nbytes)
        got = _run(d, lambda: L.ggb_dev_add_q(t, pq, px, po, nrows, k, None), po, q.shape, np.uint8)
        assert np.array_equal(got, want), int(np.argmax((got != want).any(1)))
        got2 = _run(d, lambda: L.ggb_dev_add_q(t, pq, px, pq, nrows, k, None), pq, q.shape, np.uint8)     # ggml_add_inplace: dst aliases src0
        assert np.array_equal(got2, want)
        assert L.ggb_dev_add_q(N.F16, pq, px, po, nrows, k, None) == N.E_UNSUPPORTED
        assert L.ggb_dev_add_q(t, pq, px, po, nrows, k + 16, None) == N.E_INVALID
    finally:
        d.close()


def _ffn_oracle(x, nw, w1b, w3b, w2b, K, F, tq):
    cur = orc.mul_f32(orc.repeat_f32(nw[None, :], x.shape[0], K), orc.rms_norm_f32(x))
    h1 = orc.mul_mat_2d(tq, w1b, F, K, cur, nth=8)
    h3 = orc.mul_mat_2d(tq, w3b, F, K, cur, nth=8)
    act = orc.mul_f32(orc.silu_f32(h1), h3)
    out = orc.mul_mat_2d(tq, w2b, K, F, act, nth=8)
    return cur, h1, h3, act, orc.add_f32(out, x)


@pytest.mark.parametrize("tq,Nn", [(N.Q4_0, 1), (N.Q4_0, 24), (N.F16, 3)])
def test_ffn_block_stays_on_the_device(tq, Nn):
    # x -> rms_norm -> * norm weights -> w1, w3 -> silu(w1 x) * (w3 x) -> w2 -> + x -> scale: the FFN half of a Llama layer
    rng = np.random.default_rng(50 + Nn)
    K, F = 256, 512
    x = rng.standard_normal((Nn, K)).astype(np.float32)
    nw = rng.uniform(0.5, 1.5, K).astype(np.float32)
    w1b, w3b, w2b = (orc.encode_weights(tq, weights(rng, F, K)), orc.encode_weights(tq, weights(rng, F, K)), orc.encode_weights(tq, weights(rng, K, F)))
    with ggml.Context(64 << 20) as c:
        tx = c.tensor_from(N.F32, K, Nn, data=x)
        tnw = c.tensor_from(N.F32, K, data=nw)
        tw1, tw3, tw2 = c.tensor_from(tq, K, F, data=w1b), c.tensor_from(tq, K, F, data=w3b), c.tensor_from(tq, F, K, data=w2b)
        tv = c.tensor_from(N.F32, 1, data=np.float32([0.5]))
        norm = c.op("rms_norm", tx)
        cur = c.op("mul", c.op("repeat", tnw, norm), norm)
        h1, h3 = c.mul_mat(tw1, cur), c.mul_mat(tw3, cur)
        act = c.op("mul", c.op("silu", h1), h3)
        out = c.op("add", c.mul_mat(tw2, act), tx)
        fin = c.op("scale", out, tv)
        g = c.build_forward(fin)
        N.lib().ggb_reset_stats()
        c.graph_compute(g)                                      # raises if any node was left to the (absent) CPU loop
        assert N.stats().nodes_executed == g.n_nodes
        got = {k: ggml.tensor_f32(v).reshape(Nn, -1).copy() for k, v in (("cur", cur), ("h1", h1), ("h3", h3), ("act", act), ("fin", fin))}
    cur_w, h1_w, h3_w, act_w, out_w = _ffn_oracle(x, nw, w1b, w3b, w2b, K, F, tq)
    tol = 1e-3 if Nn >= 16 else 5e-6
    assert rel_l2(got["cur"], cur_w) <= 1e-6
    assert rel_l2(got["h1"], h1_w) <= tol and rel_l2(got["h3"], h3_w) <= tol
    assert rel_l2(got["act"], act_w) <= 5 * tol + 2e-3          # silu rounds its input to fp16: a 1-ulp input change can move a table step
    # `out` and `fin` share their bytes (ggml_scale returns a view): the host sees the scaled tensor
    assert rel_l2(got["fin"], orc.scale_f32(out_w, np.float32(0.5))) <= 5 * tol + 2e-3


def test_test3_backward_shape_cont_transpose_feeds_mul_mat():
    # MUL_MAT's backward issues mul_mat(cont(transpose(src0)), grad) (Ggml.cs:7453-7462): Test3's shapes, F32 end to end
    rng = np.random.default_rng(3)
    nrows, ncols = 256, 4096                                   # Test3/Program.cs:22-28: x [ncols? no: ne = (256, 4096)]
    W = rng.standard_normal((ncols, nrows)).astype(np.float32)  # ne = [nrows, ncols]: 4096 rows of 256
    gvec = rng.standard_normal((1, ncols)).astype(np.float32)
    with ggml.Context(64 << 20) as c:
        tw = c.tensor_from(N.F32, nrows, ncols, data=W)
        tg = c.tensor_from(N.F32, ncols, 1, data=gvec)
        wt = c.op("cont", c.op("transpose", tw))                # [ncols, nrows] contiguous: 256 rows of 4096
        y = c.mul_mat(wt, tg)
        g = c.build_forward(y)
        c.graph_compute(g)
        got_t = ggml.tensor_f32(wt).reshape(nrows, ncols).copy()
        got_y = ggml.tensor_f32(y).reshape(1, nrows).copy()
    assert np.array_equal(got_t, np.ascontiguousarray(W.T))
    want = orc.mul_mat_2d(orc.F32, orc.encode_weights(orc.F32, np.ascontiguousarray(W.T)), nrows, ncols, gvec)
    assert rel_l2(got_y, want) <= 2e-6


def test_inplace_ops_and_mul_mat_only_flag():
    rng = np.random.default_rng(8)
    K, Nn = 128, 4
    x = rng.standard_normal((Nn, K)).astype(np.float32)
    y = rng.standard_normal((Nn, K)).astype(np.float32)
    with ggml.Context(16 << 20) as c:
        tx, ty = c.tensor_from(N.F32, K, Nn, data=x), c.tensor_from(N.F32, K, Nn, data=y)
        s = c.op("silu_inplace", c.op("add_inplace", tx, ty))   # tx <- silu(tx + ty), in place on a leaf
        g = c.build_forward(s)
        c.graph_compute(g)
        assert np.array_equal(ggml.tensor_f32(tx).reshape(Nn, K), orc.silu_f32(orc.add_f32(x, y)))
        # the round-1 contract is still available: with GGB_GRAPH_MUL_MAT_ONLY no neighbour is taken
        done = (C.c_uint8 * g.n_nodes)()
        from ggmlsharp_b200.native import lib
        pool = _pool_of(c)
        if pool is not None:
            assert lib().ggb_graph_compute_mul_mats(pool, C.byref(g), N.GRAPH_MUL_MAT_ONLY, done) == 0
            assert not any(done)


def _pool_of(c):
    f = getattr(N.host(), "ggml_host_pool_of", None)
    if f is None:
        return None
    f.restype, f.argtypes = C.c_void_p, [C.POINTER(N.ggml_context)]
    return f(c.ctx)
