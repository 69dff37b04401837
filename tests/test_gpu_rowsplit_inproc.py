"""Row split INSIDE the shim: one process, the unchanged ggml_graph_compute, several GPUs.

The reference splits the rows of src0 over the OS threads of ggml_graph_compute (Ggml.cs:3231-3252, 6665-6672); here each "thread" is
a host thread driving one GPU: device g multiplies rows [ceil(M/G) g, ceil(M/G) (g+1)) and its kernels store the results into every
device's copy of dst (the all-gather fused into the epilogue), element-wise neighbours are replicated, device 0 returns the results.
A row's dot products do not depend on which device computes them, so the split results must equal the single-GPU ones BIT FOR BIT,
and both are held to the oracle.  Needs >= 2 devices; skipped otherwise (the driver's 1-GPU test box)."""
import ctypes as C

import numpy as np
import pytest

from gpu_util import rel_l2
from ggmlsharp_b200 import ggml, native as N
from oracle import pyoracle as orc

pytestmark = pytest.mark.gpu


def _ngpu():
    n = C.c_int(0)
    N.lib().ggb_device_count(C.byref(n))
    return n.value


def _need2():
    if _ngpu() < 2:
        pytest.skip("row split needs >= 2 GPUs")


def weights(rng, M, K):
    return (rng.standard_normal((M, K)) * 0.02).astype(np.float32)


def _layer_graph(c, rng, Nn, types, cache):
    """x -> rms_norm -> {w1, w3} -> silu(w1 x) * (w3 x) -> w2 -> + x -> a small head: the FFN block of a Llama layer plus nodes that
    are too small to split, all in one graph with four dependent mul_mat levels."""
    K, F, H = 512, 1408, 48                      # F = 1408: ceil(1408 / 2) = 704 rows each
    X = rng.standard_normal((Nn, K)).astype(np.float32)
    W1, W3, W2, Wh = weights(rng, F, K), weights(rng, F, K), weights(rng, K, F), weights(rng, H, K)
    enc = {k: orc.encode_weights(t, w) for k, t, w in (("w1", types[0], W1), ("w3", types[1], W3), ("w2", types[2], W2), ("wh", N.F32, Wh))}
    if cache:
        N.check(N.host().ggml_host_set_weight_cache(c.ctx, 1))
    x = c.tensor_from(N.F32, K, Nn, data=X)
    w1 = c.tensor_from(types[0], K, F, data=enc["w1"])
    w3 = c.tensor_from(types[1], K, F, data=enc["w3"])
    w2 = c.tensor_from(types[2], F, K, data=enc["w2"])
    wh = c.tensor_from(N.F32, K, H, data=enc["wh"])
    xn = c.op("rms_norm", x)
    a = c.mul_mat(w1, xn)
    b = c.mul_mat(w3, xn)
    h = c.op("mul", c.op("silu", a), b)
    o = c.op("add", c.mul_mat(w2, h), x)
    head = c.mul_mat(wh, o)                      # 48 x 512 F32 = 96 KB: below the split threshold -> device 0 alone, broadcast
    g = c.build_forward(head)
    return g, dict(x=x, xn=xn, a=a, b=b, h=h, o=o, head=head), enc, X, (K, F, H)


def _run(Nn, types, split, cache, min_bytes=64 << 10, repeats=1):
    rng = np.random.default_rng(1234 + Nn)
    with ggml.Context(64 << 20) as c:
        g, nodes, enc, X, dims = _layer_graph(c, rng, Nn, types, cache)
        N.check(N.host().ggml_host_set_row_split(c.ctx, split, min_bytes))
        outs = None
        for _ in range(repeats):
            c.graph_compute(g)
            got = {k: ggml.tensor_f32(t).copy() for k, t in nodes.items()}
            if outs is not None:
                for k in got:
                    assert np.array_equal(got[k], outs[k]), ("run-to-run", k)
            outs = got
    return outs, enc, X, dims


@pytest.mark.parametrize("Nn", [1, 5, 32])
@pytest.mark.parametrize("types", [(N.Q4_0, N.Q4_0, N.Q4_0), (N.Q4_1, N.F16, N.Q8_0), (N.F32, N.Q5_0, N.F16)])
def test_row_split_equals_single_gpu_bit_for_bit(Nn, types):
    _need2()
    one, enc, X, (K, F, H) = _run(Nn, types, split=0, cache=False)
    two, _, _, _ = _run(Nn, types, split=2, cache=False)
    for k in one:
        assert np.array_equal(one[k], two[k]), (k, rel_l2(two[k], one[k]))
    # ... and the first mul_mat level against the oracle (like with like: the device's own rms_norm output)
    xn = one["xn"].reshape(Nn, K)
    tol = 1e-3 if Nn >= 16 else 6e-6
    assert rel_l2(two["a"].reshape(Nn, F), orc.mul_mat_2d(types[0], enc["w1"], F, K, xn, nth=4)) <= (tol if types[0] != N.F32 else 1e-5)
    assert rel_l2(two["b"].reshape(Nn, F), orc.mul_mat_2d(types[1], enc["w3"], F, K, xn, nth=4)) <= (tol if types[1] != N.F32 else 1e-5)


def test_row_split_with_resident_weight_slices_and_reruns():
    _need2()
    types = (N.Q4_0, N.Q4_1, N.Q4_0)
    one, _, _, _ = _run(1, types, split=0, cache=True, repeats=2)
    N.lib().ggb_reset_stats()
    two, _, _, _ = _run(1, types, split=2, cache=True, repeats=3)
    s = N.stats()
    assert s.weight_uploads == 4 and s.weight_cache_hits == 8            # device 0's view: 4 weights uploaded once (its slices), hit twice more
    for k in one:
        assert np.array_equal(one[k], two[k]), k
    # prompt-sized on the tensor cores, resident slices with their own row exponents
    one, _, _, _ = _run(48, types, split=0, cache=True, repeats=2)
    two, _, _, _ = _run(48, types, split=2, cache=True, repeats=2)
    for k in one:
        assert np.array_equal(one[k], two[k]), k


def test_row_split_uses_every_device_it_is_given():
    n = _ngpu()
    if n < 2:
        pytest.skip("row split needs >= 2 GPUs")
    types = (N.Q4_0, N.Q4_0, N.Q4_0)
    one, _, _, _ = _run(1, types, split=0, cache=False)
    alln, _, _, _ = _run(1, types, split=-1, cache=False)                   # all devices with mutual peer access
    for k in one:
        assert np.array_equal(one[k], alln[k]), k


def test_small_graphs_are_not_split():
    # "only for matrices large enough to benefit": with the default 4 MiB threshold nothing in this graph qualifies; on a 1-GPU box
    # the setting is simply inert
    types = (N.Q4_0, N.Q4_0, N.Q4_0)
    a, _, _, _ = _run(1, types, split=0, cache=False)
    b, _, _, _ = _run(1, types, split=8, cache=False, min_bytes=4 << 20)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
