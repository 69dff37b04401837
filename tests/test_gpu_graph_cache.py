"""The executor's CUDA-graph cache (ggb_graph_compute_mul_mats): the second identical compute of a cgraph is recorded, later ones are
replays.  The reference re-plans and re-walks the node list on every ggml_graph_compute (Ggml.cs:3260-3704) and reads tensor->data
afresh each time, so a replay must (a) follow whatever the user wrote into tensor->data since, (b) not be used when anything the
enqueue depended on changed (a SCALE factor, a shape, a data pointer, a resident weight made stale), and (c) keep the per-compute
statistics and perf counters going."""
import ctypes as C

import numpy as np
import pytest

from gpu_util import rel_l2
from ggmlsharp_b200 import ggml, native as N
from oracle import pyoracle as orc
from test_gpu_parity import weights

pytestmark = pytest.mark.gpu


def _ffn(c, rng, Nn, cache):
    K, F = 256, 512
    X = rng.standard_normal((Nn, K)).astype(np.float32)
    W1, W2 = weights(rng, F, K), weights(rng, K, F)
    w1b, w2b = orc.encode_weights(N.Q4_0, W1), orc.encode_weights(N.F16, W2)
    if cache:
        N.check(N.host().ggml_host_set_weight_cache(c.ctx, 1))
    x = c.tensor_from(N.F32, K, Nn, data=X)
    w1 = c.tensor_from(N.Q4_0, K, F, data=w1b)
    w2 = c.tensor_from(N.F16, F, K, data=w2b)
    f = c.tensor_from(N.F32, 1, data=np.array([0.5], np.float32))
    h = c.op("silu", c.mul_mat(w1, c.op("rms_norm", x)))
    y = c.op("scale", c.mul_mat(w2, h), f)
    out = c.op("add", y, x)
    return c.build_forward(out), x, f, w1, out, (w1b, w2b, K, F)


def _expect(X, fval, w1b, w2b, K, F, Nn):
    """The same chain through the oracle's restatements of the reference ops (rms_norm Ggml.cs:5858-5921, silu through table_silu_f16
    5705-5747, scale 6746-6780, add 4622-4685)."""
    xn = orc.rms_norm_f32(np.ascontiguousarray(X, dtype=np.float32))
    a = orc.mul_mat_2d(N.Q4_0, w1b, F, K, xn, nth=2)
    h = orc.silu_f32(a)
    y = orc.mul_mat_2d(N.F16, w2b, K, F, h, nth=2)
    return orc.add_f32(orc.scale_f32(y, np.float32(fval)), np.ascontiguousarray(X, dtype=np.float32))


@pytest.mark.parametrize("Nn", [1, 24])
@pytest.mark.parametrize("cache", [False, True])
def test_replays_follow_the_tensor_data(Nn, cache):
    rng = np.random.default_rng(100 + Nn)
    with ggml.Context(16 << 20) as c:
        g, x, f, w1, out, (w1b, w2b, K, F) = _ffn(c, rng, Nn, cache)
        N.lib().ggb_reset_stats()
        launches = []
        for rep in range(6):
            X = rng.standard_normal((Nn, K)).astype(np.float32) * np.float32(1 + rep)
            ggml.tensor_f32(x).reshape(Nn, K)[...] = X                   # the user writes tensor->data between computes
            before = N.stats().kernel_launches
            c.graph_compute(g)
            launches.append(N.stats().kernel_launches - before)
            got = ggml.tensor_f32(out).reshape(Nn, K).copy()
            want = _expect(X, 0.5, w1b, w2b, K, F, Nn)
            assert rel_l2(got, want) <= (1e-3 if Nn >= 16 else 1e-4), (rep, rel_l2(got, want))
        s = N.stats()
        assert s.graph_replays == 4, s.graph_replays                      # computes 3..6; the 1st runs eagerly, the 2nd records
        assert len(set(launches[1:])) == 1, launches                      # the statistics keep counting per compute
        assert out.contents.perf_runs == 6


def test_a_changed_scale_factor_or_rewritten_weights_are_not_replayed_stale():
    rng = np.random.default_rng(200)
    Nn = 2
    with ggml.Context(16 << 20) as c:
        g, x, f, w1, out, (w1b, w2b, K, F) = _ffn(c, rng, Nn, cache=True)
        X = ggml.tensor_f32(x).reshape(Nn, K).copy()
        for _ in range(3):
            c.graph_compute(g)
        assert N.stats().graph_replays >= 1
        # the SCALE node reads its factor on the host when it is enqueued (Ggml.cs:6763): a new factor is a new graph
        ggml.tensor_f32(f).reshape(-1)[0] = 2.0
        c.graph_compute(g)
        assert rel_l2(ggml.tensor_f32(out).reshape(Nn, K), _expect(X, 2.0, w1b, w2b, K, F, Nn)) <= 1e-4
        # a resident weight rewritten through the API: the mirror goes, and with it every recorded graph that points into it
        W1n = weights(rng, F, K)
        w1n = orc.encode_weights(N.Q4_0, W1n)
        ggml.tensor_bytes(w1)[:] = w1n.ravel()
        N.check(N.lib().ggb_tensor_invalidate(N.host().ggml_host_pool_of(c.ctx), w1))
        for _ in range(4):
            c.graph_compute(g)
            assert rel_l2(ggml.tensor_f32(out).reshape(Nn, K), _expect(X, 2.0, w1n, w2b, K, F, Nn)) <= 1e-4


def test_uncached_weights_are_reread_by_every_replay():
    # default pool (no weight residency): the recorded graph contains the upload of src0 from the host arena, so a replay
    # multiplies whatever bytes the weight tensor holds NOW -- the reference's semantics (Ggml.cs:6139-6164)
    rng = np.random.default_rng(300)
    M, K = 96, 256
    x = rng.standard_normal((1, K)).astype(np.float32)
    with ggml.Context(8 << 20) as c:
        a = c.tensor_from(N.F32, K, M, data=weights(rng, M, K))
        b = c.tensor_from(N.F32, K, data=x)
        y = c.mul_mat(a, b)
        g = c.build_forward(y)
        N.lib().ggb_reset_stats()
        for rep in range(5):
            W = weights(rng, M, K)
            ggml.tensor_bytes(a)[:] = W.view(np.uint8).ravel()
            c.graph_compute(g)
            want = orc.mul_mat_2d(orc.F32, W.view(np.uint8).reshape(M, -1), M, K, x)
            assert rel_l2(ggml.tensor_f32(y).reshape(1, M), want) <= 2e-6, rep
        assert N.stats().graph_replays == 3


def test_perf_time_is_attributed_per_node():
    """Ggml.cs:3695-3703 fills perf_runs / perf_time_us per node.  The executor times each dependency level with CUDA events and splits
    a level's time by the bytes its nodes move, so a 64 MB mul_mat must be charged far more than the 64 KB one that follows it."""
    rng = np.random.default_rng(400)
    K, Mbig, Msmall = 4096, 4096, 16
    with ggml.Context(96 << 20) as c:
        wb = c.tensor_from(N.F32, K, Mbig, data=weights(rng, Mbig, K))
        ws = c.tensor_from(N.F32, Mbig, Msmall, data=weights(rng, Msmall, Mbig))
        x = c.tensor_from(N.F32, K, data=rng.standard_normal((1, K)).astype(np.float32))
        y1 = c.mul_mat(wb, x)
        y2 = c.mul_mat(ws, y1)
        g = c.build_forward(y2)
        for _ in range(4):                                                # eager, recorded, replayed twice
            c.graph_compute(g)
        assert y1.contents.perf_runs == 4 and y2.contents.perf_runs == 4
        assert y1.contents.perf_time_us > 3 * max(y2.contents.perf_time_us, 1), (y1.contents.perf_time_us, y2.contents.perf_time_us)
        assert g.perf_runs == 4 and g.perf_time_us >= y1.contents.perf_time_us
