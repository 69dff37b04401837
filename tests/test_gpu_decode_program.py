"""The decode program (ggb_gemv.cu: k_decode_program): a dependent chain of single-token nodes -- what one ggml_graph_compute of a
decode step is (Ggml.cs:3539-3704 executes the nodes in order) -- runs as ONE persistent launch instead of one or two launches per
dependency level.  It must leave exactly the bytes the per-level path leaves: the row arithmetic is the same code (dot_units, the
Q8 staging of ggb_act_q8.cuh, k_rms_norm_f32's reduction order), only the scheduling differs.  The per-level path is still there
(wide levels, prompt-sized nodes, other ops, row split); ggb_set_decode_program(0) selects it, which is how the two are compared."""
import ctypes as C

import numpy as np
import pytest

from gpu_util import rel_l2
from ggmlsharp_b200 import ggml, native as N
from oracle import pyoracle as orc

pytestmark = pytest.mark.gpu


def weights(rng, M, K):
    return (rng.standard_normal((M, K)) * (1.0 / np.sqrt(K))).astype(np.float32)


def _layers(c, rng, K, F, types, layers, inplace=False, norm_weights=False):
    """The chain of benchmarks/bench_inproc.py's dependent_chain record: per layer rms_norm -> {wq, wk, wv} -> two adds -> wo -> add ->
    rms_norm -> {w1, w3} -> silu, mul -> w2 -> add.  Returns the graph, every node's tensor and the encoded weights."""
    x = c.tensor_from(N.F32, K, 1, data=rng.standard_normal((1, K)).astype(np.float32))
    nodes, enc = {"x": x}, []
    cur = x
    for l in range(layers):
        tq, tk, tv, to, t1, t3, t2 = [types[(7 * l + j) % len(types)] for j in range(7)]
        W = [(tq, K, K), (tk, K, K), (tv, K, K), (to, K, K), (t1, F, K), (t3, F, K), (t2, K, F)]
        wt = []
        for t, M, Kk in W:
            wb = orc.encode_weights(t, weights(rng, M, Kk))
            enc.append((t, wb, M, Kk))
            wt.append(c.tensor_from(t, Kk, M, data=wb))
        xn = c.op("rms_norm", cur)
        if norm_weights:                                         # cur = ggml_mul(ggml_repeat(norm, cur), cur), as a Llama layer has it (one token: ggml_repeat returns norm itself)
            nw = c.tensor_from(N.F32, K, 1, data=rng.uniform(0.5, 1.5, (1, K)).astype(np.float32))
            xn = c.op("mul", c.op("repeat", nw, xn), xn)
        q, k, v = c.mul_mat(wt[0], xn), c.mul_mat(wt[1], xn), c.mul_mat(wt[2], xn)
        qkv = c.op("add", c.op("add", q, k), v)
        o = c.mul_mat(wt[3], qkv)
        x1 = c.op("add_inplace", o, cur) if inplace else c.op("add", cur, o)
        xn2 = c.op("rms_norm", x1)
        a, b = c.mul_mat(wt[4], xn2), c.mul_mat(wt[5], xn2)
        h = c.op("mul", c.op("silu_inplace", a) if inplace else c.op("silu", a), b)
        d = c.mul_mat(wt[6], h)
        cur = c.op("add", x1, d)
        nodes.update({"xn%d" % l: xn, "q%d" % l: q, "v%d" % l: v, "qkv%d" % l: qkv, "o%d" % l: o, "x1_%d" % l: x1, "xn2_%d" % l: xn2,
                      "b%d" % l: b, "h%d" % l: h, "d%d" % l: d, "cur%d" % l: cur})
    return c.build_forward(cur), nodes, enc


def _both_routes(build, arena=96 << 20, cache=True, computes=1):
    """Runs the graph `build(c, rng)` makes through the decode program and through the per-level launches; returns both sets of node
    values, the launches each route made per compute and what build returned besides."""
    out = []
    for program in (1, 0):
        rng = np.random.default_rng(4242)
        with ggml.Context(arena) as c:
            if cache:
                N.check(N.host().ggml_host_set_weight_cache(c.ctx, 1))
            g, nodes, extra = build(c, rng)
            N.check(N.lib().ggb_set_decode_program(program))
            try:
                vals = None
                for _ in range(computes):
                    before = N.stats().kernel_launches
                    c.graph_compute(g)
                    launches = N.stats().kernel_launches - before
                    got = {k: ggml.tensor_f32(t).copy() for k, t in nodes.items()}
                    if vals is not None:
                        for k in got:
                            assert np.array_equal(got[k].view(np.uint32), vals[k].view(np.uint32)), ("run to run", program, k)
                    vals = got
            finally:
                N.check(N.lib().ggb_set_decode_program(1))
            out.append((vals, launches, extra))
    return out


@pytest.mark.parametrize("types", [(N.Q4_0,), (N.Q4_1,), (N.F16,), (N.F32,), (N.Q4_0, N.Q4_1, N.F16, N.F32, N.Q4_0),
                                   (N.Q4_2,), (N.Q5_0,), (N.Q5_1,), (N.Q8_0,), (N.Q5_0, N.Q8_0, N.Q4_2, N.Q5_1, N.Q4_0, N.F16)])
@pytest.mark.parametrize("K,F", [(256, 768), (4096, 11008)])
def test_program_equals_the_per_level_path_bit_for_bit(types, K, F):
    if K == 4096 and (len(types) > 1 or types[0] in (N.Q4_2, N.Q5_1, N.Q8_0)):
        pytest.skip("covered at the small size")
    layers = 2 if K == 256 else 1
    (prog, lp, enc), (lvl, ll, _) = _both_routes(lambda c, rng: _layers(c, rng, K, F, types, layers), arena=((1200 if N.F32 in types else 768) << 20) if K == 4096 else (96 << 20), computes=3)
    assert lp <= 3 and ll >= 12 * layers, (lp, ll)                         # one program launch (+ the result copy) against a launch or two per level
    for k in prog:
        assert np.array_equal(prog[k].view(np.uint32), lvl[k].view(np.uint32)), (k, rel_l2(prog[k], lvl[k]))
    # ... and the first dependent mul_mats against the oracle, like with like (the device's own rms_norm output)
    t, wb, M, Kk = enc[0]
    tol = 1e-5 if t == N.F32 else 6e-6 if t in (N.F16, N.Q4_0, N.Q4_1) else 2e-5
    assert rel_l2(prog["q0"].reshape(1, M), orc.mul_mat_2d(t, wb, M, Kk, prog["xn0"].reshape(1, Kk), nth=4)) <= tol
    t, wb, M, Kk = enc[6]
    assert rel_l2(prog["d0"].reshape(1, M), orc.mul_mat_2d(t, wb, M, Kk, prog["h0"].reshape(1, Kk), nth=4)) <= tol
    assert np.array_equal(prog["xn0"].reshape(1, K), orc.rms_norm_f32(prog["x"].reshape(1, K)))


def test_in_place_ops_and_a_scale_factor():
    def build(c, rng):
        g, nodes, enc = _layers(c, rng, 256, 768, (N.Q4_0, N.Q4_1), 2, inplace=True, norm_weights=True)
        return g, nodes, enc
    (prog, lp, _), (lvl, ll, _) = _both_routes(build, computes=2)
    assert lp <= 3 < ll
    for k in prog:
        assert np.array_equal(prog[k].view(np.uint32), lvl[k].view(np.uint32)), k

    def build2(c, rng):
        K, M = 512, 384
        w1b, w2b = orc.encode_weights(N.Q4_0, weights(rng, M, K)), orc.encode_weights(N.F16, weights(rng, K, M))
        x = c.tensor_from(N.F32, K, 1, data=rng.standard_normal((1, K)).astype(np.float32))
        f = c.tensor_from(N.F32, 1, data=np.array([0.37], np.float32))
        w1, w2 = c.tensor_from(N.Q4_0, K, M, data=w1b), c.tensor_from(N.F16, M, K, data=w2b)
        y1 = c.mul_mat(w1, x)                                    # level 0 multiplies a leaf: the program loads it
        s1 = c.op("scale", y1, f)                                # in place on a mul_mat result
        y2 = c.mul_mat(w2, s1)
        s2 = c.op("scale", c.op("scale", y2, f), f)              # in place twice in a row
        out = c.op("add", s2, x)
        return c.build_forward(out), {"s1": s1, "y2": y2, "out": out}, (w1b, w2b, K, M)
    (prog, lp, (w1b, w2b, K, M)), (lvl, ll, _) = _both_routes(build2, cache=False, computes=2)
    assert lp <= 3 < ll
    for k in prog:
        assert np.array_equal(prog[k].view(np.uint32), lvl[k].view(np.uint32)), k


def test_graphs_the_program_does_not_take_run_as_before():
    # two mul_mats of ONE level on different rows (a wide batch), a prompt-sized chain, and a single level: all per-level
    def wide(c, rng):
        K, M = 256, 128
        ws = [c.tensor_from(N.Q4_0, K, M, data=orc.encode_weights(N.Q4_0, weights(rng, M, K))) for _ in range(2)]
        xs = [c.tensor_from(N.F32, K, 1, data=rng.standard_normal((1, K)).astype(np.float32)) for _ in range(2)]
        w3 = c.tensor_from(N.Q4_0, M, M, data=orc.encode_weights(N.Q4_0, weights(rng, M, M)))
        y = c.op("add", c.mul_mat(ws[0], xs[0]), c.mul_mat(ws[1], xs[1]))
        z = c.mul_mat(w3, y)
        return c.build_forward(z), {"y": y, "z": z}, None

    def prompt(c, rng):
        K, M, Nn = 256, 128, 16
        w1 = c.tensor_from(N.Q4_0, K, M, data=orc.encode_weights(N.Q4_0, weights(rng, M, K)))
        w2 = c.tensor_from(N.Q4_0, M, M, data=orc.encode_weights(N.Q4_0, weights(rng, M, M)))
        x = c.tensor_from(N.F32, K, Nn, data=rng.standard_normal((Nn, K)).astype(np.float32))
        z = c.mul_mat(w2, c.op("silu", c.mul_mat(w1, x)))
        return c.build_forward(z), {"z": z}, None
    for build in (wide, prompt):
        (a, la, _), (b, lb, _) = _both_routes(build, cache=False)
        assert la == lb and la >= 4, (la, lb)
        for k in a:
            assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), k


def test_a_long_chain_is_split_into_several_launches():
    # 160 dependent mul_mat levels: more steps than one parameter block holds (DP_MAX_STEPS = 144)
    def build(c, rng):
        K = 128
        x = c.tensor_from(N.F32, K, 1, data=rng.standard_normal((1, K)).astype(np.float32))
        ws = [c.tensor_from(N.Q4_0, K, K, data=orc.encode_weights(N.Q4_0, weights(rng, K, K))) for _ in range(8)]
        cur, nodes = x, {}
        for l in range(160):
            cur = c.op("add", c.op("rms_norm", c.mul_mat(ws[l % 8], cur)), x)
            if l % 8 == 7:
                nodes["c%d" % l] = cur
        return c.build_forward(cur), nodes, None
    (prog, lp, _), (lvl, ll, _) = _both_routes(build, cache=False, computes=3)
    assert 2 <= lp <= 8 and ll >= 480, (lp, ll)                    # two program launches (+ a result copy for what they did not ship themselves)
    for k in prog:
        assert np.array_equal(prog[k].view(np.uint32), lvl[k].view(np.uint32)), k
    assert np.isfinite(prog["c159"]).all()


def _random_chain(c, rng, n_nodes):
    """A random single-token graph: mul_mats between rows of 256 / 512 / 768 elements and row ops (some in place) on whatever rows
    exist so far.  Most of these are not layer-shaped: mul_mats of one level on different rows, operands produced several steps ago,
    in-place ops on rows other nodes still read -- what the program's builder has to get right or refuse."""
    lens = (256, 512, 768)
    rows = {L: [c.tensor_from(N.F32, L, 1, data=rng.standard_normal((1, L)).astype(np.float32))] for L in lens}
    f = c.tensor_from(N.F32, 1, data=np.array([rng.uniform(0.5, 1.5)], np.float32))
    nodes = {}
    wcache = {}
    for i in range(n_nodes):
        kind = rng.choice(["mm", "mm", "add", "mul", "silu", "rms_norm", "scale", "add_inplace", "silu_inplace"])
        L = int(rng.choice(lens))
        src = rows[L][int(rng.integers(len(rows[L])))] if rng.random() < 0.5 else rows[L][-1]
        if kind == "mm":
            M = int(rng.choice(lens))
            t = [N.Q4_0, N.Q4_1, N.F16, N.F32, N.Q4_2, N.Q5_0, N.Q5_1, N.Q8_0][int(rng.integers(8))]
            key = (t, M, L, int(rng.integers(2)))
            if key not in wcache:
                wcache[key] = c.tensor_from(t, L, M, data=orc.encode_weights(t, weights(rng, M, L)))
            r = c.mul_mat(wcache[key], src)
            rows[M].append(r)
        elif kind in ("add", "mul", "add_inplace"):
            other = rows[L][int(rng.integers(len(rows[L])))]
            if kind == "add_inplace" and (other is src or src is rows[L][0]):
                kind = "add"
            r = c.op(kind, src, other)
            rows[L].append(r)
        elif kind == "scale":
            if src is rows[L][0]:
                continue                                           # (scale works in place: keep the leaves intact for the second route)
            r = c.op("scale", src, f)
            rows[L].append(r)
        else:
            if kind == "silu_inplace" and src is rows[L][0]:
                kind = "silu"
            r = c.op(kind, src)
            rows[L].append(r)
        nodes["n%d" % i] = r
    # one result that depends on everything of its length, so that build_forward reaches most nodes
    outs = [rows[L][-1] for L in lens if len(rows[L]) > 1]
    g = N.ggml_cgraph()
    N.host().ggml_build_forward_into(C.byref(g), outs[0])
    for o in outs[1:]:
        N.host().ggml_build_forward_expand(C.byref(g), o)
    reach = {}
    for k, t in nodes.items():
        for j in range(g.n_nodes):
            if C.addressof(g.nodes[j].contents) == C.addressof(t.contents):
                reach[k] = t
                break
    return g, reach, None


def test_random_single_token_graphs_agree_between_the_routes():
    took = 0
    for seed in range(32):
        def build(c, rng):
            return _random_chain(c, np.random.default_rng(9000 + seed), 28)
        (prog, lp, _), (lvl, ll, _) = _both_routes(build, cache=bool(seed & 1), computes=2)
        assert prog.keys() == lvl.keys() and len(prog) >= 3
        for k in prog:
            assert np.array_equal(prog[k].view(np.uint32), lvl[k].view(np.uint32)), (seed, k, lp, ll)
        took += lp < ll
    print("decode program taken by %d of 32 random graphs" % took)
    assert took >= 12, took                                        # (21 when written; the rest are refused: levels whose mul_mats multiply different rows, a single level)
