"""Randomised graphs through the reference-shaped host API (ggml_new_tensor / ggml_mul_mat / ggml_add / ... /
ggml_graph_compute): chains and fan-outs of MUL_MAT (every weight type) with the F32 neighbours the executor keeps on the
device, every node's result compared with the oracle's evaluation of the same graph.  Exercises the executor's dependency
levels, device-resident intermediates, in-place SCALE, views (TRANSPOSE looked through) and shared activations."""
import ctypes as C

import numpy as np
import pytest

from gpu_util import rel_l2
from ggmlsharp_b200 import ggml, native as N
from oracle import pyoracle as orc
from test_gpu_parity import weights

pytestmark = pytest.mark.gpu
WTYPES = [N.F32, N.F16, N.Q4_0, N.Q4_1, N.Q4_2, N.Q5_0, N.Q5_1, N.Q8_0]
DIMS = [32, 64, 96, 128, 160, 256, 384, 512]


@pytest.mark.parametrize("seed", range(32))
def test_random_graph_vs_oracle(seed):
    rng = np.random.default_rng(8100 + seed)
    Nn = int(rng.choice([1, 1, 2, 5, 16, 24, 40]))
    K0 = int(rng.choice(DIMS))
    x = rng.standard_normal((Nn, K0)).astype(np.float32)
    with ggml.Context(96 << 20) as c:
        tx = c.tensor_from(N.F32, K0, Nn, data=x)
        live = [(tx, x)]                      # (tensor handle, oracle value [Nn, width]) that later nodes may read
        made = []                             # every compute node: (handle, oracle value, accumulated tolerance)
        tol_of = {id(tx): 0.0}
        outs = []
        for step in range(int(rng.integers(3, 9))):
            h, v = live[int(rng.integers(0, len(live)))]
            tol = tol_of[id(h)]
            op = rng.choice(["mul_mat", "mul_mat", "mul_mat", "add", "mul", "silu", "rms_norm", "scale", "normw"])
            width = v.shape[1]
            if op == "mul_mat":
                t = int(rng.choice(WTYPES))
                M = int(rng.choice(DIMS))
                wb = orc.encode_weights(t, weights(rng, M, width, "weights" if rng.integers(0, 2) else "uniform"))
                nh = c.mul_mat(c.tensor_from(t, width, M, data=wb), h)
                nv = orc.mul_mat_2d(t, wb, M, width, v, nth=4)
                ntol = tol * 3 + (1e-3 if (Nn >= 16 and t != N.F32) else 6e-6)
            elif op in ("add", "mul"):
                same = [(hh, vv) for hh, vv in live if vv.shape == v.shape and hh is not h]
                if same and rng.integers(0, 2):
                    oh, ov = same[int(rng.integers(0, len(same)))]
                    otol = tol_of[id(oh)]
                else:
                    ov = rng.uniform(0.5, 1.5, v.shape).astype(np.float32)
                    oh, otol = c.tensor_from(N.F32, width, Nn, data=ov), 0.0
                nh = c.op(op, h, oh)
                nv = orc.add_f32(v, ov) if op == "add" else orc.mul_f32(v, ov)
                ntol = (tol + otol) * 2 + 1e-6
            elif op == "silu":
                nh, nv = c.op("silu", h), orc.silu_f32(v)
                ntol = tol * 3 + 2e-3             # silu rounds its input to fp16: a 1-ulp change can move a table step
            elif op == "rms_norm":
                nh, nv = c.op("rms_norm", h), orc.rms_norm_f32(v)
                ntol = tol * 3 + 1e-6
            elif op == "normw":                   # the Llama pattern: rms_norm, then * repeat(weight vector)
                nw = rng.uniform(0.5, 1.5, width).astype(np.float32)
                nrm = c.op("rms_norm", h)
                nh = c.op("mul", c.op("repeat", c.tensor_from(N.F32, width, data=nw), nrm), nrm)
                nv = orc.mul_f32(orc.repeat_f32(nw[None, :], Nn, width), orc.rms_norm_f32(v))
                ntol = tol * 3 + 2e-6
            else:                                 # scale: in place on a view of its source -> the source handle dies with it
                if h is tx or any(hh is h for hh, _ in live[:1]):
                    continue
                sv = np.float32(rng.uniform(0.25, 2.0))
                nh = c.op("scale", h, c.tensor_from(N.F32, 1, data=np.float32([sv])))
                nv = orc.scale_f32(v, sv)
                ntol = tol + 1e-6
                live = [(hh, vv) for hh, vv in live if hh is not h]
                made = [(hh, vv, tt) for hh, vv, tt in made if hh is not h]
            tol_of[id(nh)] = ntol
            live.append((nh, nv))
            made.append((nh, nv, ntol))
            outs.append(nh)
        if not made:
            return
        # expanded in creation order, so the cgraph's node order -- which is what defines the reference's semantics for a node
        # that writes over its source (SCALE) -- is the order the oracle values above were computed in
        g = c.build_forward(outs[0])
        for o in outs[1:]:
            N.host().ggml_build_forward_expand(C.byref(g), o)
        N.lib().ggb_reset_stats()
        c.graph_compute(g)                        # raises if a node was left to the (absent) CPU loop
        assert N.stats().nodes_executed == g.n_nodes
        for i, (h, v, tol) in enumerate(made):
            got = ggml.tensor_f32(h).reshape(Nn, -1)
            assert got.shape == v.shape
            err = rel_l2(got, v)
            assert err <= max(tol, 1e-6) * 1.5, (seed, i, err, tol)
