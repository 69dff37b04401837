#!/usr/bin/env python3
"""The smallest case of every kernel family, for compute-sanitizer (VERDICT r1 #8c):

    compute-sanitizer --tool memcheck  python benchmarks/sanitize_cases.py
    compute-sanitizer --tool racecheck python benchmarks/sanitize_cases.py
    compute-sanitizer --tool synccheck python benchmarks/sanitize_cases.py

GEMV (TMA-staged and plain-load), tcgen05 GEMM (one node, a grouped launch, 256-column tiles are picked by the cost model only for
larger groups), the expansion fallback, row exponents + activation staging, the codecs, K segments, the executor (two lanes, result
copy kernel, CUDA-graph replay) and the one-rank exchange kernel.  Every result is checked against the oracle so that a sanitizer
run is also a correctness run."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from ggmlsharp_b200 import ggml, native as N, rowsplit
    from oracle import pyoracle as orc
    from test_gpu_parity import dev_mul_mat, weights
    from test_gpu_gemm import dev_mul_mat_batch
    from gpu_util import rel_l2
    rng = np.random.default_rng(1)
    only = set(sys.argv[1:])

    def case(name):
        ok = not only or name in only
        if ok:
            print("case", name, flush=True)
        return ok

    if case("gemv"):
        for t, M, K in ((N.Q4_0, 40, 256), (N.Q4_1, 33, 128), (N.F16, 20, 64), (N.F32, 17, 96), (N.Q5_0, 16, 128), (N.Q4_0, 9, 4128)):
            wb = orc.encode_weights(t, weights(rng, M, K))
            x = rng.standard_normal((1, K)).astype(np.float32)
            assert rel_l2(dev_mul_mat(t, wb, M, K, x), orc.mul_mat_2d(t, wb, M, K, x)) <= 1e-5
    if case("gemm"):
        for t, M, K, Nn in ((N.Q4_0, 128, 128, 16), (N.Q4_1, 130, 256, 17), (N.F16, 128, 128, 16), (N.Q8_0, 128, 128, 16), (N.Q4_0, 64, 160, 16)):
            wb = orc.encode_weights(t, weights(rng, M, K))
            X = rng.standard_normal((Nn, K)).astype(np.float32)
            assert rel_l2(dev_mul_mat(t, wb, M, K, X), orc.mul_mat_2d(t, wb, M, K, X)) <= 1e-3
    if case("grouped"):
        nodes = []
        for t, (M, K, Nn) in ((N.Q4_0, (128, 128, 16)), (N.Q4_0, (256, 256, 32)), (N.F16, (128, 128, 16)), (N.F16, (128, 256, 16))):
            nodes.append((t, orc.encode_weights(t, weights(rng, M, K)), M, K, rng.standard_normal((Nn, K)).astype(np.float32)))
        for (t, wb, M, K, X), y in zip(nodes, dev_mul_mat_batch(nodes)):
            assert rel_l2(y, orc.mul_mat_2d(t, wb, M, K, X)) <= 1e-3
    if case("codecs"):
        x = rng.standard_normal((8, 256)).astype(np.float32)
        for t in (N.Q4_0, N.Q4_1, N.Q5_0, N.Q8_0, N.Q4_2, N.Q5_1):
            q = ggml.quantize_rows(t, x)
            assert np.array_equal(q, orc.quantize_rows(t, x))
            assert np.array_equal(ggml.dequantize_rows(t, q, 256), orc.dequantize_rows(t, q, 256))
    if case("ksegments"):
        os.environ.setdefault("GGB200_KSEG_TEST", "1")
        K = 98304
        wb = orc.encode_weights(N.Q4_0, weights(rng, 3, K))
        x = rng.standard_normal((1, K)).astype(np.float32)
        assert rel_l2(dev_mul_mat(N.Q4_0, wb, 3, K, x), orc.mul_mat_2d(N.Q4_0, wb, 3, K, x)) <= 1e-5
    if case("executor"):
        K, F = 128, 256
        with ggml.Context(8 << 20) as c:
            X = rng.standard_normal((1, K)).astype(np.float32)
            x = c.tensor_from(N.F32, K, 1, data=X)
            ws = [orc.encode_weights(N.Q4_0, weights(rng, F, K)) for _ in range(9)]
            ys = [c.mul_mat(c.tensor_from(N.Q4_0, K, F, data=w), x) for w in ws]            # 9 nodes in one level: the two-lane path
            h = c.op("silu", ys[0])
            g = c.build_forward(h)
            for y in ys[1:]:
                N.host().ggml_build_forward_expand(C.byref(g), y)
            for _ in range(3):                                                              # eager, recorded, replayed
                c.graph_compute(g)
            for w, y in zip(ws, ys):
                assert rel_l2(ggml.tensor_f32(y).reshape(1, F), orc.mul_mat_2d(N.Q4_0, w, F, K, X)) <= 1e-5
    if case("exchange"):
        sym = rowsplit.SymmetricBuffer(4096, 0, 1, lambda h: [h])
        sym.push_barrier(None, 0, 1024, 2048, 2)
        N.check(N.lib().ggb_stream_sync(None))
        sym.close()
    print("sanitize_cases: all cases passed", flush=True)


if __name__ == "__main__":
    main()
