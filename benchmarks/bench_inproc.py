#!/usr/bin/env python3
"""The row split INSIDE the shim, measured through the reference-shaped API: one process, ggml_graph_compute over a host arena,
G GPUs (include/ggb200.h: ggb_pool_set_row_split).  Host buffers in, host buffers out: activations are read from the pinned arena
by every device, device 0 returns the results; weights are resident slices (weight cache opted in), as in bench.py's `e2e`.

  ring     configs[1] weak-scaled like bench.py --gpus G: 32 nodes, each (4096 G) x 4096 Q4_0, one token
  stack    configs[4]: the Llama-7B-shaped stack, 32 layers x 7 Q4_0 matrices (4.05 GB), one token (strong scaling: G = 1 vs G)
  prompt   configs[4] with 512 tokens on a few layers (the policy caps prompt-sized computes at 2 devices, ggb_shim.cu: row_split_width)

Each record: wall-clock ms per ggml_graph_compute (the call a user makes; includes the exchange, which is fused into the kernels),
the CUDA-event span on device 0, algorithmic GB/s or TFLOP/s.  Standalone:  python benchmarks/bench_inproc.py --gpus 2
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

Q4_0 = 2
LAYER = [(4096, 4096)] * 4 + [(11008, 4096)] * 2 + [(4096, 11008)]      # (M, K)


def _quantized_host(torch, N, L, dev, sp, shapes, seed):
    """Q4_0 weights for `shapes` = [(M, K)], quantized on the device (bit-exact kernel) and brought to the host."""
    import numpy as np
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    out = []
    for (M, K) in shapes:
        rb = K // 32 * 20
        w = torch.empty((M, rb), dtype=torch.uint8, device=dev)
        step = 8192
        for r0 in range(0, M, step):
            r = min(step, M - r0)
            wf = torch.randn((r, K), generator=gen, device=dev) * 0.02
            N.check(L.ggb_dev_quantize_rows(Q4_0, wf.data_ptr(), w[r0:r0 + r].data_ptr(), r, K, sp))
        torch.cuda.synchronize()
        out.append(w.cpu().numpy())
        del w
    return out


def run_graph(torch, N, ggml, shapes, wq, Nn, G, iters, warm=3, seed=7):
    """One graph of independent MUL_MAT nodes (one per shape) through ggml_graph_compute with the row split set to G devices."""
    import numpy as np
    rng = np.random.default_rng(seed)
    arena = sum(w.nbytes + 4 * K * Nn + 4 * M * Nn + 3 * 256 + 1024 for w, (M, K) in zip(wq, shapes)) + (16 << 20)
    L = N.lib()
    with ggml.Context(arena) as c:
        N.check(N.host().ggml_host_set_weight_cache(c.ctx, 1))
        N.check(N.host().ggml_host_set_row_split(c.ctx, G, 0))
        ys, g = [], None
        for w, (M, K) in zip(wq, shapes):
            a = c.tensor_from(Q4_0, K, M, data=w)
            b = c.tensor_from(N.F32, K, Nn, data=rng.standard_normal((Nn, K)).astype(np.float32))
            y = c.mul_mat(a, b)
            ys.append(y)
            if g is None:
                g = c.build_forward(y)
            else:
                N.host().ggml_build_forward_expand(C.byref(g), y)
        for _ in range(warm):
            c.graph_compute(g)
        L.ggb_reset_stats()
        t0 = time.perf_counter()
        for _ in range(iters):
            c.graph_compute(g)
        dt = (time.perf_counter() - t0) / iters
        st = N.stats()
        check = [ggml.tensor_f32(ys[i]).reshape(-1).copy() for i in (0, len(ys) - 1)]
    abytes = sum(M * (K // 32 * 20) + 4 * K * Nn + 4 * M * Nn for (M, K) in shapes)
    flop = sum(2.0 * M * K * Nn for (M, K) in shapes)
    rec = {"n_gpus": G, "nodes": len(shapes), "N": Nn, "ms_per_compute": dt * 1e3, "device0_ms": float(st.last_graph_device_ms),
           "h2d_bytes_per_compute": int(st.h2d_bytes // iters), "d2h_bytes_per_compute": int(st.d2h_bytes // iters)}
    if Nn < 16:
        rec.update({"achieved": abytes / dt / 1e9, "unit": "GB/s"})
    else:
        rec.update({"achieved": flop / dt / 1e12, "unit": "TFLOP/s"})
    return rec, check


def chain_record(torch, N, ggml, layers=32, iters=10, dev_index=0):
    """a DEPENDENT graph, as a decode step really is -- `layers` Llama-7B-shaped layers, one token, every mul_mat fed
    by the previous level through the element-wise neighbours the reference implements (rms_norm, add, silu, mul; it ports no
    attention ops, so q + k + v stands in for the mixing):
        xn = rms_norm(x); a = wo . (wq.xn + wk.xn + wv.xn); x1 = x + a; xn2 = rms_norm(x1); x = x1 + w2 . (silu(w1.xn2) * (w3.xn2))
    15 nodes and 10 dependency levels per layer, 4 of them mul_mat levels (3, 1, 2, 1 nodes wide).  Through ggml_graph_compute on a
    host arena, weights resident.  The floor is the time to stream the weights once at the measured HBM peak."""
    import numpy as np
    L = N.lib()
    dev = torch.device("cuda", dev_index)
    sp = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    shapes = LAYER * layers
    wq = _quantized_host(torch, N, L, dev, sp, shapes, 13)
    wbytes = sum(w.nbytes for w in wq)
    rng = np.random.default_rng(3)
    arena = wbytes + len(shapes) * 1024 + layers * 16 * (4 * 11008 + 1024) + (16 << 20)
    out = {}
    with ggml.Context(arena) as c:
        N.check(N.host().ggml_host_set_weight_cache(c.ctx, 1))
        W = [c.tensor_from(Q4_0, K, M, data=w) for w, (M, K) in zip(wq, shapes)]
        del wq
        x = c.tensor_from(N.F32, 4096, 1, data=rng.standard_normal((1, 4096)).astype(np.float32))
        cur = x
        for l in range(layers):
            w_q, w_k, w_v, w_o, w_1, w_3, w_2 = W[7 * l:7 * l + 7]
            xn = c.op("rms_norm", cur)
            qkv = c.op("add", c.op("add", c.mul_mat(w_q, xn), c.mul_mat(w_k, xn)), c.mul_mat(w_v, xn))
            x1 = c.op("add", cur, c.mul_mat(w_o, qkv))
            xn2 = c.op("rms_norm", x1)
            h = c.op("mul", c.op("silu", c.mul_mat(w_1, xn2)), c.mul_mat(w_3, xn2))
            cur = c.op("add", x1, c.mul_mat(w_2, h))
        g = c.build_forward(cur)
        for mode in ("eager", "graph"):
            os.environ.pop("GGB200_NO_GRAPH_CACHE", None)
            # (the library reads GGB200_NO_GRAPH_CACHE once; the eager figure comes from the first two computes of the pool, which are
            #  never replays, the replay figure from the later ones)
            if mode == "eager":
                t0 = time.perf_counter()
                c.graph_compute(g)
                first = time.perf_counter() - t0          # includes the one-time upload of the weights
                t0 = time.perf_counter()
                c.graph_compute(g)                        # second sighting: enqueued eagerly into a stream capture, then launched
                out["ms_second_compute_recording"] = (time.perf_counter() - t0) * 1e3
                out["ms_first_compute_with_upload"] = first * 1e3
            else:
                L.ggb_reset_stats()
                t0 = time.perf_counter()
                for _ in range(iters):
                    c.graph_compute(g)
                dt = (time.perf_counter() - t0) / iters
                st = N.stats()
                out.update({"ms_per_compute": dt * 1e3, "device_ms": float(st.last_graph_device_ms), "graph_replays": int(st.graph_replays),
                            "kernel_launches_per_compute": int(st.kernel_launches // iters)})
        res = ggml.tensor_f32(cur).reshape(-1).copy()
        # the same chain as one or two launches per dependency level (ggb_set_decode_program(0)): what the persistent launch replaces
        N.check(L.ggb_set_decode_program(0))
        try:
            for _ in range(3):                                # eager, recorded, first replay
                c.graph_compute(g)
            L.ggb_reset_stats()
            t0 = time.perf_counter()
            for _ in range(iters):
                c.graph_compute(g)
            dt = (time.perf_counter() - t0) / iters
            st = N.stats()
            out["per_level_route"] = {"ms_per_compute": dt * 1e3, "device_ms": float(st.last_graph_device_ms), "kernel_launches_per_compute": int(st.kernel_launches // iters),
                                      "bit_identical": bool(np.array_equal(ggml.tensor_f32(cur).reshape(-1).view(np.uint32), res.view(np.uint32)))}
        finally:
            N.check(L.ggb_set_decode_program(1))
    n_levels_mm = 4 * layers
    out.update({"config": "dependent chain through ggml_graph_compute: %d Llama-7B-shaped layers, one token, 15 nodes / 10 levels per layer (4 mul_mat levels: 3, 1, 2, 1 nodes wide), Q4_0 weights resident (%.2f GB)" % (layers, wbytes / 1e9),
                "nodes": int(g.n_nodes), "mul_mat_levels": n_levels_mm, "weights_GB": wbytes / 1e9, "achieved": wbytes / (out["ms_per_compute"] * 1e-3) / 1e9, "unit": "GB/s",
                "finite": bool(np.all(np.isfinite(res))),
                "route": "decode program: the chain is one persistent cooperative launch (k_decode_program) + the result copies" if out["kernel_launches_per_compute"] < n_levels_mm else "per-level launches"})
    return out


def inproc_records(torch, N, ggml, G, dev_index=0, quick=False):
    """Yields records for G GPUs driven by THIS process.  Caller makes sure devices 0..G-1 are visible and otherwise idle."""
    import numpy as np
    L = N.lib()
    dev = torch.device("cuda", dev_index)
    stream = torch.cuda.current_stream(dev)
    sp = C.c_void_p(stream.cuda_stream)
    it = 10 if quick else 30
    # ---- ring: weak scaling, as bench.py's headline ----
    one = _quantized_host(torch, N, L, dev, sp, [(4096, 4096)] * 32, 11)
    for g in sorted({1, G}):
        shapes = [(4096 * g, 4096)] * 32
        wq = [np.tile(w, (g, 1)) for w in one]              # every device's slice of a node is a distinct 10.5 MB matrix (336 MB per GPU > 2x L2)
        rec, chk = run_graph(torch, N, ggml, shapes, wq, 1, g, it)
        rec["config"] = "configs[1] ring through ggml_graph_compute, weak-scaled: 32 nodes x (%d x 4096) Q4_0, one token, row split over %d GPU(s) in ONE process" % (4096 * g, g)
        yield rec
        del wq
    del one
    # ---- stack: strong scaling ----
    shapes = LAYER * 32
    wq = _quantized_host(torch, N, L, dev, sp, shapes, 12)
    ref = None
    for g in sorted({1, G}):
        rec, chk = run_graph(torch, N, ggml, shapes, wq, 1, g, max(5, it // 3))
        rec["config"] = "configs[4] through ggml_graph_compute: 32 layers x 7 Q4_0 matrices (4.05 GB), one token, row split over %d GPU(s) in ONE process (strong scaling)" % g
        if ref is None:
            ref = chk
        else:
            rec["bit_identical_to_1_gpu"] = bool(all(np.array_equal(a, b) for a, b in zip(ref, chk)))
        yield rec
    # ---- prompt: the policy caps it at 2 devices ----
    layers = 2 if quick else 4
    shapes_p = LAYER * layers
    ref = None
    for g in sorted({1, G}):
        rec, chk = run_graph(torch, N, ggml, shapes_p, wq[:len(shapes_p)], 512, g, 3)
        rec["note"] = "tensor-core path: the tile width is chosen per launch from the tile count, so a half-height shard may sum its K steps in another order than the full matrix (same fp16 operands; fp32 rounding differs)"
        rec["config"] = ("configs[4] prompt step through ggml_graph_compute: %d layers x 7 Q4_0 matrices, 512 tokens, row split requested over %d GPU(s) "
                         "(prompt-sized computes use at most 2: the fp32 exchange outgrows the math beyond that)" % (layers, g))
        if ref is None:
            ref = chk
        else:
            rec["rel_l2_vs_1_gpu"] = float(max(np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b.astype(np.float64)), 1e-30) for a, b in zip(chk, ref)))
        yield rec


def main():
    import torch
    from ggmlsharp_b200 import ggml, native as N
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=0)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--chain", type=int, default=0, help="only the dependent-chain record, with this many layers")
    a = ap.parse_args()
    G = a.gpus or torch.cuda.device_count()
    torch.cuda.set_device(0)
    N.check(N.lib().ggb_init())
    if a.chain:
        print(json.dumps(chain_record(torch, N, ggml, layers=a.chain)), flush=True)
        return
    for rec in inproc_records(torch, N, ggml, G, quick=a.quick):
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
