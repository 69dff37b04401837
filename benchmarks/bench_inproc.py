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
        rec["config"] = ("configs[4] prompt step through ggml_graph_compute: %d layers x 7 Q4_0 matrices, 512 tokens, row split requested over %d GPU(s) "
                         "(prompt-sized computes use at most 2: the fp32 exchange outgrows the math beyond that)" % (layers, g))
        if ref is None:
            ref = chk
        else:
            rec["bit_identical_to_1_gpu"] = bool(all(np.array_equal(a, b) for a, b in zip(ref, chk)))
        yield rec


def main():
    import torch
    from ggmlsharp_b200 import ggml, native as N
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=0)
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    G = a.gpus or torch.cuda.device_count()
    torch.cuda.set_device(0)
    N.check(N.lib().ggb_init())
    for rec in inproc_records(torch, N, ggml, G, quick=a.quick):
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
