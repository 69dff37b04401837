#!/usr/bin/env python3
"""BASELINE.json configs[3]: batched Q4_0 / F16 mul_mat 4096x4096 x 4096x512 on the tcgen05 path, 1 B200.

Reports TFLOP/s (2*M*N*K / CUDA-event time) for the whole node (activation staging + GEMM) and as a fraction of
the measured dense bf16 peak in MEASURED_PEAKS.json.  A ring of distinct weight and activation buffers larger than
L2 is cycled so no launch re-reads a warm cache.  Usage: python benchmarks/bench_gemm.py [--type q4_0|q4_1|f16] [--n 512]
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from ggmlsharp_b200 import native as N
    ap = argparse.ArgumentParser()
    ap.add_argument("--type", default="q4_0")
    ap.add_argument("--m", type=int, default=4096)
    ap.add_argument("--k", type=int, default=4096)
    ap.add_argument("--n", type=int, default=512)
    ap.add_argument("--ring", type=int, default=16)
    ap.add_argument("--iters", type=int, default=64)
    ap.add_argument("--batch", type=int, default=4, help="independent nodes per ggb_dev_mul_mat_batch call in the back-to-back measurement")
    a = ap.parse_args()
    t = {"q4_0": N.Q4_0, "q4_1": N.Q4_1, "f16": N.F16, "f32": N.F32}[a.type]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    L = N.lib()
    N.check(L.ggb_init())
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sp = C.c_void_p(stream.cuda_stream)
    M, K, Nn, R = a.m, a.k, a.n, a.ring
    rb = N.TYPE_SIZE[t] * (K // N.BLCK_SIZE[t])
    W = torch.empty((R, M, rb), dtype=torch.uint8, device=dev)
    for i in range(R):
        wf = torch.randn((M, K), device=dev) * 0.02
        if t == N.F32:
            W[i] = wf.view(torch.uint8).view(M, rb)
        else:
            N.check(L.ggb_dev_quantize_rows(t, wf.data_ptr(), W[i].data_ptr(), M, K, sp))
    X = torch.randn((R, Nn, K), device=dev)
    Y = torch.zeros((R, Nn, M), device=dev)
    mm = (N.ggb_dev_mm * R)()
    for i in range(R):
        m = mm[i]
        m.type, m.M, m.K, m.N = t, M, K, Nn
        m.W, m.nb01, m.X, m.ldx_bytes, m.Y, m.ldy_bytes = W[i].data_ptr(), rb, X[i].data_ptr(), 4 * K, Y[i].data_ptr(), 4 * M
    one = (N.ggb_dev_mm * 1)()
    wsb = L.ggb_dev_workspace_bytes(mm, 1)
    ws = torch.empty(wsb * a.batch + 256, dtype=torch.uint8, device=dev)
    wsp = (ws.data_ptr() + 255) // 256 * 256
    torch.cuda.synchronize()

    def run(i):
        N.check(L.ggb_dev_mul_mat_batch(C.byref(mm[i % R]), 1, wsp, wsb, sp))

    for i in range(8):
        run(i)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.iters)]
    for i, (e0, e1) in enumerate(ev):
        e0.record(stream)
        run(i)
        e1.record(stream)
    torch.cuda.synchronize()
    ms = sorted(e0.elapsed_time(e1) for e0, e1 in ev)
    med, best = ms[len(ms) // 2], ms[0]
    # back-to-back (sustained): `batch` independent nodes per call (a graph level of prompt-batch mul_mats), calls in a stream
    nb = a.batch
    groups = [(N.ggb_dev_mm * nb)(*[mm[(g * nb + j) % R] for j in range(nb)]) for g in range(max(1, R // nb))]
    for g in groups:
        N.check(L.ggb_dev_mul_mat_batch(g, nb, wsp, wsb * nb, sp))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ncalls = max(1, a.iters // nb)
    e0.record(stream)
    for i in range(ncalls):
        N.check(L.ggb_dev_mul_mat_batch(groups[i % len(groups)], nb, wsp, wsb * nb, sp))
    e1.record(stream)
    torch.cuda.synchronize()
    sus = e0.elapsed_time(e1) / (ncalls * nb)
    flop = 2.0 * M * Nn * K
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    pk = float(peaks.get("bf16_tflops", 1590.0))
    out = {"config": "mul_mat %s W[%d x %d] . X[%d x %d]" % (a.type, M, K, Nn, K), "ring": R,
           "ms_median": med, "ms_best": best, "ms_back_to_back": sus,
           "tflops_median": flop / med / 1e9, "tflops_best": flop / best / 1e9, "tflops_back_to_back": flop / sus / 1e9,
           "frac_of_measured_bf16_peak": flop / med / 1e9 / pk, "frac_of_nominal_2250": flop / med / 1e9 / 2250.0,
           "peak_tflops": pk, "nodes_per_call": nb}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
