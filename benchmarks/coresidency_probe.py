#!/usr/bin/env python3
"""Does a small kernel (the activation staging, 128 threads x 25 registers per CTA) run BESIDE a resident persistent GEMV (512 threads x
118 registers, ~197 KB of shared memory per SM), or does it wait for the GEMV to finish?  Stream A replays a graph of back-to-back
GEMVs (ggb_dev_mul_mat_batch_phase 2, 51 us each); stream B issues one staging call (phase 1) in the middle and times it with events.
A few microseconds = co-resident; ~50 us = it waited for SMs to drain."""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from ggmlsharp_b200 import native as N
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    L = N.lib()
    N.check(L.ggb_init())
    prio = int(os.environ.get("PROBE_PRIORITY", "0"))          # -1 = high priority for the small kernel's stream
    sa, sb = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev, priority=prio)
    pa, pb = C.c_void_p(sa.cuda_stream), C.c_void_p(sb.cuda_stream)
    RING, M, K = 32, 4096, 4096
    rb = K // 32 * 20
    torch.cuda.set_stream(sa)
    Wq = torch.empty((RING, M, rb), dtype=torch.uint8, device=dev)
    for i in range(RING):
        wf = torch.randn((M, K), device=dev) * 0.02
        N.check(L.ggb_dev_quantize_rows(N.Q4_0, wf.data_ptr(), Wq[i].data_ptr(), M, K, pa))
    X = torch.randn((RING, K), device=dev)
    Y = torch.zeros((RING, M), device=dev)
    mm = (N.ggb_dev_mm * RING)()
    for i in range(RING):
        m = mm[i]
        m.type, m.M, m.K, m.N = N.Q4_0, M, K, 1
        m.W, m.nb01, m.X, m.ldx_bytes, m.Y, m.ldy_bytes = Wq[i].data_ptr(), rb, X[i].data_ptr(), 4 * K, Y[i].data_ptr(), 4 * M
    wsb = L.ggb_dev_workspace_bytes(mm, RING)
    ws = torch.empty(2 * wsb + 256, dtype=torch.uint8, device=dev)
    wsp = (ws.data_ptr() + 255) // 256 * 256
    N.check(L.ggb_dev_mul_mat_batch(mm, RING, wsp, wsb, pa))
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=sa, capture_error_mode="thread_local"):
        for _ in range(20):
            N.check(L.ggb_dev_mul_mat_batch_phase(mm, RING, wsp, wsb, pa, 2))
    out = {}
    for label, busy in (("idle GPU", False), ("beside 20 back-to-back GEMVs", True)):
        ts = []
        for _ in range(10):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if busy:
                g.replay()
                time.sleep(0.0003)                     # ~6 GEMVs in
            e0.record(sb)
            N.check(L.ggb_dev_mul_mat_batch_phase(mm, RING, wsp + wsb, wsb, pb, 1))
            e1.record(sb)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        out[label] = {"median_us": ts[len(ts) // 2], "min_us": ts[0], "max_us": ts[-1]}
    print(json.dumps({"probe": "staging kernel (k_act_batch, 32 nodes) on a second stream", "stream_priority": prio, **out}))


if __name__ == "__main__":
    main()
