#!/usr/bin/env python3
"""BASELINE.json configs[4]: the Llama-7B-shaped layer stack, row-split across the GPUs of one box (STRONG scaling).

32 layers x {wq, wk, wv, wo: 4096x4096; w1, w3: 11008x4096; w2: 4096x11008} Q4_0, distinct random weights per matrix
(4.05 GB in total).  Rank g of G owns rows [dr*g, dr*g + dr) of EVERY matrix, dr = ceil(M / G) -- the reference's thread
split (Ggml.cs:6665-6672) with nth = G -- computes its column block of every dst, and the blocks are exchanged so that
every rank ends the step holding every full dst.  The 224 nodes are independent (as in benchmarks/bench_configs.py cfg4),
so one step = all of them, then the exchange.

  --batch 1     decode step: persistent GEMV launches (HBM-bound); exchange = ggb_peer_push_barrier (one kernel over CUDA-IPC
            peer memory per dst width) or NCCL all-gather
  --batch 512   prompt step: tcgen05 GEMMs; exchange = peer stores fused into the GEMM epilogue + one flag barrier, the push
            kernel, or NCCL all-gather (+ nothing to permute: NCCL gathers into [rank][node][n][rows], which is NOT the
            reference's dst layout -- it is timed as the baseline the fused variants are compared with)
  --exchange none    outputs stay sharded (what a consumer that is itself row-sharded would need): the compute-only line

Run:  python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 benchmarks/bench_stack.py --batch 1
Prints one JSON line per (n, exchange) on rank 0.  Time = CUDA events on the launching stream, max over ranks.
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

LAYER = [(4096, 4096)] * 4 + [(11008, 4096)] * 2 + [(4096, 11008)]      # (M, K)
Q4_0 = 2


def main():
    import torch
    import torch.distributed as dist
    from ggmlsharp_b200 import native as N, rowsplit
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--exchange", default="all", help="comma list of none,push,epilogue,nccl or 'all'")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    dist.init_process_group("nccl", device_id=dev, rank=rank, world_size=world)
    os.environ.setdefault("GGB200_DEVICE", str(lr))
    L = N.lib()
    N.check(L.ggb_init())
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sp = C.c_void_p(stream.cuda_stream)
    Nn = a.batch
    shapes = LAYER * a.layers
    nn = len(shapes)
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm, tf = float(pk.get("hbm_gbs", 6650.0)), float(pk.get("bf16_tflops", 1590.0))

    def ago(o):
        out = [None] * world
        dist.all_gather_object(out, o)
        return out

    # ---- this rank's row slices, quantized on the device ----
    Ws, Xs, rows = [], [], []
    gx = torch.Generator(device=dev); gx.manual_seed(77)                 # activations are replicated: same seed everywhere
    gw = torch.Generator(device=dev); gw.manual_seed(1000 + rank)
    for (M, K) in shapes:
        r0, n = rowsplit.shard_rows(M, world, rank)
        rb = K // 32 * 20
        wf = torch.randn((max(n, 1), K), generator=gw, device=dev) * 0.02
        w = torch.empty((max(n, 1), rb), dtype=torch.uint8, device=dev)
        N.check(L.ggb_dev_quantize_rows(Q4_0, wf.data_ptr(), w.data_ptr(), max(n, 1), K, sp))
        Ws.append(w); rows.append((r0, n))
        Xs.append(torch.randn((Nn, K), generator=gx, device=dev))
        del wf
    torch.cuda.synchronize()

    # dst of every node, full width, grouped by width so one push call covers a group: [nodes of width M][Nn][M]
    widths = sorted(set(M for M, _ in shapes))
    group_nodes = {M: [i for i, s in enumerate(shapes) if s[0] == M] for M in widths}
    group_off, total = {}, 0
    for M in widths:
        group_off[M] = total
        total += len(group_nodes[M]) * Nn * M * 4
    sym = rowsplit.SymmetricBuffer(total, rank, world, ago)

    def node_off(i):
        M = shapes[i][0]
        return group_off[M] + group_nodes[M].index(i) * Nn * M * 4

    def build_mms(peer_stores):
        mms = (N.ggb_dev_mm * nn)()
        for i, (M, K) in enumerate(shapes):
            r0, n = rows[i]
            m = mms[i]
            m.type, m.M, m.K, m.N = Q4_0, n, K, Nn
            m.W, m.nb01 = Ws[i].data_ptr(), K // 32 * 20
            m.X, m.ldx_bytes = Xs[i].data_ptr(), 4 * K
            off = node_off(i) + r0 * 4
            m.Y, m.ldy_bytes = sym.payload() + off, 4 * M
            if peer_stores:
                peers = [r for r in range(world) if r != rank]
                m.n_peers = len(peers)
                for j, r in enumerate(peers):
                    m.Y_peer[j] = sym.payload(r) + off
        return mms

    mm_plain, mm_peer = build_mms(False), build_mms(True)
    wsb = L.ggb_dev_workspace_bytes(mm_plain, nn)
    ws = torch.empty(wsb + 256, dtype=torch.uint8, device=dev)
    wsp = (ws.data_ptr() + 255) // 256 * 256
    # NCCL baseline: gather the per-rank blocks [node][n][rows_r] (only defined when every rank has the same row count)
    even = all(M % world == 0 for M, _ in shapes)
    loc_elems = sum(Nn * rows[i][1] for i in range(nn))
    nccl_src = torch.zeros(loc_elems, device=dev) if even else None
    nccl_dst = torch.zeros(world * loc_elems, device=dev) if even else None

    def step(ex):
        if ex == "epilogue":
            N.check(L.ggb_dev_mul_mat_batch(mm_peer, nn, wsp, wsb, sp))
            sym.barrier(sp)
        else:
            N.check(L.ggb_dev_mul_mat_batch(mm_plain, nn, wsp, wsb, sp))
            if ex == "push":
                for M in widths:
                    r0, n = rowsplit.shard_rows(M, world, rank)
                    sym.push_barrier(sp, group_off[M] + r0 * 4, n * 4, M * 4, len(group_nodes[M]) * Nn)
            elif ex == "nccl":
                dist.all_gather_into_tensor(nccl_dst, nccl_src)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    wbytes = sum(M * (K // 32 * 20) for M, K in shapes)
    abytes = sum(M * (K // 32 * 20) + 4 * K * Nn + 4 * M * Nn for M, K in shapes)
    flop = sum(2.0 * M * K * Nn for M, K in shapes)
    exs = ["none", "push", "epilogue", "nccl"] if a.exchange == "all" else a.exchange.split(",")
    if world == 1:
        exs = ["none"]
    exs.sort(key=["none", "nccl", "push", "epilogue"].index)             # push (device-side epochs) must precede epilogue (host-side epochs)
    for ex in exs:
        if ex == "nccl" and not even:
            continue
        # push and epilogue must not share epochs on one buffer: push uses the device-side epoch, barrier() the host-side one
        if ex == "epilogue":
            sym.epoch = 1 << 40                                         # far above anything the device-side counter reached
        for _ in range(a.warmup):
            step(ex)
        barrier()
        graph = None
        if ex in ("none", "push"):                                      # host launch cost matters at N=1: replay a captured step
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=stream, capture_error_mode="thread_local"):
                step(ex)
            graph.replay()
            barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(a.steps):
            graph.replay() if graph is not None else step(ex)
        e1.record(stream)
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        # correctness of the exchange: every rank must hold what an all-gather of the per-rank column blocks gives
        ok = True
        if ex in ("push", "epilogue"):
            for i in (0, 4, 6, nn - 1):
                M = shapes[i][0]
                r0, n = rows[i]
                N.check(L.ggb_stream_sync(sp))
                host = np.empty((Nn, M), np.float32)
                N.check(L.ggb_dev_download(host.ctypes.data, sym.payload() + node_off(i), host.nbytes))
                mine = torch.from_numpy(host[:, r0:r0 + n].copy()).to(dev)
                if even:
                    allb = torch.empty((world, Nn, n), device=dev)
                    dist.all_gather_into_tensor(allb.view(-1), mine.reshape(-1))
                    want = allb.permute(1, 0, 2).reshape(Nn, M).cpu().numpy()
                    ok = ok and bool((host == want).all())
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            out = {"config": "cfg4 Llama-7B-shaped stack, %d layers x 7 Q4_0 matrices, N=%d, row-split x%d" % (a.layers, Nn, world),
                   "exchange": ex, "n_gpus": world, "ms": ms, "weights_GB_total": wbytes / 1e9, "exchange_verified": bool(flag.item()) if ex in ("push", "epilogue") else None}
            if Nn < 16:
                gbs = abytes / (ms * 1e-3) / 1e9
                out.update({"GB/s": gbs, "frac_of_measured_hbm_xN": gbs / (hbm * world)})
            else:
                tfl = flop / (ms * 1e-3) / 1e12
                out.update({"TFLOP/s": tfl, "frac_of_measured_bf16_xN": tfl / (tf * world)})
            print(json.dumps(out), flush=True)
    barrier()
    sym.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
