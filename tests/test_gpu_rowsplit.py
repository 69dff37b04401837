"""Row-split across 2 GPUs through the C ABI (needs >= 2 devices; skipped otherwise): both exchange variants -- peer stores
from the GEMV / tcgen05 GEMM epilogue (ggb_dev_mm.Y_peer) and the push+barrier kernel (ggb_peer_push_barrier) -- must leave, on EVERY
rank, the bytes the unsharded oracle computes (<= the GEMV tolerance) and the same bytes on both ranks."""
import os
import subprocess
import sys
import textwrap

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent('''
    import ctypes as C, os, sys
    sys.path.insert(0, %r)
    import numpy as np, torch, torch.distributed as dist
    from ggmlsharp_b200 import native as N, rowsplit
    from oracle import pyoracle as orc
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    L = N.lib(); N.check(L.ggb_init())
    def ago(o):
        out = [None] * world; dist.all_gather_object(out, o); return out
    rng = np.random.default_rng(5)
    M, K, NODES = 2048, 1024, 3
    Ws = [(rng.standard_normal((M, K)) * 0.02).astype(np.float32) for _ in range(NODES)]
    Xs = [rng.standard_normal((1, K)).astype(np.float32) for _ in range(NODES)]
    TYPES = [N.Q4_0, N.Q5_0, N.Q8_0]                          # one headline and two sibling formats, one node each
    wbs = [orc.quantize_rows(t, w) for t, w in zip(TYPES, Ws)]
    want = np.stack([orc.mul_mat_2d(t, wb, M, K, x)[0] for t, wb, x in zip(TYPES, wbs, Xs)])      # [NODES][M]
    ok = True
    for variant in ("epilogue", "push", "push_gemm", "epilogue_gemm"):
        NB = 24 if variant.endswith("_gemm") else 1           # 24 activation rows -> the tcgen05 path; dst is [NB][M] per node
        if variant.endswith("_gemm"):
            Xg = [rng.standard_normal((NB, K)).astype(np.float32) for _ in range(NODES)]
            wantg = np.stack([orc.mul_mat_2d(t, wb, M, K, x, nth=8) for t, wb, x in zip(TYPES, wbs, Xg)])  # [NODES][NB][M]
        sym = rowsplit.SymmetricBuffer(NODES * NB * M * 4, rank, world, ago)
        r0, n = rowsplit.shard_rows(M, world, rank)
        keep, mms = [], (N.ggb_dev_mm * NODES)()
        def put(a):
            p = C.c_void_p(); N.check(L.ggb_dev_alloc(a.nbytes, C.byref(p))); N.check(L.ggb_dev_upload(p, a.ctypes.data, a.nbytes)); keep.append(p); return p.value
        for i in range(NODES):
            m = mms[i]
            m.type, m.M, m.K, m.N = TYPES[i], n, K, NB
            m.W, m.nb01 = put(np.ascontiguousarray(wbs[i][r0:r0 + n])), wbs[i].shape[1]
            m.X, m.ldx_bytes = put(Xg[i] if variant.endswith("_gemm") else Xs[i]), 4 * K
            off = (i * NB * M + r0) * 4
            m.Y, m.ldy_bytes = sym.payload() + off, 4 * M
            if variant.startswith("epilogue"):
                peers = [r for r in range(world) if r != rank]
                m.n_peers = len(peers)
                for j, r in enumerate(peers): m.Y_peer[j] = sym.payload(r) + off
        wsb = L.ggb_dev_workspace_bytes(mms, NODES)
        ws = C.c_void_p(); N.check(L.ggb_dev_alloc(wsb + 256, C.byref(ws)))
        for rep in range(3):
            N.check(L.ggb_dev_mul_mat_batch(mms, NODES, ws, wsb, None))
            if variant.startswith("epilogue"): sym.barrier(None)
            else: sym.push_barrier(None, r0 * 4, n * 4, M * 4, NODES * NB)      # every dst row of every node is one segment
        N.check(L.ggb_stream_sync(None))
        got = np.zeros((NODES, NB, M), np.float32)
        N.check(L.ggb_dev_download(got.ctypes.data, sym.payload(), got.nbytes))
        ref_ = wantg if variant.endswith("_gemm") else want.reshape(NODES, 1, M)
        err = float(np.linalg.norm(got - ref_) / np.linalg.norm(ref_))
        t = torch.from_numpy(got).cuda(); ref = t.clone(); dist.broadcast(ref, 0)
        same = bool(torch.equal(t, ref))
        ok = ok and err <= (1e-3 if variant.endswith("_gemm") else 5e-6) and same
        dist.barrier(); sym.close()
    flag = torch.tensor([1 if ok else 0], device="cuda"); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0: print("ROWSPLIT_GPU_OK" if flag.item() == 1 else "ROWSPLIT_GPU_MISMATCH")
    dist.destroy_process_group()
''')


def test_two_gpu_rowsplit_exchange(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    env.pop("GGB200_DEVICE", None)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert "ROWSPLIT_GPU_OK" in r.stdout, (r.stdout[-2000:], r.stderr[-3000:])
