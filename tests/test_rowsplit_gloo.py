"""The N > 1 path on CPU: world_size-2 gloo processes shard a mul_mat by rows exactly as the reference's thread split does
(Ggml.cs:6665-6672), each rank computes its block (with the oracle -- this is a test of the host-side sharding / gather
logic, there is no GPU here), the blocks are all-gathered, and the result must equal the unsharded oracle bit for bit."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from ggmlsharp_b200 import rowsplit

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_rows_is_the_reference_thread_split():
    for nr, world in ((4096, 8), (11008, 8), (11008, 4), (10, 4), (3, 8), (1, 2), (0, 2)):
        covered = []
        for r in range(world):
            r0, n = rowsplit.shard_rows(nr, world, r)
            dr = (nr + world - 1) // world
            assert r0 == dr * r and n == max(min(r0 + dr, nr) - r0, 0)
            covered += list(range(r0, r0 + n))
        assert covered == list(range(nr))
    assert rowsplit.shard_bytes(11008, 2560, 8, 7) == (7 * 1376 * 2560, 1376 * 2560)      # 11008/8 = 1376 rows, ragged vs 128


def test_the_shim_splits_rows_by_the_same_rule():
    # ggb_row_split_rows is what run_nodes_sharded (the row split inside the shim, one host thread per GPU) uses per device
    import ctypes as C
    from ggmlsharp_b200 import native as N
    for nr, world in ((4096, 8), (11008, 8), (11008, 3), (1408, 2), (10, 4), (3, 8), (1, 2), (0, 2), (64 * 8, 8)):
        covered = []
        for g in range(world):
            r0, n = C.c_int64(), C.c_int64()
            assert N.lib().ggb_row_split_rows(nr, g, world, C.byref(r0), C.byref(n)) == 0
            want0, wantn = rowsplit.shard_rows(nr, world, g)
            assert n.value == wantn and (wantn == 0 or r0.value == want0)
            covered += list(range(r0.value, r0.value + n.value))
        assert covered == list(range(nr))
    r0, n = C.c_int64(), C.c_int64()
    assert N.lib().ggb_row_split_rows(16, 2, 2, C.byref(r0), C.byref(n)) == N.E_INVALID


WORKER = textwrap.dedent('''
    import os, sys
    sys.path.insert(0, %r)
    import numpy as np, torch, torch.distributed as dist
    from ggmlsharp_b200 import rowsplit
    from oracle import pyoracle as orc
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    rng = np.random.default_rng(7)
    ok = True
    for t in (orc.Q4_0, orc.F16, orc.F32, orc.Q4_1):
        for M, K, N in ((300, 256, 1), (129, 128, 3)):
            W = (rng.standard_normal((M, K)) * 0.02).astype(np.float32)
            X = rng.standard_normal((N, K)).astype(np.float32)
            wb = orc.encode_weights(t, W)                               # identical on every rank (same seed)
            r0, n = rowsplit.shard_rows(M, world, rank)
            off, nbytes = rowsplit.shard_bytes(M, wb.shape[1], world, rank)
            mine = wb.reshape(-1)[off:off + nbytes].reshape(n, wb.shape[1])      # this rank's contiguous byte slice
            part = orc.mul_mat_2d(t, mine, n, K, X) if n else np.zeros((N, 0), np.float32)
            # all-gather ragged column blocks: pad to dr rows, gather, then cut
            dr = (M + world - 1) // world
            pad = np.zeros((N, dr), np.float32); pad[:, :n] = part
            out = [torch.zeros(N, dr) for _ in range(world)]
            dist.all_gather(out, torch.from_numpy(pad))
            full = np.concatenate([out[r].numpy()[:, :rowsplit.shard_rows(M, world, r)[1]] for r in range(world)], axis=1)
            want = orc.mul_mat_2d(t, wb, M, K, X)
            ok = ok and np.array_equal(full, want)
    flag = torch.tensor([1 if ok else 0]); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0: print("ROWSPLIT_OK" if flag.item() == 1 else "ROWSPLIT_MISMATCH")
    dist.destroy_process_group()
''')


def test_rowsplit_allgather_two_gloo_ranks(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", str(script)], capture_output=True, text=True, env=env, timeout=240)
    assert "ROWSPLIT_OK" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])


def test_reference_arm_under_torchrun_prints_once():
    # bench.py --impl reference under torchrun: rank 0 alone prints one JSON line, other ranks exit 0 without work
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1", GGB_BENCH_REF_RING="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29534", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, env=env, timeout=300)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert r.returncode == 0 and len(lines) == 1, (r.stdout[-1500:], r.stderr[-1500:])
    import json
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
