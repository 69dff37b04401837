#!/usr/bin/env python3
"""bench.py -- the headline benchmark of BASELINE.json on B200.

Metric: "Q4_0 mul_mat HBM GB/s (% roofline)".  Workload at N=1 = BASELINE.json configs[1]: Q4_0 weight
4096x4096 GEMV, single token.  One matrix is 10.5 MB -- smaller than the 126 MB L2 and shorter than a kernel
launch -- so a "step" is one pass of the mul_mat path over a RING of 32 distinct 4096x4096 Q4_0 matrices
(335.5 MB > 2x L2), each with its own activation vector: 32 independent MUL_MAT graph nodes, which the executor
fuses into one activation-quantize launch + one persistent GEMV launch.  Inputs are larger than L2, so no flush.

  value     whole-job GB/s = algorithmic bytes (W + x + y per node) / device time, weights, activations
            and outputs resident in HBM, timed with CUDA events on the launching stream.
  e2e       the same bytes / wall time through the reference-shaped API (ggml_graph_compute of a 32-node graph over a
            host arena): per step the activations go host->device and the results device->host inside the timing.
  roofline  the GEMV kernel alone: algorithmic bytes per launch / its event-timed duration vs MEASURED_PEAKS.json.
  cpu_baseline / --impl reference: the C oracle's restatement of the reference's CPU mul_mat on the box's host cores
            (the reference is C#; no .NET toolchain exists here, so kind = "port").

N > 1 (torchrun, one rank per GPU): weak scaling of the row split.  Every node's weight matrix grows to
(4096*N) x 4096; rank r owns rows [4096 r, 4096 (r+1)) -- the reference's thread row split (Ggml.cs:6665-6672) with
nth = N -- and the outputs are all-gathered so every rank ends the step with the full result.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M_LOCAL, K, RING = 4096, 4096, 32
Q4_0 = 2


def alg_bytes_per_node(m, k, n=1, row_bytes=None):
    rb = row_bytes if row_bytes is not None else k // 32 * 20
    return m * rb + 4 * k * n + 4 * m * n


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_mul_mat_ring(nodes, nth, seconds_budget=None, repeats=1):
    """The oracle's CPU mul_mat (reference algorithm: serial INIT quantization + row-split scalar dots) over `nodes`
    = [(wbytes, x)], `repeats` times.  Returns seconds per pass (best)."""
    from oracle import pyoracle as orc
    best = None
    t_start = time.perf_counter()
    for _ in range(repeats):
        t0 = time.perf_counter()
        for wb, x in nodes:
            orc.mul_mat_2d(orc.Q4_0, wb, wb.shape[0], K, x, nth=nth)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        if seconds_budget and time.perf_counter() - t_start > seconds_budget:
            break
    return best


def host_ring(n_nodes, m, seed=1001):
    """Synthetic weights N(0, 0.02^2) -> oracle quantize_row_q4_0, activations N(0,1) (SURVEY 8d generators)."""
    import numpy as np
    from oracle import pyoracle as orc
    rng = np.random.default_rng(seed)
    nodes = []
    base = (rng.standard_normal((m, K)) * 0.02).astype(np.float32)
    for i in range(n_nodes):
        w = np.roll(base, i * 17 + 1, axis=0) if i else base      # distinct bytes per node without 32 full RNG passes
        nodes.append((orc.quantize_rows(orc.Q4_0, w), rng.standard_normal((1, K)).astype(np.float32)))
    return nodes


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (C restatement; see module docstring) on the host cores."""
    if rank != 0:
        return
    nth = os.cpu_count() or 1
    ring = int(os.environ.get("GGB_BENCH_REF_RING", RING))        # tests shrink it
    nodes = host_ring(ring, M_LOCAL * world)
    # bounded sample: as many ring nodes per step as keep the whole --steps/--warmup run within ~2 minutes of CPU time
    t_node = cpu_mul_mat_ring(nodes[:1], nth, repeats=2)
    budget = 120.0 / max(args.steps + args.warmup, 1)
    sample_nodes = int(max(1, min(ring, budget / max(t_node, 1e-6))))
    nodes = nodes[:sample_nodes]
    for _ in range(args.warmup):
        cpu_mul_mat_ring(nodes, nth)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_mul_mat_ring(nodes, nth)
    dt = (time.perf_counter() - t0) / args.steps
    nbytes = sum(alg_bytes_per_node(M_LOCAL * world, K) for _ in nodes)
    val = nbytes / dt / 1e9
    line = {"impl": "reference", "metric": "Q4_0 mul_mat HBM GB/s (% roofline)", "value": val, "unit": "GB/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int8 dot (Q4_0 x Q8_0), f32 scales", "data": "synthetic",
            "config": {"workload": "configs[1]: Q4_0 4096x4096 GEMV, single token; ring of %d distinct matrices (%d rows each)" % (len(nodes), M_LOCAL * world)},
            "cpu_baseline": {"value": val, "unit": "GB/s", "cores": nth, "kind": "port",
                             "sample": "%d of the %d ring GEMVs per step (bounded to ~2 min total); C restatement of ggml_compute_forward_mul_mat_q_f32 with %d threads (not .NET RyuJIT)" % (len(nodes), ring, nth)},
            "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-configuration sub-records (configs[0], [2], [3], [4]) of the N=1 line")
    ap.add_argument("--gather", default="fused", choices=["nccl", "fused"])
    ap.add_argument("--no-graph", action="store_true", help="launch every step from Python instead of replaying a captured CUDA graph")
    ap.add_argument("--no-pipeline", action="store_true", help="one ggb_dev_mul_mat_batch call per step (staging and GEMV back to back) instead of staging step i+1 under the GEMVs of step i")
    ap.add_argument("--epilogue-stores", action="store_true", help="fused gather through per-row peer stores in the GEMV epilogue instead of the push kernel")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from ggmlsharp_b200 import ggml, native as N

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the mul_mat path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    L = N.lib()
    os.environ.setdefault("GGB200_DEVICE", str(local_rank))
    N.check(L.ggb_init())
    # a real (non-default) stream: a NULL stream handle means "the library's own stream" in the C ABI, and CUDA
    # events only see work on the stream they are recorded on
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sptr = C.c_void_p(stream.cuda_stream)
    assert sptr.value, "expected a non-default stream"

    # ---- device-resident workload: RING nodes, this rank's 4096-row slice of each ----
    rb = K // 32 * 20
    M_total = M_LOCAL * world
    gen = torch.Generator(device=dev); gen.manual_seed(1001 + rank)
    Wq = torch.empty((RING, M_LOCAL, rb), dtype=torch.uint8, device=dev)
    for i in range(RING):
        wf = torch.randn((M_LOCAL, K), generator=gen, device=dev, dtype=torch.float32) * 0.02
        N.check(L.ggb_dev_quantize_rows(Q4_0, wf.data_ptr(), Wq[i].data_ptr(), M_LOCAL, K, sptr))
    xgen = torch.Generator(device=dev); xgen.manual_seed(2001)     # same activations on every rank (replicated)
    X = torch.randn((RING, K), generator=xgen, device=dev, dtype=torch.float32)
    gather = args.gather if world > 1 else "none"
    sym = None
    # Pipelined steps (default): the steps of this benchmark are independent passes over the ring, so the activation staging of step
    # i+2 (ggb_dev_mul_mat_batch_phase 3: no wait for its predecessor, into the third of three workspaces) is issued IN the GEMV
    # stream between the GEMVs of steps i and i+1 (phase 2) and runs beside them (its 128-thread CTAs fit next to a resident GEMV
    # CTA), and -- N > 1 -- the exchange of step i on a second stream under step i+1.  Consecutive GEMVs then stay back-to-back
    # programmatic launches, as in the kernel-only roofline pass.  Why the orders hold: a GEMV's consumers wait for the staging
    # kernel in front of it to complete, and that kernel completes only after the GEMV in front of IT has (trailing
    # griddepcontrol.wait) -- so GEMV j+1 starts after GEMV j, which had waited for staging j+1; and staging j+3, which overwrites
    # workspace j mod 3, is launched only once GEMV j+1 has passed that wait, i.e. after GEMV j is complete.  Every step still stages
    # AND multiplies inside the timed region.  --no-pipeline: one ggb_dev_mul_mat_batch call per step, as a lone decode step is issued.
    pipelined = not args.no_pipeline and gather in ("none", "fused") and not args.epilogue_stores
    NBUF = 2 if (pipelined or (gather == "fused" and not args.epilogue_stores)) else 1
    mm_bufs = [(N.ggb_dev_mm * RING)() for _ in range(NBUF)]
    if gather == "fused":
        # every rank holds the FULL dst of every node, [node][rows_total] (x2: double buffered).  The exchange is one kernel
        # over CUDA-IPC mapped peer memory: coalesced push of this rank's blocks into every peer's copy + a flag barrier.
        # It runs on its own stream, so the exchange of step i overlaps the GEMVs of step i+1 (the nodes of consecutive
        # steps are independent); the timed region ends only when every step's exchange has completed.
        from ggmlsharp_b200 import rowsplit

        def ago(obj):
            out = [None] * world
            dist.all_gather_object(out, obj)
            return out
        sym = rowsplit.SymmetricBuffer(NBUF * RING * M_total * 4, rank, world, ago)
        Y = None
    else:
        Y = torch.zeros((world, RING, M_LOCAL), dtype=torch.float32, device=dev)     # all-gather layout [rank][node][rows]
        Yloc = Y[rank]
    for bsel in range(NBUF):
        for i in range(RING):
            m = mm_bufs[bsel][i]
            m.type, m.M, m.K, m.N = Q4_0, M_LOCAL, K, 1
            m.W, m.nb01 = Wq[i].data_ptr(), rb
            m.X, m.ldx_bytes = X[i].data_ptr(), 4 * K
            if sym is not None:
                off = ((bsel * RING + i) * M_total + rank * M_LOCAL) * 4
                m.Y, m.ldy_bytes = sym.payload() + off, 4 * M_total
                if args.epilogue_stores:                      # variant: the GEMV epilogue itself stores into every peer (4-byte NVLink writes)
                    peers = [r for r in range(world) if r != rank]
                    m.n_peers = len(peers)
                    for j, r in enumerate(peers):
                        m.Y_peer[j] = sym.payload(r) + off
            else:
                m.Y, m.ldy_bytes = Yloc[i].data_ptr(), 4 * M_LOCAL
    mms = mm_bufs[0]
    wsb = L.ggb_dev_workspace_bytes(mms, RING)
    NWS = 3 if pipelined else NBUF
    ws = torch.empty(NWS * wsb + 256, dtype=torch.uint8, device=dev)
    wsp = (ws.data_ptr() + 255) // 256 * 256
    # HIGH priority: the next GEMV's CTAs are already pending (programmatic launch) when the push is issued, and the block scheduler
    # serves pending CTAs of equal priority in order -- the 64 small push CTAs would wait a whole GEMV behind CTAs that cannot be
    # placed yet (benchmarks/coresidency_probe.py: 50 us at equal priority, 19 us = launch + run at high priority)
    comm = torch.cuda.Stream(device=dev, priority=-1) if (NBUF == 2 and gather == "fused") else None
    cptr = C.c_void_p(comm.cuda_stream) if comm is not None else None

    class Ev:                                                # one set of events per buffer: staged, multiplied, exchanged
        def __init__(self):
            self.stage = [torch.cuda.Event() for _ in range(2)]
            self.mul = [torch.cuda.Event() for _ in range(2)]
            self.comm = [torch.cuda.Event() for _ in range(2)]
            self.fork = torch.cuda.Event()
            self.staged, self.multiplied, self.pushed = [False, False], [False, False], [False, False]
    ev_live = Ev()
    total_steps = [0]
    torch.cuda.synchronize()

    def pipe(n, E):
        """n pipelined steps (see `pipelined` above); workspaces 0, 1, 2, 0, ... from the start of every call, dst buffers by step parity."""
        if n <= 0:
            return

        def stage(w, phase):
            N.check(L.ggb_dev_mul_mat_batch_phase(mm_bufs[0], RING, wsp + w * wsb, wsb, sptr, phase))
        # the first two stagings of a sequence wait for whatever is still in flight on the stream (an earlier sequence's GEMVs)
        stage(0, 1)
        if n > 1:
            stage(1, 1)
        for j in range(n):
            b = (total_steps[0] + j) & 1 if NBUF == 2 else 0
            if comm is not None and E.pushed[b]:
                stream.wait_event(E.comm[b])                # the exchange that last read / filled dst buffer b is done
            N.check(L.ggb_dev_mul_mat_batch_phase(mm_bufs[b], RING, wsp + (j % 3) * wsb, wsb, sptr, 2))
            if comm is not None:
                E.mul[b].record(stream)
            if j + 2 < n:
                stage((j + 2) % 3, 3)                       # two steps ahead, beside the GEMVs of steps j and j+1
            if comm is not None:
                comm.wait_event(E.mul[b])
                dbg = os.environ.get("GGB_BENCH_EXCHANGE", "")      # diagnosis only: "none" = no exchange kernel, "barrier" = flag barrier without the copy
                if dbg == "barrier":
                    sym.push_barrier(cptr, 0, 16, 16, 0)
                elif dbg != "none":
                    sym.push_barrier(cptr, (b * RING * M_total + rank * M_LOCAL) * 4, M_LOCAL * 4, M_total * 4, RING)
                E.comm[b].record(comm)
                E.pushed[b] = True
        total_steps[0] += n

    def step_body(b, E, wait=True):
        # unpipelined fused step on buffer b: GEMVs on the compute stream, exchange on the comm stream (overlaps the next step's GEMVs)
        if wait and E.pushed[b]:
            stream.wait_event(E.comm[b])                    # the exchange that last read / filled buffer b is done
        N.check(L.ggb_dev_mul_mat_batch(mm_bufs[b], RING, wsp + b * wsb, wsb, sptr))
        E.mul[b].record(stream)
        comm.wait_event(E.mul[b])
        sym.push_barrier(cptr, (b * RING * M_total + rank * M_LOCAL) * 4, M_LOCAL * 4, M_total * 4, RING)
        E.comm[b].record(comm)
        E.pushed[b] = True

    def steps_eager(n, E):
        if pipelined:
            pipe(n, E)
            return
        for _ in range(n):
            if gather == "fused" and NBUF == 2:
                step_body(total_steps[0] & 1, E)
            else:
                N.check(L.ggb_dev_mul_mat_batch(mms, RING, wsp, wsb, sptr))
                if gather == "fused":
                    sym.barrier(sptr)
                elif gather == "nccl":
                    dist.all_gather_into_tensor(Y.view(-1), Yloc.reshape(-1))
            total_steps[0] += 1

    # Host launch cost (2 kernels + exchange + ~8 event ops per step from Python) is comparable to the ~55 us of GPU work, so
    # GSTEPS steps are captured once into a CUDA graph (all streams) and replayed; the flag-barrier epoch lives in device memory.
    GSTEPS = 8
    graph = None
    use_graph = (pipelined or (gather == "fused" and NBUF == 2)) and not args.no_graph

    def build_graph(nsteps):
        g = torch.cuda.CUDAGraph()
        E = Ev()                                             # events that live inside the capture
        saved = total_steps[0]
        with torch.cuda.graph(g, stream=stream, capture_error_mode="thread_local"):
            # replays serialise on the stream, so the first use of each buffer in a replay has nothing to wait for
            steps_eager(nsteps, E)
            if comm is not None:
                for b in range(2):
                    if E.pushed[b]:
                        stream.wait_event(E.comm[b])        # join the exchange stream back before the capture ends
        total_steps[0] = saved
        return g

    rem_graphs = {}

    def run_steps(n):
        """exactly n steps: replays of the GSTEPS-step graph, then ONE replay of a graph of the remaining n % GSTEPS steps (captured on
        first use, outside any timed region: prepare_steps), so that a short --steps run is not a mix of replayed and Python-issued steps"""
        if graph is not None:
            if comm is not None:
                stream.wait_stream(comm)                    # Python-issued steps may still have exchange work in flight on the side stream
            for _ in range(n // GSTEPS):
                graph.replay()
            total_steps[0] += (n // GSTEPS) * GSTEPS
            n = n % GSTEPS
            if n and n in rem_graphs and not (total_steps[0] & 1):
                rem_graphs[n].replay()
                total_steps[0] += n
                n = 0
        if n:
            steps_eager(n, ev_live)

    def prepare_steps(n):
        r = n % GSTEPS
        if graph is not None and r and r not in rem_graphs:
            if total_steps[0] & 1:
                steps_eager(1, ev_live)
            torch.cuda.synchronize()
            rem_graphs[r] = build_graph(r)
            rem_graphs[r].replay()
            total_steps[0] += r
            if total_steps[0] & 1:
                steps_eager(1, ev_live)                     # leave the buffer parity even: every graph starts on buffer 0
            torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    steps_eager(args.warmup, ev_live)
    barrier()
    if use_graph:
        if total_steps[0] & 1:
            steps_eager(1, ev_live)                         # graphs start on buffer 0
        torch.cuda.synchronize()
        graph = build_graph(GSTEPS)
        graph.replay()
        total_steps[0] += GSTEPS
        prepare_steps(args.steps)
        barrier()
    L.ggb_reset_stats()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # keep the timed region long enough for nvidia-smi to sample it (>= ~0.5 s), but time EXACTLY --steps steps
    barrier()
    e0.record(stream)
    run_steps(args.steps)
    if comm is not None:
        for b in range(2):
            if ev_live.pushed[b]:
                stream.wait_event(ev_live.comm[b])          # every step's exchange is inside the timed region (graph replays join by themselves)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = int(N.stats().kernel_launches)
    if graph is not None:
        replayed = (args.steps // GSTEPS) * GSTEPS + (args.steps % GSTEPS if args.steps % GSTEPS in rem_graphs else 0)
        launches += replayed * (3 if comm is not None else 2)      # replayed steps: act + GEMV (+ exchange) kernels each
    # extra untimed steps under the sampler so short runs still see clocks under load; the SAME count on every rank
    # (each step ends in a collective / flag barrier)
    n_extra = int(min(20000, max(10, 0.6 / max(ms / args.steps * 1e-3, 1e-6))))
    if world > 1:
        t = torch.tensor([n_extra], device=dev, dtype=torch.int64)
        dist.broadcast(t, 0)
        n_extra = int(t.item())
    run_steps(n_extra)
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    step_bytes_rank = RING * alg_bytes_per_node(M_LOCAL, K)
    value = world * step_bytes_rank / (ms_per_step * 1e-3) / 1e9

    # ---- the pipelined steps must leave the bytes one plain call leaves (N = 1; N > 1 is checked against an all-gather below) ----
    if pipelined and world == 1:
        torch.cuda.synchronize()
        got = Y.clone()
        Y.zero_()
        N.check(L.ggb_dev_mul_mat_batch(mms, RING, wsp, wsb, sptr))
        torch.cuda.synchronize()
        assert torch.equal(got, Y), "pipelined steps differ from a plain ggb_dev_mul_mat_batch call"

    # ---- the fused exchange must leave the same bytes everywhere as a plain all-gather of the per-rank blocks ----
    if gather == "fused":
        mine = np.zeros((RING, M_total), dtype=np.float32)
        N.check(L.ggb_stream_sync(sptr))
        torch.cuda.synchronize()
        lastb = (total_steps[0] - 1) & 1 if NBUF == 2 else 0
        N.check(L.ggb_dev_download(mine.ctypes.data, sym.payload() + lastb * RING * M_total * 4, mine.nbytes))
        loc = torch.from_numpy(mine[:, rank * M_LOCAL:(rank + 1) * M_LOCAL].copy()).to(dev)
        allb = torch.zeros((world, RING, M_LOCAL), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(allb.view(-1), loc.reshape(-1))
        want = allb.permute(1, 0, 2).reshape(RING, M_total).cpu().numpy()
        assert os.environ.get("GGB_BENCH_EXCHANGE") or np.array_equal(mine, want), "fused row-split exchange differs from the all-gather of the per-rank blocks"

    # ---- roofline of the dominant kernel (the persistent GEMV): CUDA events around each GEMV launch, inside the library,
    #      on the launching stream (ggb_set_kernel_timing).  The brackets disable the PDL overlap, so this is a separate pass.
    #      Mode 2 relaunches the GEMV alone on the activations the last step staged: a back-to-back stream of the dominant kernel,
    #      bracketed by one event pair, so per-launch duration = elapsed / launches (no host or event gap inside).
    L.ggb_reset_stats()
    N.check(L.ggb_dev_mul_mat_batch(mms, RING, wsp, wsb, sptr))          # stage activations for mms in workspace 0
    N.check(L.ggb_set_kernel_timing(2))
    # The launches are captured into a CUDA graph and replayed, so that the figure does not depend on how fast this host can issue
    # launches from Python (round 1: 57.3 us per launch on one box, 51.8 on another -- the slower one was the host, not the kernel).
    KINNER = 10
    n_prof = max(2, min(args.steps, 100) // KINNER) * KINNER
    for _ in range(3):
        N.check(L.ggb_dev_mul_mat_batch(mms, RING, wsp, wsb, sptr))
    torch.cuda.synchronize()
    st0 = N.stats()
    kgraph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(kgraph, stream=stream, capture_error_mode="thread_local"):
        for _ in range(KINNER):
            N.check(L.ggb_dev_mul_mat_batch(mms, RING, wsp, wsb, sptr))
    st = N.stats()
    assert st.timed_kernel_launches - st0.timed_kernel_launches == KINNER, "expected exactly one GEMV launch per call in the kernel-only pass"
    kgraph.replay()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    k0.record(stream)
    for _ in range(n_prof // KINNER):
        kgraph.replay()
    k1.record(stream)
    torch.cuda.synchronize()
    N.check(L.ggb_set_kernel_timing(0))
    gemv_ms = k0.elapsed_time(k1) / n_prof
    peak, peak_src = load_peaks()
    achieved = step_bytes_rank / (gemv_ms * 1e-3) / 1e9
    traffic = None                                          # DRAM bytes per launch from the committed ncu --set full capture
    tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)["k_gemv_fast"]
        traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
    line = {"metric": "Q4_0 mul_mat HBM GB/s (% roofline)", "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int8 dot (Q4_0 x Q8_0), f32 scales", "data": "synthetic",
            "config": {"workload": "configs[1]: Q4_0 4096x4096 GEMV, single token", "ring": RING,
                       "step": "%d independent MUL_MAT nodes (distinct weights, %.1f MB > 2x L2, no flush needed), 1 act + 1 GEMV launch" % (RING, RING * M_LOCAL * rb / 1e6),
                       "pipelining": ("steps are independent: the activation staging of step i+2 is issued between the GEMVs of steps i and i+1 and runs beside them (three workspaces; consecutive GEMVs stay back-to-back programmatic launches)%s; every step stages and multiplies inside the timed region; %d steps per CUDA-graph replay" % ("; the exchange of step i under step i+1" if world > 1 else "", GSTEPS)) if pipelined else "none (--no-pipeline)",
                       "rows_per_rank": M_LOCAL, "rows_total": M_total, "k": K,
                       "parallelism": "row-split x%d + all-gather (%s)" % (world, args.gather) if world > 1 else "1 GPU"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of this kernel on this workload from the committed ncu --set full capture (profiles/r02_gemv_q4_0_ncu_full.txt); ncu cannot run inside the timed bench",
                         "kernel": "k_gemv_fast<Q4_0,1,xreg>", "launch_ms": gemv_ms, "launches_timed": n_prof, "how": "GEMV kernel relaunched back to back on staged activations (CUDA graph of %d launches, replayed), one CUDA-event pair around %d launches" % (KINNER, n_prof), "peak_source": peak_src,
                         "frac_of_nominal_8TBs": achieved / 8000.0, "bytes_per_launch": step_bytes_rank},
            "gpu_launches": launches, "clocks": clocks}

    # ---- e2e: reference-shaped API over a host arena (weights cached on device after the first compute).  Every rank runs
    #      ggml_graph_compute over its own row slice (activations host->device, this rank's dst block device->host each step);
    #      the time is the max over ranks, the bytes are all ranks'. ----
    arena = RING * (M_LOCAL * rb + 4 * K + 4 * M_LOCAL + 3 * 256) + (8 << 20)
    wq_host = Wq.cpu().numpy()
    x_host = X.cpu().numpy()
    with ggml.Context(arena) as c:
        # weight residency is opt-in (include/ggb200.h: ggb_pool_set_weight_cache); the reference re-reads src0 on every compute
        N.check(N.host().ggml_host_set_weight_cache(c.ctx, 1))
        ys, g = [], None
        for i in range(RING):
            a = c.tensor_from(N.Q4_0, K, M_LOCAL, data=wq_host[i])
            b = c.tensor_from(N.F32, K, data=x_host[i])
            y = c.mul_mat(a, b)
            ys.append(y)
            if g is None:
                g = c.build_forward(y)
            else:
                N.host().ggml_build_forward_expand(C.byref(g), y)
        for _ in range(3):
            c.graph_compute(g)
        L.ggb_reset_stats()
        n_e2e = max(10, min(args.steps, 100))
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            c.graph_compute(g)
        dt = (time.perf_counter() - t0) / n_e2e
        s_e2e = N.stats()
        y0 = ggml.tensor_f32(ys[0]).reshape(-1).copy()
    if world > 1:
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    e2e = {"value": world * step_bytes_rank / dt / 1e9, "unit": "GB/s", "ms_per_step": dt * 1e3,
           "h2d_bytes_per_step": int(s_e2e.h2d_bytes // n_e2e) * world, "d2h_bytes_per_step": int(s_e2e.d2h_bytes // n_e2e) * world,
           "device_ms_per_step": float(s_e2e.last_graph_device_ms),       # CUDA-event span of the last call: staging reads over PCIe + kernels + result copy
           "api": "ggml_graph_compute over a %d-node graph per rank (host arena, weights device-cached)%s" % (RING, "; each rank reads back its own dst block" if world > 1 else "")}
    # ---- every other BASELINE.json configuration on this GPU, as sub-records of the same line (device-resident, CUDA events on the
    #      launching stream; benchmarks/bench_configs.py: Runner).  configs[1] -- the ring above -- stays the line's `value`. ----
    if world == 1 and not args.no_configs:
        from benchmarks.bench_configs import Runner
        pk = {}
        if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                pk = json.load(f)
        del Wq, X, ws
        torch.cuda.empty_cache()
        runner = Runner(torch, N, L, dev, stream, float(pk.get("hbm_gbs", 6650.0)), float(pk.get("bf16_tflops", 1590.0)))
        line["configs"] = list(runner.baseline_records(iters=20))
        # ... and a DEPENDENT decode step (32 layers x 4 mul_mat levels) through ggml_graph_compute: what the launch-latency work is for
        try:
            from benchmarks.bench_inproc import chain_record
            line["dependent_chain"] = chain_record(torch, N, ggml, layers=32, iters=10, dev_index=local_rank)
            floor_ms = line["dependent_chain"]["weights_GB"] * 1e9 / (float(pk.get("hbm_gbs", 6650.0)) * 1e9) * 1e3
            line["dependent_chain"]["floor_ms_at_measured_hbm_peak"] = floor_ms
            line["dependent_chain"]["overhead_us_per_mul_mat_level"] = (line["dependent_chain"]["ms_per_compute"] - floor_ms) * 1e3 / line["dependent_chain"]["mul_mat_levels"]
        except Exception as ex:
            line["dependent_chain"] = {"error": repr(ex)}
        line["configs_note"] = ("one record per BASELINE.json configuration: ms per call of ggb_dev_mul_mat_batch over the listed nodes, achieved = "
                                "algorithmic bytes (W + x + y) or 2MNK flop / ms, frac of MEASURED_PEAKS.json (hbm_gbs / bf16_tflops burst), "
                                "frac_of_nominal of 8 TB/s / 2.25 PFLOP/s")
    # ---- N > 1: the row split INSIDE the shim (one process, the unchanged ggml_graph_compute, all N GPUs), measured by rank 0 while the
    #      other ranks sit in a CPU-side wait (a NCCL barrier would spin on their GPUs).  This is the end-to-end number that includes
    #      the exchange: host arena in, host arena out, every device's copy of dst complete. ----
    if world > 1 and not args.no_configs:
        if sym is not None:
            barrier()
            sym.close()
            sym = None
        del Wq, X, ws
        torch.cuda.empty_cache()
        store = dist.distributed_c10d._get_default_store()
        if rank == 0:
            from benchmarks.bench_inproc import inproc_records
            os.environ["GGB200_DEVICES"] = ",".join(str(i) for i in range(world))      # rank 0 may open the other ranks' GPUs now
            try:
                recs = list(inproc_records(torch, N, ggml, world, dev_index=local_rank, quick=True))
            except Exception as ex:                                  # report, never lose the headline line
                recs = [{"error": repr(ex)}]
            line["in_process_row_split"] = recs
            ring = [r for r in recs if r.get("config", "").startswith("configs[1]") and r.get("n_gpus") == world]
            if ring:
                r = ring[0]
                e2e = {"value": r["achieved"], "unit": "GB/s", "ms_per_step": r["ms_per_compute"], "h2d_bytes_per_step": r["h2d_bytes_per_compute"],
                       "d2h_bytes_per_step": r["d2h_bytes_per_compute"], "device_ms_per_step": r["device0_ms"],
                       "api": "ONE process: ggml_graph_compute over a 32-node graph of (%d x 4096) Q4_0 matrices on a host arena, row split over %d GPUs inside the shim "
                              "(resident weight slices; the exchange is fused into the GEMV epilogues and inside the timed region)" % (M_total, world)}
            store.set("ggb_inproc_done", "1")
        else:
            store.wait(["ggb_inproc_done"])
    if rank == 0:
        line["e2e"] = e2e
        if not args.no_cpu_baseline and world == 1:
            nth = os.cpu_count() or 1
            nodes = host_ring(8, M_LOCAL)
            sec = cpu_mul_mat_ring(nodes, nth, seconds_budget=15, repeats=5)
            line["cpu_baseline"] = {"value": 8 * alg_bytes_per_node(M_LOCAL, K) / sec / 1e9, "unit": "GB/s", "cores": nth, "kind": "port",
                                    "sample": "8 of the %d ring GEMVs, best of <=5 passes; C restatement of the reference algorithm (not .NET RyuJIT)" % RING}
        print(json.dumps(line), flush=True)
    if sym is not None:
        barrier()
        sym.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
