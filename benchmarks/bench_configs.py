#!/usr/bin/env python3
"""Every configuration of BASELINE.json on one B200, device-resident, CUDA-event timed.

  cfg0  F32  4096x4096 . 4096x1                      (GEMV, HBM)
  cfg1  Q4_0 4096x4096 . 4096x1                      (GEMV, HBM)  + quantize_row_q4_0 over the 4096x4096 source
  cfg2  Q4_1 / F16, 11008x4096 (w1/w3) and 4096x11008 (w2), N=1   (GEMV, HBM)
  cfg3  Q4_0 / F16 4096x4096 . 4096x512              (tcgen05 GEMM)
  cfg4  32 layers x {wq,wk,wv,wo 4096x4096; w1,w3 11008x4096; w2 4096x11008} Q4_0, N=1 decode and N=512 prompt, 1 GPU

GEMV configs cycle a ring of distinct weight matrices larger than 2x L2 and submit the whole ring as ONE batch (a graph
level of independent nodes), which is the steady state the bandwidth target applies to; "isolated" is one node per call.
Prints one JSON object per line; fractions are of MEASURED_PEAKS.json ("of measured") and of nominal 8 TB/s / 2.25 PF.
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


KERNELS = {"gemv": "k_gemv_fast<%s>", "gemm1": "k_gemm_q<%s> (one node: k_act_f16_dequant + one launch)", "gemm1_f16": "k_gemm_f16 (one node)",
           "gemmg": "k_gemm_q_grouped<%s>", "gemmg_f16": "k_gemm_f16_grouped"}
TN = {0: "f32", 1: "f16", 2: "q4_0", 3: "q4_1", 4: "q4_2", 6: "q5_0", 7: "q5_1", 8: "q8_0"}


class Runner:
    """Device-resident timing of batches of independent mul_mat nodes through ggb_dev_mul_mat_batch (CUDA events on the launching
    stream).  Shared by this script and by bench.py, which puts one record per BASELINE.json configuration into the driver's line."""

    def __init__(self, torch, N, L, dev, stream, hbm, tf):
        self.torch, self.N, self.L, self.dev, self.stream = torch, N, L, dev, stream
        self.sp = C.c_void_p(stream.cuda_stream)
        self.hbm, self.tf = hbm, tf

    def make_w(self, t, M, K):
        torch, N, L = self.torch, self.N, self.L
        rb = N.TYPE_SIZE[t] * (K // N.BLCK_SIZE[t])
        wf = torch.randn((M, K), device=self.dev) * 0.02
        if t == N.F32:
            return wf.view(torch.uint8).view(M, rb).clone(), rb
        w = torch.empty((M, rb), dtype=torch.uint8, device=self.dev)
        N.check(L.ggb_dev_quantize_rows(t, wf.data_ptr(), w.data_ptr(), M, K, self.sp))
        return w, rb

    def time_calls(self, fn, iters, warm=3):
        torch = self.torch
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        for _ in range(iters):
            fn()
        e1.record(self.stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    def run_nodes(self, shapes, Nn, label, iters, share_x=False, resident_rowexp=True):
        """shapes: list of (type, M, K); all submitted as one batch per call.  share_x: every node multiplies the same activations.
        resident_rowexp: the power-of-two row exponents of the tensor-core path are computed once, as the executor does for resident
        weights (ggb_dev_weight_rowexp), instead of inside every call."""
        torch, N, L = self.torch, self.N, self.L
        keep, mm = [], (N.ggb_dev_mm * len(shapes))()
        wbytes = flop = abytes = 0
        xs = {}
        for i, (t, M, K) in enumerate(shapes):
            w, rb = self.make_w(t, M, K)
            x = xs.get(K) if share_x else None
            if x is None:
                x = torch.randn((Nn, K), device=self.dev)
                xs[K] = x
            y = torch.zeros((Nn, M), device=self.dev)
            keep += [w, x, y]
            m = mm[i]
            m.type, m.M, m.K, m.N = t, M, K, Nn
            m.W, m.nb01, m.X, m.ldx_bytes, m.Y, m.ldy_bytes = w.data_ptr(), rb, x.data_ptr(), 4 * K, y.data_ptr(), 4 * M
            if resident_rowexp and Nn >= 16 and t not in (N.F32, N.F16):
                e = torch.empty((M,), dtype=torch.int32, device=self.dev)
                N.check(L.ggb_dev_weight_rowexp(t, w.data_ptr(), rb, M, K, e.data_ptr(), self.sp))
                keep.append(e)
                m.W_rowexp = e.data_ptr()
            wbytes += M * rb
            abytes += M * rb + 4 * K * Nn + 4 * M * Nn
            flop += 2.0 * M * K * Nn
        wsb = L.ggb_dev_workspace_bytes(mm, len(shapes))
        ws = torch.empty(wsb + 256, dtype=torch.uint8, device=self.dev)
        wsp = (ws.data_ptr() + 255) // 256 * 256
        ms = self.time_calls(lambda: N.check(L.ggb_dev_mul_mat_batch(mm, len(shapes), wsp, wsb, self.sp)), iters)
        t0 = shapes[0][0]
        out = {"config": label, "nodes": len(shapes), "N": Nn, "weight_MB": wbytes / 1e6, "ms": ms}
        if Nn < 16:
            gbs = abytes / (ms * 1e-3) / 1e9
            out.update({"bound": "hbm", "kernel": KERNELS["gemv"] % TN[t0], "achieved": gbs, "unit": "GB/s", "peak": self.hbm, "frac": gbs / self.hbm,
                        "frac_of_nominal": gbs / 8000.0, "alg_bytes": abytes})
        else:
            tfl = flop / (ms * 1e-3) / 1e12
            if t0 == N.F32:
                kern = "k_gemv_fast<f32> column passes (F32 weights stay off the tensor cores: 1e-5 contract)"
            else:
                key = ("gemm1" if len(shapes) == 1 else "gemmg") + ("_f16" if t0 == N.F16 else "")
                kern = KERNELS[key] % TN[t0] if "%s" in KERNELS[key] else KERNELS[key]
            out.update({"bound": "tensor", "kernel": kern, "achieved": tfl, "unit": "TFLOP/s", "peak": self.tf, "frac": tfl / self.tf,
                        "frac_of_nominal": tfl / 2250.0, "flop": flop, "includes": "activation staging (k_act_f16_dequant) + GEMM"})
        del keep
        torch.cuda.empty_cache()
        return out

    def baseline_records(self, iters=20, quick=False):
        """One record per BASELINE.json configuration on one GPU (configs[1]'s ring is bench.py's own headline).  Yields dicts."""
        N = self.N
        it = max(4, iters // 4) if quick else iters
        yield self.run_nodes([(N.F32, 4096, 4096)] * 5, 1, "configs[0] on the device: F32 4096x4096 GEMV, ring of 5 distinct matrices (336 MB > 2x L2)", it)
        yield self.run_nodes([(N.Q4_0, 4096, 4096)], 1, "configs[1] isolated: ONE Q4_0 4096x4096 GEMV per call (L2-warm, launch-latency bound: one launch, the GEMV quantizes its own activation row)", it * 4)
        for t in (N.Q4_1, N.F16):
            n_ring = 10 if t == N.Q4_1 else 4
            yield self.run_nodes([(t, 11008, 4096)] * n_ring, 1, "configs[2]: %s 11008x4096 (w1/w3) GEMV, ring of %d (> 2x L2)" % (TN[t], n_ring), it)
            yield self.run_nodes([(t, 4096, 11008)] * n_ring, 1, "configs[2]: %s 4096x11008 (w2, K=11008) GEMV, ring of %d (> 2x L2)" % (TN[t], n_ring), it)
        for t in (N.Q4_0, N.F16):
            yield self.run_nodes([(t, 4096, 4096)], 512, "configs[3] literal: ONE %s 4096x4096 . 4096x512 node per call" % TN[t], it)
            yield self.run_nodes([(t, 4096, 4096)] * 8, 512, "configs[3] batch: 8 independent %s 4096x4096 . 4096x512 nodes per call (a graph level)" % TN[t], max(4, it // 4))
        yield self.run_nodes([(N.Q4_0, 4096, 4096)] * 3, 512, "configs[3] as it occurs in a layer: wq/wk/wv, 3 x Q4_0 4096x4096 on ONE shared 4096x512 activation tensor", it, share_x=True)
        layer = [(N.Q4_0, 4096, 4096)] * 4 + [(N.Q4_0, 11008, 4096)] * 2 + [(N.Q4_0, 4096, 11008)]
        yield self.run_nodes(layer * 32, 1, "configs[4] on 1 GPU: Llama-7B-shaped stack, 32 layers x 7 Q4_0 matrices (4.05 GB), N=1 decode step (224 independent nodes)", max(3, it // 4))
        yield self.run_nodes(layer * 32, 512, "configs[4] on 1 GPU: the same 224 matrices, N=512 prompt step", 3 if quick else 4)


def main():
    import torch
    from ggmlsharp_b200 import native as N
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--sib-types", default="q4_2,q5_0,q5_1,q8_0", help="sibling formats to run in --only siblings (for targeted ncu captures)")
    ap.add_argument("--sib-parts", default="gemv,gemm,codecs")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    L = N.lib()
    N.check(L.ggb_init())
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sp = C.c_void_p(stream.cuda_stream)
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm, tf = float(pk.get("hbm_gbs", 6650.0)), float(pk.get("bf16_tflops", 1590.0))
    R = Runner(torch, N, L, dev, stream, hbm, tf)
    make_w, time_calls = R.make_w, R.time_calls

    def run_nodes(shapes, Nn, label, iters, share_x=False):
        print(json.dumps(R.run_nodes(shapes, Nn, label, iters, share_x)), flush=True)

    sel = set(a.only.split(",")) if a.only else None

    def want(k):
        return sel is None or k in sel

    if want("cfg0"):
        run_nodes([(N.F32, 4096, 4096)] * 5, 1, "cfg0 F32 4096x4096 GEMV, ring of 5 (336 MB)", a.iters)
        run_nodes([(N.F32, 4096, 4096)], 1, "cfg0 F32 4096x4096 GEMV, isolated (L2-warm)", a.iters)
    if want("cfg1"):
        run_nodes([(N.Q4_0, 4096, 4096)] * 32, 1, "cfg1 Q4_0 4096x4096 GEMV, ring of 32 (336 MB)", a.iters)
        run_nodes([(N.Q4_0, 4096, 4096)], 1, "cfg1 Q4_0 4096x4096 GEMV, isolated (L2-warm)", a.iters)
        src = torch.randn((8, 4096, 4096), device=dev) * 0.02                      # 8 distinct 64 MiB sources (> L2)
        dst = torch.empty((4096, 4096 // 32 * 20), dtype=torch.uint8, device=dev)
        it = [0]

        def q():
            N.check(L.ggb_dev_quantize_rows(N.Q4_0, src[it[0] % 8].data_ptr(), dst.data_ptr(), 4096, 4096, sp))
            it[0] += 1
        ms = time_calls(q, a.iters)
        by = 4096 * 4096 * 4 + 4096 * 4096 // 32 * 20
        print(json.dumps({"config": "cfg1 quantize_row_q4_0 over 4096x4096 F32 (bit-exact kernel)", "ms": ms, "GB/s": by / ms / 1e6,
                          "frac_of_measured_hbm": by / ms / 1e6 / hbm}), flush=True)
        del src, dst
    if want("codecs"):
        # the codec column on its own: 4 B read + 0.625 / 0.75 B written per element (quantize), the reverse (dequantize);
        # sources cycle through > 2x L2 of distinct data
        for rows in (4096, 11008):
            nsrc = max(2, (300 << 20) // (rows * 4096 * 4) + 1)
            src = torch.randn((nsrc, rows, 4096), device=dev) * 0.02
            for t in (N.Q4_0, N.Q4_1):
                rb = 4096 // 32 * N.TYPE_SIZE[t]
                dst = torch.empty((nsrc, rows, rb), dtype=torch.uint8, device=dev)
                it = [0]

                def q():
                    N.check(L.ggb_dev_quantize_rows(t, src[it[0] % nsrc].data_ptr(), dst[it[0] % nsrc].data_ptr(), rows, 4096, sp))
                    it[0] += 1
                ms = time_calls(q, a.iters)
                by = rows * 4096 * 4 + rows * rb
                print(json.dumps({"config": "quantize_row_%s over %dx4096 F32" % (TN[t], rows), "ms": ms, "GB/s": by / ms / 1e6,
                                  "frac_of_measured_hbm": by / ms / 1e6 / hbm}), flush=True)
                for _ in range(nsrc):
                    q()                                                                # every dst slot holds valid blocks
                back = torch.empty((nsrc, rows, 4096), device=dev)

                def dq():
                    N.check(L.ggb_dev_dequantize_rows(t, dst[it[0] % nsrc].data_ptr(), back[it[0] % nsrc].data_ptr(), rows, 4096, sp))
                    it[0] += 1
                ms = time_calls(dq, a.iters)
                print(json.dumps({"config": "dequantize_row_%s over %dx4096" % (TN[t], rows), "ms": ms, "GB/s": by / ms / 1e6,
                                  "frac_of_measured_hbm": by / ms / 1e6 / hbm}), flush=True)
                del dst, back
            del src
            torch.cuda.empty_cache()
    if want("siblings"):
        # the sibling weight formats through the same kernels (GEMV rings > 2x L2, codecs, tensor-core batch)
        sib = [t for t in (N.Q4_2, N.Q5_0, N.Q5_1, N.Q8_0) if TN[t] in a.sib_types.split(",")]
        parts = a.sib_parts.split(",")
        for t in sib:
            rb = 4096 // N.BLCK_SIZE[t] * N.TYPE_SIZE[t]
            n_ring = (336 << 20) // (4096 * rb) + 1
            n2 = max(2, n_ring * 4096 // 11008 + 1)
            if "gemv" in parts:
                run_nodes([(t, 4096, 4096)] * n_ring, 1, "sibling %s 4096x4096 GEMV, ring of %d" % (TN[t], n_ring), a.iters)
                run_nodes([(t, 11008, 4096)] * n2, 1, "sibling %s 11008x4096 (w1/w3) GEMV, ring of %d" % (TN[t], n2), a.iters)
                run_nodes([(t, 4096, 11008)] * n2, 1, "sibling %s 4096x11008 (w2, K=11008) GEMV, ring of %d" % (TN[t], n2), a.iters)
            if "gemm" not in parts:
                continue
            run_nodes([(t, 4096, 4096)] * 8, 512, "sibling %s 4096x4096 . 4096x512, batch of 8 nodes (%s)" % (TN[t], "in-kernel dequant"), max(4, a.iters // 4))
            run_nodes([(t, 4096, 4096)], 512, "sibling %s 4096x4096 . 4096x512, isolated" % TN[t], a.iters)
        rows = 11008
        nsrc = 2
        src = torch.randn((nsrc, rows, 4096), device=dev) * 0.02
        back = torch.empty((nsrc, rows, 4096), device=dev)
        for t in (sib if "codecs" in parts else []):
            rb = 4096 // N.BLCK_SIZE[t] * N.TYPE_SIZE[t]
            dst = torch.empty((nsrc, rows, rb), dtype=torch.uint8, device=dev)
            it = [0]

            def q():
                N.check(L.ggb_dev_quantize_rows(t, src[it[0] % nsrc].data_ptr(), dst[it[0] % nsrc].data_ptr(), rows, 4096, sp))
                it[0] += 1

            def dq():
                N.check(L.ggb_dev_dequantize_rows(t, dst[it[0] % nsrc].data_ptr(), back[it[0] % nsrc].data_ptr(), rows, 4096, sp))
                it[0] += 1
            by = rows * 4096 * 4 + rows * rb
            ms = time_calls(q, a.iters)
            print(json.dumps({"config": "quantize_row_%s over %dx4096 F32" % (TN[t], rows), "ms": ms, "GB/s": by / ms / 1e6,
                              "frac_of_measured_hbm": by / ms / 1e6 / hbm}), flush=True)
            ms = time_calls(dq, a.iters)
            print(json.dumps({"config": "dequantize_row_%s over %dx4096" % (TN[t], rows), "ms": ms, "GB/s": by / ms / 1e6,
                              "frac_of_measured_hbm": by / ms / 1e6 / hbm}), flush=True)
            del dst
        del src, back
        torch.cuda.empty_cache()
    if want("cfg2"):
        for t in (N.Q4_1, N.F16):
            n_ring = 10 if t == N.Q4_1 else 4
            run_nodes([(t, 11008, 4096)] * n_ring, 1, "cfg2 %s 11008x4096 (w1/w3) GEMV, ring of %d" % (TN[t], n_ring), a.iters)
            run_nodes([(t, 4096, 11008)] * n_ring, 1, "cfg2 %s 4096x11008 (w2, K=11008) GEMV, ring of %d" % (TN[t], n_ring), a.iters)
    if want("cfg3"):
        for t in (N.Q4_0, N.F16):
            run_nodes([(t, 4096, 4096)] * 8, 512, "cfg3 %s 4096x4096 . 4096x512, batch of 8 nodes" % TN[t], max(4, a.iters // 4))
            run_nodes([(t, 4096, 4096)], 512, "cfg3 %s 4096x4096 . 4096x512, isolated" % TN[t], a.iters)
        # how the shape occurs in a layer: wq / wk / wv multiply ONE activation tensor, which is then staged once
        run_nodes([(N.Q4_0, 4096, 4096)] * 3, 512, "cfg3 q4_0 wq/wk/wv: 3 x 4096x4096 on one shared 4096x512 activation tensor", a.iters, share_x=True)
        run_nodes([(N.Q4_0, 4096, 4096)] * 3, 512, "cfg3 q4_0 3 x 4096x4096, each with its own 4096x512 activations", a.iters)
    if want("cfg4"):
        layer = [(N.Q4_0, 4096, 4096)] * 4 + [(N.Q4_0, 11008, 4096)] * 2 + [(N.Q4_0, 4096, 11008)]
        run_nodes(layer * 32, 1, "cfg4 Llama-7B-shaped stack, 32 layers x 7 Q4_0 matrices (4.05 GB), N=1 decode step, 1 GPU", max(3, a.iters // 6))
        run_nodes(layer * 4, 512, "cfg4 Llama-7B-shaped layers (4 of 32: 28 matrices), N=512 prompt step, 1 GPU", 3)


if __name__ == "__main__":
    main()
