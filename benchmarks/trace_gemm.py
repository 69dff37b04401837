#!/usr/bin/env python3
"""Debug helper: clock64 timeline of the tcgen05 GEMM (per-CTA stamps written when GGB200_GEMM_TRACE holds a device pointer)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from ggmlsharp_b200 import native as N
t = {"q4_0": N.Q4_0, "f16": N.F16, "q4_1": N.Q4_1}[sys.argv[1] if len(sys.argv) > 1 else "q4_0"]
M = K = 4096; Nn = 512
dev = torch.device("cuda", 0); L = N.lib(); N.check(L.ggb_init())
st = torch.cuda.Stream(); torch.cuda.set_stream(st); sp = C.c_void_p(st.cuda_stream)
rb = N.TYPE_SIZE[t] * (K // N.BLCK_SIZE[t])
W = torch.empty((M, rb), dtype=torch.uint8, device=dev)
N.check(L.ggb_dev_quantize_rows(t, (torch.randn((M, K), device=dev) * 0.02).data_ptr(), W.data_ptr(), M, K, sp))
X = torch.randn((Nn, K), device=dev); Y = torch.zeros((Nn, M), device=dev)
trace = torch.zeros((256, 128), dtype=torch.int64, device=dev)
os.environ["GGB200_GEMM_TRACE"] = hex(trace.data_ptr())
mm = N.ggb_dev_mm(); mm.type, mm.M, mm.K, mm.N = t, M, K, Nn
mm.W, mm.nb01, mm.X, mm.ldx_bytes, mm.Y, mm.ldy_bytes = W.data_ptr(), rb, X.data_ptr(), 4 * K, Y.data_ptr(), 4 * M
wsb = L.ggb_dev_workspace_bytes(C.byref(mm), 1); ws = torch.empty(wsb + 256, dtype=torch.uint8, device=dev); wsp = (ws.data_ptr() + 255) // 256 * 256
for _ in range(3):
    N.check(L.ggb_dev_mul_mat_batch(C.byref(mm), 1, wsp, wsb, sp))
torch.cuda.synchronize()
tr = trace.cpu().numpy()
import numpy as np
st, en = tr[:128, 4], tr[:128, 5]
print('globaltimer (ns): CTA start spread %d, first start -> last end %d, per-CTA duration min/med/max %d/%d/%d' % (st.max() - st.min(), en.max() - st.min(), (en - st).min(), np.median(en - st), (en - st).max()))
print('per-CTA cycles (clock64) min/med/max: %d/%d/%d' % ((tr[:128,3]-tr[:128,0]).min(), np.median(tr[:128,3]-tr[:128,0]), (tr[:128,3]-tr[:128,0]).max()))
for cta in (0, 1, 64, 127):
    r = tr[cta]; t0 = r[0]
    ready = [(int(r[8 + 2 * k] - t0), int(r[9 + 2 * k] - t0)) for k in range(20)]
    per = [ready[k + 1][1] - ready[k][1] for k in range(19)]
    print("cta %3d: setup %d  first-ready %d  acc_full %d  epilogue_end %d  | B-ready deltas: %s" % (cta, r[1] - t0, ready[0][1], r[2] - t0, r[3] - t0, per))
    print("         A-ready vs B-ready (A-B): %s" % [a - b for a, b in ready])
    for base, nm, nit in ((48, "w4(q0)", 8),):
        for it in range(nit):
            st = [int(r[base + it * 6 + i] - t0) for i in range(6)]
            print("         dq %s ks=%2d: top %6d | raw-wait +%5d | lds+A_EMPTY +%5d | dequant+st +%4d | wait::st +%3d | arrive +%3d -> %6d" % (nm, it * 4, st[0], st[1] - st[0], st[2] - st[1], st[3] - st[2], st[4] - st[3], st[5] - st[4], st[5]))
    for k in range(8):
        st = [int(r[96 + 4 * k + i] - t0) for i in range(4)]
        print("         mma ks=%2d: issue-start %6d | 8 MMAs + commits + probes +%4d | residual waits +%4d" % (k + 8, st[0], st[1] - st[0], st[2] - st[1]))
    print("         mma    : A-ready   %s" % [a for a, b in ready[0:20:3]])
