import ctypes as C, json, os, sys
sys.path.insert(0, ".")
import torch
from ggmlsharp_b200 import native as N
from benchmarks.bench_configs import Runner
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
L = N.lib(); N.check(L.ggb_init())
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
R = Runner(torch, N, L, dev, stream, 6535.4, 1640.6)
layer = [(N.Q4_0, 4096, 4096)] * 4 + [(N.Q4_0, 11008, 4096)] * 2 + [(N.Q4_0, 4096, 11008)]
for label, shapes in (("cfg4 prompt 4 layers", layer * 4), ("cfg3 batch 8", [(N.Q4_0, 4096, 4096)] * 8)):
    N.check(L.ggb_set_kernel_timing(0))
    r0 = R.run_nodes(shapes, 512, label, 6)
    L.ggb_reset_stats()
    N.check(L.ggb_set_kernel_timing(1))
    r1 = R.run_nodes(shapes, 512, label, 6)
    st = N.stats()
    N.check(L.ggb_set_kernel_timing(0))
    print(label, "total us", round(r0["ms"] * 1e3, 1), "| with brackets", round(r1["ms"] * 1e3, 1), "| GEMM kernels only us/call", round(st.timed_kernel_ms / 9 * 1e3, 1), "launches", st.timed_kernel_launches)
