"""Bulk throughput of the element-wise neighbours (ggb_ops.cu) -- they have no device-level entry point of their own, so this runs
one graph over a 4096 x 4096 F32 tensor (64 MB, larger than L2 with its result) through ggml_graph_compute and the kernel durations
are read from an ncu launch list:

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ops.csv python benchmarks/ops_throughput.py
    python benchmarks/ops_throughput.py --summarise gpurun_out/ops.csv > profiles/r02_ops_throughput.txt
"""
import csv
import sys

sys.path.insert(0, ".")

ROWS, COLS = 4096, 4096
# bytes each kernel moves per element: (kernel-name fragment, node, algorithmic bytes per element)
MOVES = [("k_binary_f32", "add / mul", 12), ("k_silu_f32", "silu (fp16 table)", 8), ("k_rms_norm_f32", "rms_norm (row held in registers between the passes)", 8),
         ("k_scale_f32", "scale (in place)", 8)]


def run():
    import numpy as np
    from ggmlsharp_b200 import ggml, native as N
    rng = np.random.default_rng(5)
    with ggml.Context(640 << 20) as c:
        x = c.tensor_from(N.F32, COLS, ROWS, data=rng.standard_normal((ROWS, COLS)).astype(np.float32))
        y = c.tensor_from(N.F32, COLS, ROWS, data=rng.standard_normal((ROWS, COLS)).astype(np.float32))
        f = c.tensor_from(N.F32, 1, data=np.array([0.5], np.float32))
        out = c.op("scale", c.op("rms_norm", c.op("silu", c.op("mul", c.op("add", x, y), y))), f)
        g = c.build_forward(out)
        for _ in range(3):
            c.graph_compute(g)
        print("ran", g.n_nodes, "nodes", float(ggml.tensor_f32(out).reshape(-1)[0]))


def summarise(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    ki, vi = H.index("Kernel Name"), H.index("Metric Value")
    n = ROWS * COLS
    print("# element-wise neighbours on a %d x %d F32 tensor (%.0f MB), kernel durations from %s (best of the launches listed)" % (ROWS, COLS, n * 4 / 1e6, path))
    for frag, what, bpe in MOVES:
        ds = [float(r[vi].replace(",", "")) for r in rows[hdr + 1:] if len(r) > vi and frag in r[ki]]
        if not ds:
            continue
        best = min(ds)
        print("%-18s %-52s %8.1f us  %6.0f GB/s algorithmic (%d B per element), %d launches" % (frag, what, best / 1e3, n * bpe / best, bpe, len(ds)))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--summarise":
        summarise(sys.argv[2])
    else:
        run()
