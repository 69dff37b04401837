#!/usr/bin/env python3
"""Summarise an .ncu-rep (read with `ncu -i`, no GPU needed) into the text kept under profiles/.

usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt
"""
import csv
import io
import subprocess
import sys
from collections import Counter

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.max"]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    print("# %s" % rep)
    for r in rows[2:]:
        print("\n== kernel: %s  (id %s)" % (r[hdr.index("Kernel Name")], r[0]))
        for k in KEYS:
            if k in hdr:
                print("  %-72s %16s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "sass"]))))
    hidx = [i for i, r in enumerate(src) if r and r[0] == "Address"]
    if not hidx:
        return
    h = src[hidx[0]]
    body = src[hidx[0] + 1:(hidx[1] - 1 if len(hidx) > 1 else len(src))]
    print("\n== first launch: warp-stall samples by reason (source page, summed over SASS)")
    agg = {}
    for i, c in enumerate(h):
        if c.startswith("stall_") and "Not Issued" not in c:
            agg[c] = sum(int(r[i]) for r in body if len(r) > i and r[i].isdigit())
    tot = sum(agg.values()) or 1
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]:
        print("  %-28s %8d  %5.1f%%" % (k, v, 100.0 * v / tot))
    ia, isrc = h.index("Instructions Executed"), h.index("Source")
    ops = Counter()
    for r in body:
        t = r[isrc].split()
        if t:
            ops[(t[1] if t[0].startswith("@") else t[0]).split(".")[0]] += int(r[ia]) if r[ia].isdigit() else 0
    n = sum(ops.values()) or 1
    print("\n== first launch: executed warp-instructions by opcode (total %d)" % n)
    print("  " + ", ".join("%s %.1f%%" % (k, 100.0 * v / n) for k, v in ops.most_common(16)))


if __name__ == "__main__":
    main()
