#!/usr/bin/env python3
"""SASS opcode evidence per kernel of libggb200.so (cuobjdump -sass): the Blackwell-native instructions B200_PROFILING.md lists
(UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA tensor load, UBLKCP = cp.async.bulk, IDP = dp4a, SYNCS = mbarrier).
Usage: python profiles/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "ggmlsharp_b200", "lib", "libggb200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
KEYS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTMAPF", "IDP", "SYNCS", "HMMA", "IMMA", "ACQBULK", "LDS", "STS", "LDG", "STG", "SHFL"]
cur, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        counts[cur][m.group(1)] += 1
        total[cur] += 1
dem = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print("libggb200.so: %d kernels (sm_100a).  Columns: total SASS instructions, then the opcodes that prove the Blackwell-native path." % len(counts))
print("%-112s %7s  %s" % ("kernel", "instrs", "  ".join(KEYS)))
for (k, c), name in zip(counts.items(), dem):
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"^void ggb::", "", name)
    name = re.sub(r"\(.*$", "", name)
    row = "  ".join("%*d" % (len(key), c.get(key, 0)) for key in KEYS)
    print("%-112s %7d  %s" % (name[:112], total[k], row))
agg = collections.Counter()
for c in counts.values():
    for key in KEYS:
        agg[key] += c.get(key, 0)
print("\nTOTAL over all kernels: " + ", ".join("%s %d" % (k, agg[k]) for k in KEYS))
