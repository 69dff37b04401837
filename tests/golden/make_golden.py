#!/usr/bin/env python3
"""Generates tests/golden/*.json.

An INDEPENDENT pure-Python restatement (numpy float32 scalars, Python's
round-half-even ``round``) of the reference routines on the mul_mat path, written
from /root/reference/GGMLSharp/Ggml.cs without looking at oracle/ggb_oracle.c.
The C oracle and the CUDA kernels are both checked against the vectors this
script emits, and the first three quantizer vectors are the hand-derived
known-answer tests of SURVEY.md Appendix B (asserted below).

The reference is C#; no .NET toolchain exists in this image, so these vectors
cannot be produced by running the reference itself (parity for the quantized
path is therefore "unpinned" in the sense of DESIGN.md section 3).

Run:  python tests/golden/make_golden.py      (needs only numpy)
"""
import json
import os
import struct
import warnings

import numpy as np

f32 = np.float32
HERE = os.path.dirname(os.path.abspath(__file__))


def rne(v):
    """Math.Round(double) -> MidpointRounding.ToEven.  Python's round() is the same."""
    return int(round(float(v)))


def q4_0_block(x):  # Ggml.cs:341-376
    amax, mx = f32(0), f32(0)
    for v in x:
        if amax < abs(v):
            amax, mx = abs(v), v
    d = f32(mx / f32(-8))
    idv = f32(f32(1) / d) if d != 0 else f32(0)
    qs = bytearray(16)
    for l in range(0, 32, 2):
        v0, v1 = f32(x[l] * idv), f32(x[l + 1] * idv)
        vi0 = min(15, rne(v0) + 8)
        vi1 = min(15, rne(v1) + 8)
        qs[l // 2] = (vi0 | (vi1 << 4)) & 0xFF
    return struct.pack("<f", d) + bytes(qs)


def q4_1_block(x):  # Ggml.cs:494-527
    mn, mx = f32(np.finfo(np.float32).max), f32(-np.finfo(np.float32).max)
    for v in x:
        if v < mn:
            mn = v
        if v > mx:
            mx = v
    d = f32(f32(mx - mn) / f32(15))
    idv = f32(f32(1) / d) if d != 0 else f32(0)
    qs = bytearray(16)
    for l in range(0, 32, 2):
        v0, v1 = f32(f32(x[l] - mn) * idv), f32(f32(x[l + 1] - mn) * idv)
        qs[l // 2] = (rne(v0) | (rne(v1) << 4)) & 0xFF
    return struct.pack("<ff", d, mn) + bytes(qs)


def q8_block(x, with_sums):  # Ggml.cs:738-761 / 786-822 with D2-D4 repaired
    amax = f32(0)
    for v in x:
        if amax < abs(v):
            amax = abs(v)
    d = f32(amax / f32(127))
    idv = f32(f32(1) / d) if d != 0 else f32(0)
    q = [rne(f32(v * idv)) for v in x]
    body = struct.pack("<32b", *q)
    if not with_sums:
        return struct.pack("<f", d) + body
    s0 = f32(d * f32(sum(q[:16])))
    s1 = f32(d * f32(sum(q[16:])))
    return struct.pack("<fff", d, s0, s1) + body


def quant_row(fn, x):
    return b"".join(fn(x[i:i + 32]) for i in range(0, len(x), 32))


def deq4_0_row(b):  # Ggml.cs:886-910
    out = []
    for i in range(0, len(b), 20):
        d = f32(struct.unpack_from("<f", b, i)[0])
        for j in range(16):
            vi = b[i + 4 + j]
            out += [f32(f32((vi & 15) - 8) * d), f32(f32((vi >> 4) - 8) * d)]
    return np.array(out, dtype=np.float32)


def deq4_1_row(b):  # Ggml.cs:962-987
    out = []
    for i in range(0, len(b), 24):
        d, m = (f32(v) for v in struct.unpack_from("<ff", b, i))
        for j in range(16):
            vi = b[i + 8 + j]
            out += [f32(f32(f32(vi & 15) * d) + m), f32(f32(f32(vi >> 4) * d) + m)]
    return np.array(out, dtype=np.float32)


def dot_f32(x, y):  # Ggml.cs:2631-2640
    s = 0.0
    for a, b in zip(x, y):
        s += float(f32(a * b))
    return f32(s)


def dot_f16(xh, yh):  # Ggml.cs:2642-2651  (inputs are np.float16 arrays)
    s = 0.0
    for a, b in zip(xh, yh):
        s += float(f32(f32(a) * f32(b)))
    return f32(s)


def dot_q4_0_q8_0(wb, xb, n):  # Ggml.cs:1136-1161, q8 signed
    sumf = f32(0)
    for i in range(n // 32):
        d0 = f32(struct.unpack_from("<f", wb, 20 * i)[0])
        d1 = f32(struct.unpack_from("<f", xb, 36 * i)[0])
        p1 = struct.unpack_from("<32b", xb, 36 * i + 4)
        sumi = 0
        for j in range(16):
            v0 = wb[20 * i + 4 + j]
            sumi += ((v0 & 15) - 8) * p1[2 * j] + ((v0 >> 4) - 8) * p1[2 * j + 1]
        sumf = f32(sumf + f32(f32(d0 * d1) * f32(sumi)))
    return sumf


def dot_q4_1_q8_1(wb, xb, n):  # Ggml.cs:1176-1200, q8 signed
    sumf = f32(0)
    for i in range(n // 32):
        d0, m0 = (f32(v) for v in struct.unpack_from("<ff", wb, 24 * i))
        d1 = f32(struct.unpack_from("<f", xb, 44 * i)[0])
        p1 = struct.unpack_from("<32b", xb, 44 * i + 12)
        for j in range(16):
            v0 = wb[24 * i + 8 + j]
            f0 = f32(f32(d0 * f32(v0 & 15)) + m0)
            f1 = f32(f32(d0 * f32(v0 >> 4)) + m0)
            f2 = f32(d1 * f32(p1[2 * j]))
            f3 = f32(d1 * f32(p1[2 * j + 1]))
            sumf = f32(sumf + f32(f32(f0 * f2) + f32(f1 * f3)))
    return sumf


def hexf(a):
    return np.ascontiguousarray(a, dtype=np.float32).tobytes().hex()


def main():
    rng = np.random.default_rng(20231018)
    blocks = {
        "A_ramp": np.arange(32, dtype=np.float32) - 16,
        "B_tenths": (f32(0.1) * np.arange(32, dtype=np.float32)).astype(np.float32),
        "C_zeros": np.zeros(32, dtype=np.float32),
        "D_tie_pos_first": np.array([3, -3] + [1] * 30, dtype=np.float32),
        "E_tie_neg_first": np.array([-3, 3] + [1] * 30, dtype=np.float32),
        "F_halves": (np.arange(32, dtype=np.float32) - 16) * f32(0.5) + f32(0.25),
        "G_const": np.full(32, 2.5, dtype=np.float32),
        "H_one_spike": np.array([0] * 31 + [-7.75], dtype=np.float32),
        "I_tiny": (rng.standard_normal(32) * 1e-30).astype(np.float32),
        "J_normal": rng.standard_normal(32).astype(np.float32),
        "K_uniform": rng.uniform(-1, 1, 32).astype(np.float32),
        "L_weights": (rng.standard_normal(32) * 0.02).astype(np.float32),
        "M_big": (rng.standard_normal(32) * 1e20).astype(np.float32),
    }
    kats = []
    for name, x in blocks.items():
        kats.append({"name": name, "x": hexf(x),
                     "q4_0": q4_0_block(x).hex(), "q4_1": q4_1_block(x).hex(),
                     "q8_0": q8_block(x, False).hex(), "q8_1": q8_block(x, True).hex(),
                     "deq4_0": hexf(deq4_0_row(q4_0_block(x))), "deq4_1": hexf(deq4_1_row(q4_1_block(x)))})
    # SURVEY.md Appendix B, derived by hand from Ggml.cs:334-377, 487-528
    assert kats[0]["q4_0"] == "00000040" + "0021224344656687" + "88a9aacbccedeeff"
    assert kats[0]["q4_1"] == "44440440" + "000080c1" + "00112233445566778899aabbccddeeff"
    assert kats[1]["q4_0"] == "6766c6be" + "88777766665555444433332222111100"
    assert kats[1]["q4_1"] == "6ea0533e" + "00000000" + "00112233445566778899aabbccddeeff"
    assert kats[2]["q4_0"] == "00000080" + "88" * 16
    assert kats[2]["q4_1"] == "00" * 24
    with open(os.path.join(HERE, "quant_kat.json"), "w") as f:
        json.dump({"source": "tests/golden/make_golden.py", "blocks": kats}, f, indent=1)

    # small mul_mat cases: W [M=6][K=96], X [N=3][K=96]
    M, K, N = 6, 96, 3
    W = (rng.standard_normal((M, K)) * 0.05).astype(np.float32)
    X = rng.standard_normal((N, K)).astype(np.float32)
    out = {"source": "tests/golden/make_golden.py", "M": M, "K": K, "N": N, "W": hexf(W), "X": hexf(X)}
    out["f32"] = hexf(np.array([[dot_f32(W[m], X[n]) for m in range(M)] for n in range(N)]))
    Wh, Xh = W.astype(np.float16), X.astype(np.float16)          # numpy casts RNE, like (Half)float
    out["W_f16"] = Wh.tobytes().hex()
    out["f16"] = hexf(np.array([[dot_f16(Wh[m], Xh[n]) for m in range(M)] for n in range(N)]))
    W40 = [quant_row(q4_0_block, W[m]) for m in range(M)]
    X80 = [quant_row(lambda b: q8_block(b, False), X[n]) for n in range(N)]
    out["W_q4_0"] = b"".join(W40).hex()
    out["q4_0"] = hexf(np.array([[dot_q4_0_q8_0(W40[m], X80[n], K) for m in range(M)] for n in range(N)]))
    W41 = [quant_row(q4_1_block, W[m]) for m in range(M)]
    X81 = [quant_row(lambda b: q8_block(b, True), X[n]) for n in range(N)]
    out["W_q4_1"] = b"".join(W41).hex()
    out["q4_1"] = hexf(np.array([[dot_q4_1_q8_1(W41[m], X81[n], K) for m in range(M)] for n in range(N)]))
    with open(os.path.join(HERE, "mul_mat_small.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote quant_kat.json (%d blocks), mul_mat_small.json" % len(kats))


# ---- neighbours of mul_mat: element-wise F32 ops, add_q_f32, cont(transpose) ----

def silu_table():  # Ggml.cs:1461-1471 + 2723-2726: table_silu_f16[i] = (Half)(f / (1.0f + MathF.Exp(-f))), f = (float)Half(i)
    # MathF.Exp is the C runtime's expf.  It is evaluated here as float32(exp(double)) -- the correctly rounded value, which is
    # what glibc's expf returns for every fp16-representable argument (numpy's vectorised float32 exp is 1 ulp off for three
    # of them, which moves three table entries, so it is NOT used).
    import math
    f = np.arange(1 << 16, dtype=np.uint16).view(np.float16).astype(np.float32)
    e = np.empty_like(f)
    for i, v in enumerate(f):
        v = float(v)
        if v != v:
            e[i] = np.nan
        elif -v > 700:
            e[i] = np.inf
        else:
            e[i] = f32(math.exp(-v))
    with np.errstate(over="ignore", invalid="ignore"):
        s = (f / (f32(1) + e)).astype(np.float32)
        return s.astype(np.float16)


def silu_row(x, table):  # Ggml.cs:2737-2746 (GGML_SILU_FP16 is defined in GGMLSharp.csproj:9)
    with np.errstate(over="ignore"):
        idx = np.asarray(x, dtype=np.float32).astype(np.float16).view(np.uint16)
    return table[idx].astype(np.float32)


def rms_norm_row(x):  # Ggml.cs:5895-5917
    s = 0.0
    for v in x:
        s += float(f32(v * v))
    mean = f32(s / len(x))
    scale = f32(f32(1) / np.sqrt(f32(mean + f32(1e-6))))
    return np.array([f32(v * scale) for v in x], dtype=np.float32)


def add_q_row(block_fn, deq_fn, bs, qrow, x):  # Ggml.cs:4893-4904
    k = len(x)
    w = np.concatenate([deq_fn(qrow[i * bs:(i + 1) * bs]) for i in range(k // 32)]).astype(np.float32)
    w = (w + np.asarray(x, dtype=np.float32)).astype(np.float32)
    return quant_row(block_fn, w)


def main_ops():
    rng = np.random.default_rng(20231019)
    R, K = 3, 64
    x = rng.standard_normal((R, K)).astype(np.float32)
    y = (rng.standard_normal((R, K)) * 3).astype(np.float32)
    x[0, :6] = [0.0, -0.0, 1e-8, -20.0, 20.0, 70000.0]           # silu: zeros, tiny, saturating, beyond fp16 range (-> inf)
    table = silu_table()
    out = {"source": "tests/golden/make_golden.py", "R": R, "K": K, "x": hexf(x), "y": hexf(y), "v": hexf(np.array([0.125 * 3.3], dtype=np.float32))}
    out["add"] = hexf((x + y).astype(np.float32))
    out["mul"] = hexf((x * y).astype(np.float32))
    out["scale"] = hexf((x * np.frombuffer(bytes.fromhex(out["v"]), dtype=np.float32)[0]).astype(np.float32))
    with np.errstate(invalid="ignore"):
        out["silu"] = hexf(np.stack([silu_row(r, table) for r in x]))
    out["silu_table_crc"] = int(np.bitwise_xor.reduce(table.view(np.uint16).astype(np.uint64) * (np.arange(1 << 16, dtype=np.uint64) * 2 + 1) % 1000003))
    out["silu_table_head"] = table[0x3000:0x3010].tobytes().hex()
    out["rms_norm"] = hexf(np.stack([rms_norm_row(r) for r in y]))
    w = (rng.standard_normal((R, K)) * 0.05).astype(np.float32)
    q40 = [quant_row(q4_0_block, r) for r in w]
    q41 = [quant_row(q4_1_block, r) for r in w]
    out["w"] = hexf(w)
    small = (x * f32(0.03)).astype(np.float32)
    small[0, 5] = 0.01
    out["addq_x"] = hexf(small)
    out["addq_q4_0"] = b"".join(add_q_row(q4_0_block, deq4_0_row, 20, q40[i], small[i]) for i in range(R)).hex()
    out["addq_q4_1"] = b"".join(add_q_row(q4_1_block, deq4_1_row, 24, q41[i], small[i]) for i in range(R)).hex()
    out["repeat_2x3"] = hexf(np.tile(x[:, :8], (2, 3)))           # ggml_repeat of [R][8] into [2R][24] (Ggml.cs:5371-5382)
    out["cont_transpose"] = hexf(np.ascontiguousarray(x.T))      # cont(transpose(x)): dst[i1][i0] = x[i0][i1]
    with open(os.path.join(HERE, "ops_small.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote ops_small.json")


# ---- sibling weight formats: Q4_2, Q5_0, Q5_1, Q8_0 ----
# fp16 scales are stored / read as IEEE binary16 bit patterns (defect D9 of oracle/ggb_oracle.c's header: the C# stores
# `(ushort)(Half)d`, a numeric cast); np.float16 casts are RNE like (Half)float.

def h16(v):
    return np.float32(v).astype(np.float16)


def trunc_int(v):  # C# (int)float for the in-range values these quantizers produce
    return int(np.trunc(float(v)))


def q4_2_block(x):  # Ggml.cs:554-589, 16 elements
    amax, mx = f32(0), f32(0)
    for v in x:
        if amax < abs(v):
            amax, mx = abs(v), v
    d = f32(mx / f32(-8))
    idv = f32(f32(1) / d) if d != 0 else f32(0)
    qs = bytearray(8)
    for l in range(0, 16, 2):
        vi0 = min(15, rne(f32(x[l] * idv)) + 8)
        vi1 = min(15, rne(f32(x[l + 1] * idv)) + 8)
        qs[l // 2] = (vi0 | (vi1 << 4)) & 0xFF
    return h16(d).tobytes() + bytes(qs)


def q5_0_block(x):  # Ggml.cs:614-652
    amax, mx = f32(0), f32(0)
    for v in x:
        if amax < abs(v):
            amax, mx = abs(v), v
    d = f32(mx / f32(-16))
    idv = f32(f32(1) / d) if d != 0 else f32(0)
    qs, qh = bytearray(16), 0
    for l in range(0, 32, 2):
        vi0 = min(31, trunc_int(f32(f32(x[l] * idv) + f32(16.5))))
        vi1 = min(31, trunc_int(f32(f32(x[l + 1] * idv) + f32(16.5))))
        qs[l // 2] = (vi0 & 15) | ((vi1 & 15) << 4)
        qh |= ((vi0 >> 4) & 1) << l
        qh |= ((vi1 >> 4) & 1) << (l + 1)
    return h16(d).tobytes() + struct.pack("<I", qh) + bytes(qs)


def q5_1_block(x):  # Ggml.cs:677-714
    mn, mx = f32(np.finfo(np.float32).max), f32(-np.finfo(np.float32).max)
    for v in x:
        if v < mn:
            mn = v
        if v > mx:
            mx = v
    d = f32(f32(mx - mn) / f32(31))
    idv = f32(f32(1) / d) if d != 0 else f32(0)
    qs, qh = bytearray(16), 0
    for l in range(0, 32, 2):
        vi0 = trunc_int(f32(f32(f32(x[l] - mn) * idv) + f32(0.5)))
        vi1 = trunc_int(f32(f32(f32(x[l + 1] - mn) * idv) + f32(0.5)))
        qs[l // 2] = (vi0 & 15) | ((vi1 & 15) << 4)
        qh |= ((vi0 >> 4) & 1) << l
        qh |= ((vi1 >> 4) & 1) << (l + 1)
    return h16(d).tobytes() + h16(mn).tobytes() + struct.pack("<I", qh) + bytes(qs)


def _h(b, off):
    return f32(np.frombuffer(b, dtype=np.float16, count=1, offset=off)[0])


def deq4_2_row(b):  # Ggml.cs:1001-1021
    out = []
    for i in range(0, len(b), 10):
        d = _h(b, i)
        for j in range(8):
            vi = b[i + 2 + j]
            out += [f32(f32((vi & 15) - 8) * d), f32(f32((vi >> 4) - 8) * d)]
    return np.array(out, dtype=np.float32)


def _q5(b, qoff, qhoff):
    qh = struct.unpack_from("<I", b, qhoff)[0]
    q = []
    for j in range(16):
        vi = b[qoff + j]
        q += [(vi & 15) | (((qh >> (2 * j)) & 1) << 4), (vi >> 4) | (((qh >> (2 * j + 1)) & 1) << 4)]
    return q


def deq5_0_row(b):  # Ggml.cs:1034-1060
    out = []
    for i in range(0, len(b), 22):
        d = _h(b, i)
        out += [f32(f32(q - 16) * d) for q in _q5(b, i + 6, i + 2)]
    return np.array(out, dtype=np.float32)


def deq5_1_row(b):  # Ggml.cs:1073-1100
    out = []
    for i in range(0, len(b), 24):
        d, m = _h(b, i), _h(b, i + 2)
        out += [f32(f32(f32(q) * d) + m) for q in _q5(b, i + 8, i + 4)]
    return np.array(out, dtype=np.float32)


def deq8_0_row(b):  # Ggml.cs:1111-1121, quants signed (D4)
    out = []
    for i in range(0, len(b), 36):
        d = f32(struct.unpack_from("<f", b, i)[0])
        out += [f32(f32(q) * d) for q in struct.unpack_from("<32b", b, i + 4)]
    return np.array(out, dtype=np.float32)


def dot_q4_2_q8_0(wb, xb, n):  # Ggml.cs:1216-1251
    sumf = f32(0)
    for i in range(n // 32):
        yd = f32(struct.unpack_from("<f", xb, 36 * i)[0])
        p = struct.unpack_from("<32b", xb, 36 * i + 4)
        for half in range(2):
            o = 10 * (2 * i + half)
            d = _h(wb, o)
            sumi = 0
            for j in range(8):
                v = wb[o + 2 + j]
                sumi += ((v & 15) - 8) * p[16 * half + 2 * j] + ((v >> 4) - 8) * p[16 * half + 2 * j + 1]
            sumf = f32(sumf + f32(f32(d * yd) * f32(sumi)))
    return sumf


def dot_q5_0_q8_0(wb, xb, n):  # Ggml.cs:1270-1298
    sumf = f32(0)
    for i in range(n // 32):
        yd = f32(struct.unpack_from("<f", xb, 36 * i)[0])
        p = struct.unpack_from("<32b", xb, 36 * i + 4)
        d = _h(wb, 22 * i)
        sxy = sum((q - 16) * pv for q, pv in zip(_q5(wb, 22 * i + 6, 22 * i + 2), p))
        sumf = f32(sumf + f32(f32(d * f32(sxy)) * yd))
    return sumf


def dot_q5_1_q8_1(wb, xb, n):  # Ggml.cs:1316-1345
    sumf = f32(0)
    for i in range(n // 32):
        yd, s0, s1 = (f32(v) for v in struct.unpack_from("<fff", xb, 44 * i))
        p = struct.unpack_from("<32b", xb, 44 * i + 12)
        d, m = _h(wb, 24 * i), _h(wb, 24 * i + 2)
        sxy = sum(q * pv for q, pv in zip(_q5(wb, 24 * i + 8, 24 * i + 4), p))
        sumf = f32(sumf + f32(f32(f32(d * f32(sxy)) * yd) + f32(m * f32(s0 + s1))))
    return sumf


def dot_q8_0_q8_0(wb, xb, n):  # Ggml.cs:1362-1377
    sumf = f32(0)
    for i in range(n // 32):
        xd = f32(struct.unpack_from("<f", wb, 36 * i)[0])
        yd = f32(struct.unpack_from("<f", xb, 36 * i)[0])
        a = struct.unpack_from("<32b", wb, 36 * i + 4)
        p = struct.unpack_from("<32b", xb, 36 * i + 4)
        sumf = f32(sumf + f32(f32(xd * yd) * f32(sum(u * v for u, v in zip(a, p)))))
    return sumf


def main_siblings():
    rng = np.random.default_rng(20231020)
    blocks = {
        "A_ramp": np.arange(32, dtype=np.float32) - 16,
        "B_tenths": (f32(0.1) * np.arange(32, dtype=np.float32)).astype(np.float32),
        "C_zeros": np.zeros(32, dtype=np.float32),
        "D_tie_pos_first": np.array([3, -3] + [1] * 30, dtype=np.float32),
        "E_tie_neg_first": np.array([-3, 3] + [1] * 30, dtype=np.float32),
        "F_halves": (np.arange(32, dtype=np.float32) - 16) * f32(0.5) + f32(0.25),
        "G_const": np.full(32, 2.5, dtype=np.float32),
        "H_one_spike": np.array([0] * 31 + [-7.75], dtype=np.float32),
        "J_normal": rng.standard_normal(32).astype(np.float32),
        "K_uniform": rng.uniform(-1, 1, 32).astype(np.float32),
        "L_weights": (rng.standard_normal(32) * 0.02).astype(np.float32),
        "N_small": (rng.standard_normal(32) * 1e-4).astype(np.float32),      # d is a subnormal half
    }
    q42 = lambda x: q4_2_block(x[:16]) + q4_2_block(x[16:])
    kats = []
    for name, x in blocks.items():
        b42, b50, b51, b80 = q42(x), q5_0_block(x), q5_1_block(x), q8_block(x, False)
        kats.append({"name": name, "x": hexf(x), "q4_2": b42.hex(), "q5_0": b50.hex(), "q5_1": b51.hex(), "q8_0": b80.hex(),
                     "deq4_2": hexf(deq4_2_row(b42)), "deq5_0": hexf(deq5_0_row(b50)),
                     "deq5_1": hexf(deq5_1_row(b51)), "deq8_0": hexf(deq8_0_row(b80))})
    # hand-derived from Ggml.cs:609-653 for x[l] = l - 16: max = -16 -> d = 1 (0x3c00), x*id + 16.5 truncates to l, so
    # qs[j] = (2j & 15) | ((2j+1 & 15) << 4) and the fifth bits are set for l >= 16
    assert kats[0]["q5_0"] == "003c" + "0000ffff" + "1032547698badcfe" * 2
    # Ggml.cs:672-714: min = -16, d = 1, the same quants, m = -16 (0xcc00)
    assert kats[0]["q5_1"] == "003c" + "00cc" + "0000ffff" + "1032547698badcfe" * 2
    # Ggml.cs:547-590 on the two halves: first max = -16 -> d = 2 (0x4000): round(-8, -7.5, .. -0.5) + 8 = 0 0 1 2 2 3 4 4 ...
    assert kats[0]["q4_2"][:4] == "0040"
    M, K, N = 6, 96, 3
    W = (rng.standard_normal((M, K)) * 0.05).astype(np.float32)
    X = rng.standard_normal((N, K)).astype(np.float32)
    out = {"source": "tests/golden/make_golden.py", "blocks": kats, "M": M, "K": K, "N": N, "W": hexf(W), "X": hexf(X)}
    X80 = [quant_row(lambda b: q8_block(b, False), X[n]) for n in range(N)]
    X81 = [quant_row(lambda b: q8_block(b, True), X[n]) for n in range(N)]
    for key, qfn, dot, xs in (("q4_2", q42, dot_q4_2_q8_0, X80), ("q5_0", q5_0_block, dot_q5_0_q8_0, X80),
                              ("q5_1", q5_1_block, dot_q5_1_q8_1, X81), ("q8_0", lambda b: q8_block(b, False), dot_q8_0_q8_0, X80)):
        Wq = [quant_row(qfn, W[m]) for m in range(M)]
        out["W_" + key] = b"".join(Wq).hex()
        out[key] = hexf(np.array([[dot(Wq[m], xs[n], K) for m in range(M)] for n in range(N)]))
    with open(os.path.join(HERE, "sibling_small.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote sibling_small.json (%d blocks)" % len(kats))


if __name__ == "__main__":
    main()
    main_ops()
    main_siblings()
