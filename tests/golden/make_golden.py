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


if __name__ == "__main__":
    main()
