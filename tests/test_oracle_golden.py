"""The C oracle against the committed golden vectors (tests/golden/make_golden.py,
an independent pure-Python restatement of Ggml.cs) and SURVEY.md Appendix B KATs."""
import json
import os

import numpy as np
import pytest

from oracle import pyoracle as orc

G = os.path.join(os.path.dirname(__file__), "golden")


def _f32(h):
    return np.frombuffer(bytes.fromhex(h), dtype=np.float32)


@pytest.fixture(scope="module")
def kats():
    with open(os.path.join(G, "quant_kat.json")) as f:
        return json.load(f)["blocks"]


@pytest.mark.parametrize("t,key", [(orc.Q4_0, "q4_0"), (orc.Q4_1, "q4_1"), (orc.Q8_0, "q8_0"), (orc.Q8_1, "q8_1")])
def test_quantize_kats_bit_exact(kats, t, key):
    for b in kats:
        got = orc.quantize_rows(t, _f32(b["x"])).tobytes().hex()
        assert got == b[key], (b["name"], key)


def test_appendix_b_literals(kats):
    # SURVEY.md Appendix B (derived by hand from Ggml.cs:334-377, 487-528)
    x = np.arange(32, dtype=np.float32) - 16
    assert orc.quantize_rows(orc.Q4_0, x).tobytes().hex() == "000000400021224344656687" + "88a9aacbccedeeff"
    # ties-to-even, not C roundf: with roundf the bytes would be 00 11 22 33 ...
    assert orc.quantize_rows(orc.Q4_0, x).tobytes().hex() != "00000040" + "0011223344556677" + "98a9bacbdcedfeff"
    z = np.zeros(32, dtype=np.float32)
    assert orc.quantize_rows(orc.Q4_0, z).tobytes().hex() == "00000080" + "88" * 16      # d = -0.0f


@pytest.mark.parametrize("t,key", [(orc.Q4_0, "deq4_0"), (orc.Q4_1, "deq4_1")])
def test_dequantize_kats_bit_exact(kats, t, key):
    src = {"deq4_0": "q4_0", "deq4_1": "q4_1"}[key]
    for b in kats:
        q = np.frombuffer(bytes.fromhex(b[src]), dtype=np.uint8)
        got = orc.dequantize_rows(t, q, 32)
        assert got.tobytes().hex() == b[key], b["name"]


def test_f16_conversions_match_numpy():
    rng = np.random.default_rng(7)
    x = np.concatenate([rng.standard_normal(20000).astype(np.float32) * s for s in (1e-8, 1e-5, 1e-3, 1, 100, 7e4)]
                       + [np.array([0, -0.0, 65504, 65519.99, 65520, 2.0**-24, 2.0**-25, 2.0**-25 * 1.0001,
                                    6.1e-5, np.inf, -np.inf], dtype=np.float32)])
    with np.errstate(over="ignore"):
        want = x.astype(np.float16).view(np.uint16)
    np.testing.assert_array_equal(orc.f32_to_f16(x), want)
    h = np.arange(65536, dtype=np.uint16)
    got = orc.f16_to_f32(h)
    ref = h.view(np.float16).astype(np.float32)
    np.testing.assert_array_equal(got.view(np.uint32)[~np.isnan(ref)], ref.view(np.uint32)[~np.isnan(ref)])


@pytest.mark.parametrize("t,key,wkey", [(orc.F32, "f32", "W"), (orc.F16, "f16", "W_f16"),
                                        (orc.Q4_0, "q4_0", "W_q4_0"), (orc.Q4_1, "q4_1", "W_q4_1")])
@pytest.mark.parametrize("nth", [1, 4])
def test_mul_mat_small_golden_bit_exact(t, key, wkey, nth):
    with open(os.path.join(G, "mul_mat_small.json")) as f:
        g = json.load(f)
    M, K, N = g["M"], g["K"], g["N"]
    wb = np.frombuffer(bytes.fromhex(g[wkey]), dtype=np.uint8)
    X = _f32(g["X"]).reshape(N, K)
    # the weight bytes themselves: oracle encode == golden encode
    np.testing.assert_array_equal(orc.encode_weights(t, _f32(g["W"]).reshape(M, K)).ravel(), wb)
    got = orc.mul_mat_2d(t, wb, M, K, X, nth=nth)
    assert got.tobytes().hex() == g[key]


def test_quantize_roundtrip_properties():
    rng = np.random.default_rng(3)
    x = rng.standard_normal((64, 256)).astype(np.float32)
    for t, tol in ((orc.Q4_0, 1 / 8), (orc.Q4_1, 0.5 / 15)):  # Q4_0: clamp at +8 -> 15 costs up to |d|
        q = orc.quantize_rows(t, x)
        y = orc.dequantize_rows(t, q, 256)
        xb = x.reshape(64, 8, 32)
        rngs = np.abs(xb).max(-1) if t == orc.Q4_0 else (xb.max(-1) - xb.min(-1))
        err = np.abs(y.reshape(64, 8, 32) - xb).max(-1)
        assert (err <= rngs * tol * 1.01 + 1e-12).all()


# ---- neighbours of mul_mat: element-wise F32 ops, add_q_f32, cont(transpose) ----

@pytest.fixture(scope="module")
def ops():
    with open(os.path.join(G, "ops_small.json")) as f:
        return json.load(f)


def test_elementwise_ops_bit_exact(ops):
    R, K = ops["R"], ops["K"]
    x, y = _f32(ops["x"]).reshape(R, K), _f32(ops["y"]).reshape(R, K)
    v = _f32(ops["v"])[0]
    assert orc.add_f32(x, y).tobytes().hex() == ops["add"]
    assert orc.mul_f32(x, y).tobytes().hex() == ops["mul"]
    assert orc.scale_f32(x, v).tobytes().hex() == ops["scale"]
    assert orc.silu_f32(x).tobytes().hex() == ops["silu"]
    assert orc.rms_norm_f32(y).tobytes().hex() == ops["rms_norm"]


def test_silu_table_matches_independent_restatement(ops):
    t = orc.silu_table()
    crc = int(np.bitwise_xor.reduce(t.astype(np.uint64) * (np.arange(1 << 16, dtype=np.uint64) * 2 + 1) % 1000003))
    assert crc == ops["silu_table_crc"]
    assert t[0x3000:0x3010].tobytes().hex() == ops["silu_table_head"]
    # spot values: silu(0) = 0, silu(1) = 0.731..., large x -> x, very negative -> -0
    h = lambda f: int(np.float16(f).view(np.uint16))
    assert t[h(0.0)] == 0 and t[h(1.0)] == h(0.7310586) and t[h(20.0)] == h(20.0) and t[h(-30.0)] == 0x8000


def test_add_q_f32_bit_exact(ops):
    R, K = ops["R"], ops["K"]
    w, x = _f32(ops["w"]).reshape(R, K), _f32(ops["addq_x"]).reshape(R, K)
    for t, key in ((orc.Q4_0, "addq_q4_0"), (orc.Q4_1, "addq_q4_1")):
        q = orc.quantize_rows(t, w)
        assert orc.add_q_f32(t, q, x).tobytes().hex() == ops[key]


def test_cont_transpose(ops):
    R, K = ops["R"], ops["K"]
    x = _f32(ops["x"]).reshape(R, K).copy()
    # ggml_transpose: ne = [R, K] (ne0 = R), nb = [K*4, 4]
    got = orc.dup_f32_strided(x, (R, K), (K * 4, 4))
    assert got.tobytes().hex() == ops["cont_transpose"]


def test_repeat(ops):
    R, K = ops["R"], ops["K"]
    x = _f32(ops["x"]).reshape(R, K)[:, :8].copy()
    assert orc.repeat_f32(x, 2 * R, 24).tobytes().hex() == ops["repeat_2x3"]


# ---- sibling weight formats: Q4_2, Q5_0, Q5_1, Q8_0 ----

SIB = [(orc.Q4_2, "q4_2"), (orc.Q5_0, "q5_0"), (orc.Q5_1, "q5_1"), (orc.Q8_0, "q8_0")]


@pytest.fixture(scope="module")
def sib():
    with open(os.path.join(G, "sibling_small.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("t,key", SIB)
def test_sibling_quantize_dequantize_kats_bit_exact(sib, t, key):
    for b in sib["blocks"]:
        assert orc.quantize_rows(t, _f32(b["x"])).tobytes().hex() == b[key], (b["name"], key)
        q = np.frombuffer(bytes.fromhex(b[key]), dtype=np.uint8)
        assert orc.dequantize_rows(t, q, 32).tobytes().hex() == b["de" + key], (b["name"], key)


def test_sibling_hand_derived_kats():
    # derived by hand from Ggml.cs:609-653 / 672-714 for x[l] = l - 16 (see make_golden.py::main_siblings)
    x = np.arange(32, dtype=np.float32) - 16
    assert orc.quantize_rows(orc.Q5_0, x).tobytes().hex() == "003c" + "0000ffff" + "1032547698badcfe" * 2
    assert orc.quantize_rows(orc.Q5_1, x).tobytes().hex() == "003c" + "00cc" + "0000ffff" + "1032547698badcfe" * 2
    z = np.zeros(32, dtype=np.float32)
    assert orc.quantize_rows(orc.Q5_0, z).tobytes().hex() == "0080" + "ffffffff" + "00" * 16     # d = -0.0, every quant 16
    assert orc.quantize_rows(orc.Q4_2, z).tobytes().hex() == ("0080" + "88" * 8) * 2


@pytest.mark.parametrize("t,key", SIB)
@pytest.mark.parametrize("nth", [1, 3])
def test_sibling_mul_mat_small_golden_bit_exact(sib, t, key, nth):
    M, K, N = sib["M"], sib["K"], sib["N"]
    wb = np.frombuffer(bytes.fromhex(sib["W_" + key]), dtype=np.uint8)
    X = _f32(sib["X"]).reshape(N, K)
    np.testing.assert_array_equal(orc.encode_weights(t, _f32(sib["W"]).reshape(M, K)).ravel(), wb)
    got = orc.mul_mat_2d(t, wb, M, K, X, nth=nth)
    assert got.tobytes().hex() == sib[key]


@pytest.mark.parametrize("t,key", SIB)
def test_sibling_dot_equals_dequantized_dot(t, key):
    """Independent of any golden file: the quantized dot equals sum(dequant(w) * dequant(q8(x))) up to float reassociation."""
    rng = np.random.default_rng(11)
    M, K, N = 5, 256, 2
    W = (rng.standard_normal((M, K)) * 0.02).astype(np.float32)
    X = rng.standard_normal((N, K)).astype(np.float32)
    wq = orc.quantize_rows(t, W)
    got = orc.mul_mat_2d(t, wq, M, K, X)
    wd = orc.dequantize_rows(t, wq, K).astype(np.float64)
    xd = orc.dequantize_rows(orc.Q8_0, orc.quantize_rows(orc.Q8_0, X), K).astype(np.float64)
    want = xd @ wd.T
    assert np.linalg.norm(got - want) / np.linalg.norm(want) < 2e-6


@pytest.mark.parametrize("t", [orc.Q4_0, orc.Q4_2, orc.Q5_0])
def test_symmetric_formats_requantize_to_themselves(t):
    """Size-independent property of the symmetric formats: the dequantized block's largest magnitude is exactly -8d (-16d), so
    quantizing it again reproduces scale and quants bit for bit.  (Not true for Q4_1 / Q5_1 / Q8_0, whose scale is recomputed
    from a rounded difference / product.)"""
    rng = np.random.default_rng(17 + t)
    x = np.concatenate([(rng.standard_normal((512, 256)) * s).astype(np.float32) for s in (0.02, 1.0, 300.0)])
    x[0, :32] = 0
    q = orc.quantize_rows(t, x)
    again = orc.quantize_rows(t, orc.dequantize_rows(t, q, 256))
    assert np.array_equal(q, again)
