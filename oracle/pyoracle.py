"""ctypes front-end of the CPU oracle (oracle/ggb_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, bench.py's cpu_baseline and
``--impl reference`` legs, and __graft_entry__.smoke().  Nothing under
ggmlsharp_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libggb_oracle.so")

F32, F16, Q4_0, Q4_1, Q4_2, Q5_0, Q5_1, Q8_0, Q8_1 = 0, 1, 2, 3, 4, 6, 7, 8, 9
TYPE_SIZE = {F32: 4, F16: 2, Q4_0: 20, Q4_1: 24, Q4_2: 10, Q5_0: 22, Q5_1: 24, Q8_0: 36, Q8_1: 44}
BLCK_SIZE = {F32: 1, F16: 1, Q4_0: 32, Q4_1: 32, Q4_2: 16, Q5_0: 32, Q5_1: 32, Q8_0: 32, Q8_1: 32}


def build(force=False):
    src = os.path.join(_HERE, "ggb_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


class _MM(C.Structure):
    _fields_ = [("type", C.c_int),
                ("ne0", C.c_int64 * 4), ("nb0", C.c_uint64 * 4), ("src0", C.c_void_p),
                ("ne1", C.c_int64 * 4), ("nb1", C.c_uint64 * 4), ("src1", C.c_void_p),
                ("ned", C.c_int64 * 4), ("nbd", C.c_uint64 * 4), ("dst", C.c_void_p),
                ("wdata", C.c_void_p), ("wsize", C.c_size_t), ("nth", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_mul_mat_2d.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int]
        _lib.orc_mul_mat.argtypes = [C.POINTER(_MM)]
        _lib.orc_mul_mat_work_size.argtypes = [C.c_int, C.c_int64]
        _lib.orc_mul_mat_work_size.restype = C.c_size_t
        _lib.orc_quantize_rows.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]
        _lib.orc_dequantize_rows.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]
        _lib.orc_f32_to_f16_row.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        _lib.orc_f16_to_f32_row.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        for n in ("orc_vec_dot_f32", "orc_vec_dot_f16", "orc_vec_dot_q4_0_q8_0", "orc_vec_dot_q4_1_q8_1",
                  "orc_vec_dot_q4_2_q8_0", "orc_vec_dot_q5_0_q8_0", "orc_vec_dot_q5_1_q8_1", "orc_vec_dot_q8_0_q8_0"):
            getattr(_lib, n).argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def row_bytes(t, k):
    assert k % BLCK_SIZE[t] == 0
    return TYPE_SIZE[t] * (k // BLCK_SIZE[t])


def quantize_rows(t, x):
    """x: float32 [nrows, k] -> uint8 [nrows, row_bytes] in the reference's block layout."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    if x.ndim == 1:
        x = x[None, :]
    nrows, k = x.shape
    out = np.zeros((nrows, row_bytes(t, k)), dtype=np.uint8)
    rc = lib().orc_quantize_rows(t, _p(x), _p(out), nrows, k)
    assert rc == 0, rc
    return out


def dequantize_rows(t, q, k):
    q = np.ascontiguousarray(q, dtype=np.uint8)
    if q.ndim == 1:
        q = q[None, :]
    nrows = q.shape[0]
    out = np.zeros((nrows, k), dtype=np.float32)
    rc = lib().orc_dequantize_rows(t, _p(q), _p(out), nrows, k)
    assert rc == 0, rc
    return out


def f32_to_f16(x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.zeros(x.shape, dtype=np.uint16)
    lib().orc_f32_to_f16_row(_p(x), _p(out), x.size)
    return out


def f16_to_f32(h):
    h = np.ascontiguousarray(h, dtype=np.uint16)
    out = np.zeros(h.shape, dtype=np.float32)
    lib().orc_f16_to_f32_row(_p(h), _p(out), h.size)
    return out


def encode_weights(t, w):
    """float32 [M, K] -> bytes of a src0 tensor of type t as uint8 [M, row_bytes]."""
    w = np.ascontiguousarray(w, dtype=np.float32)
    if t == F32:
        return w.view(np.uint8).reshape(w.shape[0], -1).copy()
    if t == F16:
        return f32_to_f16(w).view(np.uint8).reshape(w.shape[0], -1).copy()
    return quantize_rows(t, w)


def mul_mat_2d(t, wbytes, M, K, x, nth=1):
    """dst[N][M] of ggml_mul_mat(W, X) for contiguous W (M rows of K, type t) and X float32 [N, K]."""
    wbytes = np.ascontiguousarray(wbytes)
    x = np.ascontiguousarray(x, dtype=np.float32)
    if x.ndim == 1:
        x = x[None, :]
    N = x.shape[0]
    assert x.shape[1] == K and wbytes.nbytes == M * row_bytes(t, K)
    y = np.zeros((N, M), dtype=np.float32)
    rc = lib().orc_mul_mat_2d(t, _p(wbytes), M, K, _p(x), N, _p(y), nth)
    assert rc == 0, rc
    return y


def mul_mat_strided(t, src0, ne0, nb0, src1, ne1, nb1, dst, ned, nbd, nth=1):
    """Full ggml_compute_forward_mul_mat with explicit ne/nb (numpy buffers as backing store)."""
    p = _MM()
    p.type = t
    p.nth = nth
    for i in range(4):
        p.ne0[i], p.nb0[i] = ne0[i], nb0[i]
        p.ne1[i], p.nb1[i] = ne1[i], nb1[i]
        p.ned[i], p.nbd[i] = ned[i], nbd[i]
    p.src0, p.src1, p.dst = src0.ctypes.data, src1.ctypes.data, dst.ctypes.data
    nel1 = int(np.prod(ne1))
    ws = lib().orc_mul_mat_work_size(t, nel1)
    wbuf = np.zeros(max(ws, 1), dtype=np.uint8)
    p.wdata, p.wsize = wbuf.ctypes.data, ws
    rc = lib().orc_mul_mat(C.byref(p))
    assert rc == 0, rc
    return dst


def vec_dot(name, n, x, y):
    s = np.zeros(1, dtype=np.float32)
    getattr(lib(), "orc_vec_dot_" + name)(n, _p(s), _p(np.ascontiguousarray(x)), _p(np.ascontiguousarray(y)))
    return float(s[0])


# ---- neighbours of mul_mat: element-wise F32 ops, add_q_f32, cont(transpose) ----

def _bin(name, x, y):
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.ascontiguousarray(y, dtype=np.float32)
    assert x.shape == y.shape
    z = np.zeros_like(x)
    f = getattr(lib(), name)
    f.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    f(x.size, _p(x), _p(y), _p(z))
    return z


def add_f32(x, y):
    return _bin("orc_add_f32", x, y)


def mul_f32(x, y):
    return _bin("orc_mul_f32", x, y)


def scale_f32(x, v):
    y = np.array(x, dtype=np.float32, copy=True, order="C")
    f = lib().orc_scale_f32
    f.argtypes = [C.c_int64, C.c_void_p, C.c_float]
    f(y.size, _p(y), float(np.float32(v)))
    return y


def silu_f32(x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.zeros_like(x)
    f = lib().orc_silu_f32
    f.argtypes = [C.c_int64, C.c_void_p, C.c_void_p]
    f(x.size, _p(x), _p(y))
    return y


def silu_table():
    t = np.zeros(1 << 16, dtype=np.uint16)
    f = lib().orc_silu_table
    f.argtypes = [C.c_void_p]
    f(_p(t))
    return t


def rms_norm_f32(x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    x2 = x.reshape(-1, x.shape[-1])
    y = np.zeros_like(x2)
    f = lib().orc_rms_norm_f32
    f.argtypes = [C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
    f(x2.shape[0], x2.shape[1], _p(x2), x2.shape[1], _p(y), x2.shape[1])
    return y.reshape(x.shape)


def add_q_f32(t, q, x):
    """q: uint8 [nrows, row_bytes] of type t, x: float32 [nrows, k] -> uint8 [nrows, row_bytes] (dequantize + x, requantized)."""
    q = np.ascontiguousarray(q, dtype=np.uint8)
    x = np.ascontiguousarray(x, dtype=np.float32)
    nrows, k = x.shape
    assert q.shape == (nrows, row_bytes(t, k))
    out = np.zeros_like(q)
    f = lib().orc_add_q_f32
    f.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]
    rc = f(t, _p(q), _p(x), _p(out), nrows, k)
    assert rc == 0, rc
    return out


def dup_f32_strided(buf, ne, nb):
    """Contiguous F32 copy of the strided view (ne, nb in bytes) over `buf` -- ggml_cont of a transposed / permuted tensor."""
    ne4 = (C.c_int64 * 4)(*(list(ne) + [1] * (4 - len(ne))))
    nb4 = (C.c_uint64 * 4)(*(list(nb) + [0] * (4 - len(nb))))
    n = int(np.prod(list(ne)))
    out = np.zeros(n, dtype=np.float32)
    f = lib().orc_dup_f32_strided
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    f(_p(buf), ne4, nb4, _p(out))
    return out


def repeat_f32(x, nr, nc):
    """ggml_repeat of the 2-D float32 x [nr0, nc0] into [nr, nc]."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.zeros((nr, nc), dtype=np.float32)
    f = lib().orc_repeat_f32
    f.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int64]
    f(_p(x), x.shape[1], x.shape[0], _p(out), nc, nr)
    return out
