"""Pythonic veneer over the host mirror (libggml_host.so): the call sequence a GGMLSharp program makes
-- ggml_init, ggml_new_tensor_*, write tensor->data, ggml_mul_mat, ggml_build_forward,
ggml_graph_compute, read dst->data -- with numpy views onto the pinned arena."""
import ctypes as C

import numpy as np

from . import native as N


def tensor_bytes(t):
    """uint8 view of a tensor's data inside the arena (no copy)."""
    nbytes = int(N.host().ggml_nbytes(t))
    return np.ctypeslib.as_array(C.cast(t.contents.data, C.POINTER(C.c_uint8)), shape=(nbytes,))


def tensor_f32(t):
    assert t.contents.type == N.F32
    ne = list(t.contents.ne)
    return tensor_bytes(t).view(np.float32).reshape(ne[3], ne[2], ne[1], ne[0])


class Context:
    def __init__(self, mem_size, mem_buffer=None, no_alloc=False):
        p = N.ggml_init_params()
        p.mem_size = mem_size
        self._keep = mem_buffer
        p.mem_buffer = mem_buffer.ctypes.data if mem_buffer is not None else None
        p.no_alloc = 1 if no_alloc else 0
        self.ctx = N.host().ggml_init(p)
        if not self.ctx:
            raise N.GgbError(N.host().ggml_host_last_status(), N.lib().ggb_last_error().decode())

    def free(self):
        if self.ctx:
            N.host().ggml_free(self.ctx)
            self.ctx = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.free()

    def new_tensor(self, t, *ne):
        fn = getattr(N.host(), "ggml_new_tensor_%dd" % len(ne))
        r = fn(self.ctx, t, *ne)
        if not r:
            raise MemoryError("ggml_new_tensor: context pool exhausted")
        return r

    def tensor_from(self, t, *ne, data=None):
        r = self.new_tensor(t, *ne)
        if data is not None:
            b = tensor_bytes(r)
            src = np.ascontiguousarray(data).view(np.uint8).ravel()
            assert src.size == b.size, (src.size, b.size)
            b[:] = src
        return r

    def mul_mat(self, a, b):
        r = N.host().ggml_mul_mat(self.ctx, a, b)
        if not r:
            raise N.GgbError(N.host().ggml_host_last_status(), "ggml_mul_mat rejected the operands")
        return r

    def cpy(self, a, b):
        r = N.host().ggml_cpy(self.ctx, a, b)
        if not r:
            raise N.GgbError(N.host().ggml_host_last_status(), "ggml_cpy rejected the operands")
        return r

    def op(self, name, *operands):
        """ggml_add / ggml_mul / ggml_scale / ggml_repeat / ggml_silu / ggml_rms_norm / ggml_cont / ggml_transpose / *_inplace."""
        r = getattr(N.host(), "ggml_" + name)(self.ctx, *operands)
        if not r:
            raise N.GgbError(N.host().ggml_host_last_status(), "ggml_%s rejected the operands" % name)
        return r

    def build_forward(self, t):
        g = N.ggml_cgraph()
        N.host().ggml_build_forward_into(C.byref(g), t)
        return g

    def graph_compute(self, g):
        N.host().ggml_graph_compute(self.ctx, C.byref(g))
        rc = N.host().ggml_host_last_status()
        if rc < 0:
            raise N.GgbError(rc, N.lib().ggb_last_error().decode(errors="replace"))


def quantize_rows(t, x):
    """float32 [nrows, k] -> uint8 [nrows, row_bytes] through ggb_quantize_rows (device kernels)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    if x.ndim == 1:
        x = x[None, :]
    nrows, k = x.shape
    out = np.zeros((nrows, N.TYPE_SIZE[t] * (k // N.BLCK_SIZE[t])), dtype=np.uint8)
    N.check(N.lib().ggb_quantize_rows(t, x.ctypes.data, out.ctypes.data, nrows, k))
    return out


def dequantize_rows(t, q, k):
    q = np.ascontiguousarray(q, dtype=np.uint8)
    if q.ndim == 1:
        q = q[None, :]
    out = np.zeros((q.shape[0], k), dtype=np.float32)
    N.check(N.lib().ggb_dequantize_rows(t, q.ctypes.data, out.ctypes.data, q.shape[0], k))
    return out
