"""Seam B's node selection (ggb_graph_plan: the same selection ggb_graph_compute_mul_mats runs, without a device).

The reference executes nodes strictly in order (Ggml.cs:3539-3704).  Seam B runs the nodes it takes BEFORE the caller's CPU loop
runs the rest, so a node must be refused whenever that reordering could be observed through memory the pointer graph does not
show: an earlier CPU node that writes bytes a candidate reads (ggml_sqr_inplace on a leaf, Ggml.cs:6917; a same-type CPY into a
leaf), a candidate that writes in place over bytes an earlier CPU node still reads, and operands that overlap an earlier device
result without lying inside it (a CPY into an offset view of a cache, then a MUL_MAT over the whole cache)."""
import ctypes as C

import numpy as np

from ggmlsharp_b200 import ggml, native as N


def _ctx(nbytes=8 << 20):
    buf = np.zeros(nbytes, dtype=np.uint8)
    assert buf.ctypes.data % 16 == 0
    return ggml.Context(nbytes, mem_buffer=buf)


def _plan(g, flags=0):
    done = (C.c_uint8 * N.GGML_MAX_NODES)()
    n = N.lib().ggb_graph_plan(C.byref(g), flags, done)
    assert n >= 0, N.lib().ggb_last_error()
    return n, [int(done[i]) for i in range(g.n_nodes)]


def _index(g, t):
    addr = C.addressof(t.contents)
    for i in range(g.n_nodes):
        if C.addressof(g.nodes[i].contents) == addr:
            return i
    raise AssertionError("node not in graph")


def _expand(c, *outs):
    g = N.ggml_cgraph()
    g.n_threads = 4
    for o in outs:
        N.host().ggml_build_forward_expand(C.byref(g), o)
    return g


def test_plain_mul_mat_chain_is_taken():
    with _ctx() as c:
        w = c.new_tensor(N.F32, 64, 32)
        x = c.new_tensor(N.F32, 64, 4)
        y = c.mul_mat(w, x)
        z = c.op("silu", y)
        g = c.build_forward(z)
        n, done = _plan(g)
        assert n == 2 and done == [1, 1]


def test_cpu_inplace_writer_before_a_reader_blocks_the_reader():
    # x2 = sqr_inplace(x) is a CPU node that rewrites the leaf x; the MUL_MAT names x itself (not x2), so the pointer graph
    # shows no dependency -- but the reference runs sqr first and the MUL_MAT sees x*x.
    with _ctx() as c:
        w = c.new_tensor(N.F32, 64, 32)
        x = c.new_tensor(N.F32, 64, 4)
        x2 = c.op("sqr_inplace", x)
        y = c.mul_mat(w, x)
        g = _expand(c, x2, y)
        assert _index(g, x2) < _index(g, y)
        n, done = _plan(g)
        assert done[_index(g, x2)] == 0
        assert done[_index(g, y)] == 0, "MUL_MAT would read x before the CPU node squares it"
        assert n == 0


def test_dependents_of_a_blocked_node_are_blocked_too():
    with _ctx() as c:
        w = c.new_tensor(N.F32, 64, 32)
        x = c.new_tensor(N.F32, 64, 4)
        x2 = c.op("sqr_inplace", x)
        y = c.mul_mat(w, x)
        z = c.op("silu", y)
        g = _expand(c, x2, z)
        n, done = _plan(g)
        assert n == 0 and done[_index(g, z)] == 0


def test_cpu_writer_after_the_reader_does_not_block():
    # node order: MUL_MAT first, then the CPU node rewrites x -- running the MUL_MAT ahead of the CPU loop keeps the order
    with _ctx() as c:
        w = c.new_tensor(N.F32, 64, 32)
        x = c.new_tensor(N.F32, 64, 4)
        y = c.mul_mat(w, x)
        x2 = c.op("sqr_inplace", x)
        g = _expand(c, y, x2)
        assert _index(g, y) < _index(g, x2)
        n, done = _plan(g)
        assert done[_index(g, y)] == 1 and done[_index(g, x2)] == 0 and n == 1


def test_gpu_inplace_node_is_not_hoisted_over_an_earlier_cpu_reader():
    # s = sqr(x) (CPU, reads the leaf x), then scale(x) rewrites x in place: hoisted ahead of the CPU loop, sqr would see scaled data
    with _ctx() as c:
        x = c.new_tensor(N.F32, 64, 4)
        f = c.new_tensor(N.F32, 1)
        s = c.op("sqr", x)
        xs = c.op("scale", x, f)
        g = _expand(c, s, xs)
        assert _index(g, s) < _index(g, xs)
        n, done = _plan(g)
        assert done[_index(g, xs)] == 0 and n == 0


def test_unrelated_cpu_node_does_not_block():
    with _ctx() as c:
        w = c.new_tensor(N.F32, 64, 32)
        x = c.new_tensor(N.F32, 64, 4)
        u = c.new_tensor(N.F32, 16)
        u2 = c.op("sqr_inplace", u)
        y = c.mul_mat(w, x)
        g = _expand(c, u2, y)
        n, done = _plan(g)
        assert done[_index(g, y)] == 1 and done[_index(g, u2)] == 0


def test_cpu_same_type_cpy_into_a_weight_blocks_the_mul_mat():
    # F32 -> F32 CPY is not on the device path (validate_cpy: F16 / quantized destinations only); it rewrites the leaf w
    with _ctx() as c:
        w = c.new_tensor(N.F32, 64, 32)
        wsrc = c.new_tensor(N.F32, 64, 32)
        x = c.new_tensor(N.F32, 64, 4)
        cp = c.cpy(wsrc, w)
        y = c.mul_mat(w, x)
        g = _expand(c, cp, y)
        assert _index(g, cp) < _index(g, y)
        n, done = _plan(g)
        assert done[_index(g, cp)] == 0 and done[_index(g, y)] == 0


def _offset_view(c, base, ne0, ne1, byte_offset):
    """A fresh tensor header over part of base's bytes -- what C# user code builds by writing ->data / ->ne / ->nb (the reference
    ports no ggml_view_1d/2d builder, but the fields are public)."""
    v = N.host().ggml_view_tensor(c.ctx, base)
    v.contents.ne[0], v.contents.ne[1] = ne0, ne1
    v.contents.nb[2] = v.contents.nb[1] * ne1
    v.contents.nb[3] = v.contents.nb[2]
    v.contents.data = base.contents.data + byte_offset
    return v


def test_cpy_into_an_offset_view_then_mul_mat_over_the_whole_cache():
    # the CPY's result covers rows [8, 16) of the cache; a MUL_MAT over rows [0, 16) starts BEFORE that range and
    # overlaps it.  Start-pointer containment missed this and multiplied stale host rows.
    with _ctx() as c:
        K = 64
        cache = c.new_tensor(N.F16, K, 16)
        src = c.new_tensor(N.F32, K, 8)
        x = c.new_tensor(N.F32, K, 4)
        dst_view = _offset_view(c, cache, K, 8, 8 * K * 2)
        cp = c.cpy(src, dst_view)
        y = c.mul_mat(cache, x)
        g = _expand(c, cp, y)
        assert _index(g, cp) < _index(g, y)
        n, done = _plan(g)
        assert done[_index(g, cp)] == 1, "the F32 -> F16 CPY itself is on the path"
        assert done[_index(g, y)] == 0, "the MUL_MAT's src0 is partly the CPY's device result, partly host bytes"
        # a MUL_MAT over exactly the rows the CPY wrote reads the device result: fine
        with _ctx() as c2:
            cache2 = c2.new_tensor(N.F16, K, 16)
            src2 = c2.new_tensor(N.F32, K, 8)
            x2 = c2.new_tensor(N.F32, K, 4)
            v2 = _offset_view(c2, cache2, K, 8, 8 * K * 2)
            cp2 = c2.cpy(src2, v2)
            rd = _offset_view(c2, cache2, K, 8, 8 * K * 2)
            y2 = c2.mul_mat(rd, x2)
            g2 = _expand(c2, cp2, y2)
            n2, done2 = _plan(g2)
            assert n2 == 2 and done2[_index(g2, y2)] == 1
        # ... and one over rows the CPY did not touch reads the host bytes: fine too
        with _ctx() as c3:
            cache3 = c3.new_tensor(N.F16, K, 16)
            src3 = c3.new_tensor(N.F32, K, 8)
            x3 = c3.new_tensor(N.F32, K, 4)
            v3 = _offset_view(c3, cache3, K, 8, 8 * K * 2)
            cp3 = c3.cpy(src3, v3)
            rd3 = _offset_view(c3, cache3, K, 8, 0)
            y3 = c3.mul_mat(rd3, x3)
            g3 = _expand(c3, cp3, y3)
            n3, done3 = _plan(g3)
            assert n3 == 2 and done3[_index(g3, y3)] == 1


def test_mul_mat_only_flag_leaves_neighbours():
    with _ctx() as c:
        w = c.new_tensor(N.F32, 64, 32)
        x = c.new_tensor(N.F32, 64, 4)
        y = c.mul_mat(w, x)
        z = c.op("silu", y)
        g = c.build_forward(z)
        n, done = _plan(g, N.GRAPH_MUL_MAT_ONLY)
        assert n == 1 and done == [1, 0]
