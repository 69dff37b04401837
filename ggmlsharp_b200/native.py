"""ctypes bindings of include/ggb200.h (libggb200.so) and of the host mirror (libggml_host.so)."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIBDIR = os.path.join(HERE, "lib")

GGML_MAX_DIMS, GGML_MAX_NODES, GGML_MAX_OPT = 4, 4096, 4
F32, F16, Q4_0, Q4_1, Q4_2, Q5_0, Q5_1, Q8_0, Q8_1, I8, I16, I32 = 0, 1, 2, 3, 4, 6, 7, 8, 9, 10, 11, 12
OP_NONE, OP_MUL_MAT, OP_CPY = 0, 20, 22
OP_DUP, OP_ADD, OP_MUL, OP_REPEAT, OP_SILU, OP_RMS_NORM, OP_SCALE, OP_CONT, OP_TRANSPOSE = 1, 2, 4, 10, 17, 19, 21, 23, 27
OK, E_INVALID, E_UNSUPPORTED, E_CUDA, E_NOMEM, E_ABI, E_NODEVICE = 0, -1, -2, -3, -4, -5, -6
GRAPH_KEEP_ON_DEVICE, GRAPH_NO_WEIGHT_CACHE, GRAPH_MUL_MAT_ONLY, GRAPH_SHARD = 1, 2, 4, 8
MM_W_IN_FLIGHT = 1
MM_X_HOST = 2
OP_SQR = 6

TYPE_SIZE = {F32: 4, F16: 2, Q4_0: 20, Q4_1: 24, Q4_2: 10, Q5_0: 22, Q5_1: 24, Q8_0: 36, Q8_1: 44, I8: 1, I16: 2, I32: 4}
BLCK_SIZE = {F32: 1, F16: 1, Q4_0: 32, Q4_1: 32, Q4_2: 16, Q5_0: 32, Q5_1: 32, Q8_0: 32, Q8_1: 32, I8: 1, I16: 1, I32: 1}


class GgbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("ggb status %d: %s" % (code, msg))
        self.code = code


class ggml_tensor(C.Structure):
    pass


ggml_tensor._fields_ = [
    ("type", C.c_int32), ("n_dims", C.c_int32),
    ("ne", C.c_int64 * GGML_MAX_DIMS), ("nb", C.c_uint64 * GGML_MAX_DIMS),
    ("op", C.c_int32), ("is_param", C.c_uint8), ("_pad0", C.c_uint8 * 3),
    ("grad", C.POINTER(ggml_tensor)), ("src0", C.POINTER(ggml_tensor)), ("src1", C.POINTER(ggml_tensor)),
    ("opt", C.c_int64 * GGML_MAX_OPT),
    ("n_tasks", C.c_int32), ("perf_runs", C.c_int32), ("perf_cycles", C.c_int64), ("perf_time_us", C.c_int64),
    ("data", C.c_void_p), ("padding", C.c_uint8 * 8)]


class ggml_cgraph(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("n_leafs", C.c_int32), ("n_threads", C.c_int32), ("_pad0", C.c_int32),
                ("work_size", C.c_uint64), ("work", C.POINTER(ggml_tensor)),
                ("nodes", C.POINTER(ggml_tensor) * GGML_MAX_NODES),
                ("grads", C.POINTER(ggml_tensor) * GGML_MAX_NODES),
                ("leafs", C.POINTER(ggml_tensor) * GGML_MAX_NODES),
                ("perf_runs", C.c_int32), ("_pad1", C.c_int32), ("perf_cycles", C.c_int64), ("perf_time_us", C.c_int64)]


class ggml_init_params(C.Structure):
    _fields_ = [("mem_size", C.c_uint64), ("mem_buffer", C.c_void_p), ("no_alloc", C.c_uint8)]


class ggml_object(C.Structure):
    pass


ggml_object._fields_ = [("offs", C.c_uint64), ("size", C.c_uint64), ("next", C.POINTER(ggml_object)), ("padding", C.c_uint8 * 8)]


class ggml_context(C.Structure):
    _fields_ = [("mem_size", C.c_uint64), ("mem_buffer", C.c_void_p), ("mem_buffer_owned", C.c_uint8), ("no_alloc", C.c_uint8),
                ("n_objects", C.c_int32), ("objects_begin", C.POINTER(ggml_object)), ("objects_end", C.POINTER(ggml_object)),
                ("scratch", C.c_uint64 * 3), ("scratch_save", C.c_uint64 * 3)]


class ggb_dev_mm(C.Structure):
    _fields_ = [("type", C.c_int32), ("n_peers", C.c_int32), ("M", C.c_int64), ("K", C.c_int64), ("N", C.c_int64),
                ("W", C.c_void_p), ("nb01", C.c_int64), ("X", C.c_void_p), ("ldx_bytes", C.c_int64),
                ("Y", C.c_void_p), ("ldy_bytes", C.c_int64), ("Y_peer", C.c_void_p * 7),
                ("W_rowexp", C.c_void_p), ("flags", C.c_int32), ("_pad", C.c_int32)]


class ggb_stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("weight_uploads", C.c_uint64), ("weight_cache_hits", C.c_uint64), ("nodes_executed", C.c_uint64),
                ("last_graph_device_ms", C.c_double), ("timed_kernel_ms", C.c_double), ("timed_kernel_launches", C.c_uint64),
                ("graph_replays", C.c_uint64)]


assert C.sizeof(ggml_tensor) == 176 and ggml_tensor.data.offset == 160
assert C.sizeof(ggml_cgraph) == 98360 and ggml_cgraph.nodes.offset == 32
assert C.sizeof(ggml_context) == 88 and C.sizeof(ggml_object) == 32

TP = C.POINTER(ggml_tensor)

# every symbol include/ggb200.h declares: name -> (restype, argtypes)
GGB_SYMBOLS = {
    "ggb_last_error": (C.c_char_p, []),
    "ggb_abi_version": (C.c_int, []),
    "ggb_abi_check": (C.c_int, [C.c_int] * 6),
    "ggb_init": (C.c_int, []),
    "ggb_shutdown": (C.c_int, []),
    "ggb_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "ggb_pool_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "ggb_pool_adopt": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "ggb_pool_free": (C.c_int, [C.c_void_p]),
    "ggb_tensor_invalidate": (C.c_int, [C.c_void_p, TP]),
    "ggb_pool_set_weight_cache": (C.c_int, [C.c_void_p, C.c_int]),
    "ggb_pool_set_row_split": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t]),
    "ggb_row_split_rows": (C.c_int, [C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "ggb_graph_plan": (C.c_int, [C.POINTER(ggml_cgraph), C.c_int, C.POINTER(C.c_uint8)]),
    "ggb_dev_weight_rowexp": (C.c_int, [C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "ggb_mul_mat_node": (C.c_int, [C.c_void_p, TP]),
    "ggb_graph_compute_mul_mats": (C.c_int, [C.c_void_p, C.POINTER(ggml_cgraph), C.c_int, C.POINTER(C.c_uint8)]),
    "ggb_quantize_rows": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]),
    "ggb_dequantize_rows": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]),
    "ggb_dev_workspace_bytes": (C.c_size_t, [C.POINTER(ggb_dev_mm), C.c_int]),
    "ggb_dev_mul_mat_batch": (C.c_int, [C.POINTER(ggb_dev_mm), C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "ggb_dev_mul_mat_batch_phase": (C.c_int, [C.POINTER(ggb_dev_mm), C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int]),
    "ggb_dev_quantize_rows": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    "ggb_dev_dequantize_rows": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    "ggb_dev_binary": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "ggb_dev_scale": (C.c_int, [C.c_void_p, C.c_float, C.c_int64, C.c_void_p]),
    "ggb_dev_silu": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "ggb_dev_rms_norm": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p]),
    "ggb_dev_repeat": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p]),
    "ggb_dev_cont": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_uint64), C.c_void_p, C.c_void_p]),
    "ggb_dev_add_q": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    "ggb_dev_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "ggb_dev_free": (C.c_int, [C.c_void_p]),
    "ggb_dev_upload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "ggb_dev_download": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "ggb_stream_sync": (C.c_int, [C.c_void_p]),
    "ggb_ipc_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ggb_ipc_open": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "ggb_ipc_close": (C.c_int, [C.c_void_p]),
    "ggb_peer_barrier": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_uint64, C.c_void_p]),
    "ggb_peer_push_barrier": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_uint64, C.c_void_p]),
    "ggb_get_stats": (C.c_int, [C.POINTER(ggb_stats)]),
    "ggb_reset_stats": (C.c_int, []),
    "ggb_set_kernel_timing": (C.c_int, [C.c_int]),
    "ggb_set_decode_program": (C.c_int, [C.c_int]),
}

HOST_SYMBOLS = {
    "ggml_host_last_status": (C.c_int, []),
    "ggml_host_pool_of": (C.c_void_p, [C.POINTER(ggml_context)]),
    "ggml_host_set_weight_cache": (C.c_int, [C.POINTER(ggml_context), C.c_int]),
    "ggml_host_set_row_split": (C.c_int, [C.POINTER(ggml_context), C.c_int, C.c_size_t]),
    "ggml_sqr": (TP, [C.POINTER(ggml_context), TP]),
    "ggml_sqr_inplace": (TP, [C.POINTER(ggml_context), TP]),
    "ggml_init": (C.POINTER(ggml_context), [ggml_init_params]),
    "ggml_free": (None, [C.POINTER(ggml_context)]),
    "ggml_nelements": (C.c_int64, [TP]),
    "ggml_nrows": (C.c_int64, [TP]),
    "ggml_nbytes": (C.c_uint64, [TP]),
    "ggml_used_mem": (C.c_uint64, [C.POINTER(ggml_context)]),
    "ggml_new_tensor": (TP, [C.POINTER(ggml_context), C.c_int, C.c_int, C.POINTER(C.c_int64)]),
    "ggml_new_tensor_1d": (TP, [C.POINTER(ggml_context), C.c_int, C.c_int64]),
    "ggml_new_tensor_2d": (TP, [C.POINTER(ggml_context), C.c_int, C.c_int64, C.c_int64]),
    "ggml_new_tensor_3d": (TP, [C.POINTER(ggml_context), C.c_int, C.c_int64, C.c_int64, C.c_int64]),
    "ggml_new_tensor_4d": (TP, [C.POINTER(ggml_context), C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int64]),
    "ggml_set_f32": (TP, [TP, C.c_float]),
    "ggml_get_f32_1d": (C.c_float, [TP, C.c_int]),
    "ggml_mul_mat": (TP, [C.POINTER(ggml_context), TP, TP]),
    "ggml_cpy": (TP, [C.POINTER(ggml_context), TP, TP]),
    "ggml_add": (TP, [C.POINTER(ggml_context), TP, TP]),
    "ggml_add_inplace": (TP, [C.POINTER(ggml_context), TP, TP]),
    "ggml_mul": (TP, [C.POINTER(ggml_context), TP, TP]),
    "ggml_scale": (TP, [C.POINTER(ggml_context), TP, TP]),
    "ggml_repeat": (TP, [C.POINTER(ggml_context), TP, TP]),
    "ggml_silu": (TP, [C.POINTER(ggml_context), TP]),
    "ggml_silu_inplace": (TP, [C.POINTER(ggml_context), TP]),
    "ggml_rms_norm": (TP, [C.POINTER(ggml_context), TP]),
    "ggml_cont": (TP, [C.POINTER(ggml_context), TP]),
    "ggml_transpose": (TP, [C.POINTER(ggml_context), TP]),
    "ggml_dup_tensor": (TP, [C.POINTER(ggml_context), TP]),
    "ggml_view_tensor": (TP, [C.POINTER(ggml_context), TP]),
    "ggml_build_forward_expand": (None, [C.POINTER(ggml_cgraph), TP]),
    "ggml_build_forward_into": (None, [C.POINTER(ggml_cgraph), TP]),
    "ggml_graph_compute": (None, [C.POINTER(ggml_context), C.POINTER(ggml_cgraph)]),
}


def _load(name, table):
    path = os.path.join(LIBDIR, name)
    if not os.path.exists(path):
        raise OSError("%s is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(there is no CPU fallback for the CUDA path)" % path)
    dll = C.CDLL(path, mode=C.RTLD_GLOBAL)
    for sym, (res, args) in table.items():
        fn = getattr(dll, sym)          # AttributeError if the library does not export a declared symbol
        fn.restype, fn.argtypes = res, args
    return dll


_lib = None
_host = None


def lib():
    global _lib
    if _lib is None:
        _lib = _load("libggb200.so", GGB_SYMBOLS)
    return _lib


def host():
    global _host
    if _host is None:
        lib()
        _host = _load("libggml_host.so", HOST_SYMBOLS)
    return _host


def check(rc):
    if rc < 0:
        raise GgbError(rc, lib().ggb_last_error().decode(errors="replace"))
    return rc


def stats():
    s = ggb_stats()
    check(lib().ggb_get_stats(C.byref(s)))
    return s
