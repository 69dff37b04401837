"""Row-split of a mul_mat across the GPUs of one box.

The reference splits src0 rows over OS threads: dr = ceil(nr / nth), thread ith owns rows [dr*ith, min(dr*ith + dr, nr))
(Ggml.cs:6665-6672).  Here nth = world size and the "thread" is a GPU: rank g holds the contiguous byte slice of the
weight tensor for its rows, computes its column block of dst, and the blocks are exchanged so every rank ends with the
full dst.  Two exchanges exist:

  * ``nccl``   -- torch.distributed all_gather of the per-rank blocks (baseline);
  * ``fused``  -- the GEMV/GEMM epilogue stores each result straight into every peer's copy of dst through CUDA-IPC mapped
                  pointers (``ggb_dev_mm.Y_peer``), followed by one flag barrier over NVLink (``ggb_peer_barrier``).

Only host-side plumbing lives here; there is no arithmetic in this module.
"""
import ctypes as C


def shard_rows(nr, world, rank):
    """The reference's thread split with nth = world: (first_row, n_rows) of `rank` (Ggml.cs:6665-6672)."""
    dr = (nr + world - 1) // world
    ir0 = dr * rank
    ir1 = min(ir0 + dr, nr)
    return ir0, max(ir1 - ir0, 0)


def shard_bytes(nr, row_bytes, world, rank):
    """Byte range [off, off + n) of the weight tensor that `rank` owns (rows are contiguous nb01-byte ranges)."""
    r0, n = shard_rows(nr, world, rank)
    return r0 * row_bytes, n * row_bytes


class SymmetricBuffer:
    """One device allocation per rank with the same layout everywhere, mapped into every peer with CUDA IPC.

    layout: [world x uint64 barrier flags | uint32 CTA counter at byte 128, padded to 256 B][payload_bytes]
    """
    FLAG_BYTES = 256

    def __init__(self, payload_bytes, rank, world, all_gather_object):
        from . import native as N
        self.N, self.rank, self.world = N, rank, world
        L = N.lib()
        self.total = self.FLAG_BYTES + payload_bytes
        p = C.c_void_p()
        N.check(L.ggb_dev_alloc(self.total, C.byref(p)))
        self.base = p.value
        zeros = (C.c_uint8 * self.FLAG_BYTES)()
        N.check(L.ggb_dev_upload(self.base, zeros, self.FLAG_BYTES))
        handle = (C.c_uint8 * 64)()
        N.check(L.ggb_ipc_export(self.base, handle))
        handles = all_gather_object(bytes(handle))            # list of `world` 64-byte handles, index = rank
        self.peer_base = []
        for r, h in enumerate(handles):
            if r == rank:
                self.peer_base.append(self.base)
                continue
            q = C.c_void_p()
            hb = (C.c_uint8 * 64).from_buffer_copy(h)
            N.check(L.ggb_ipc_open(hb, C.byref(q)))
            self.peer_base.append(q.value)
        self._flags = (C.c_void_p * world)(*[C.c_void_p(b) for b in self.peer_base])
        self.epoch = 0
        self._payloads = None

    def payload(self, r=None):
        return (self.base if r is None else self.peer_base[r]) + self.FLAG_BYTES

    def barrier(self, stream):
        self.epoch += 1
        self.N.check(self.N.lib().ggb_peer_barrier(self._flags, self.rank, self.world, self.epoch, stream))

    def push_barrier(self, stream, seg_offset, seg_bytes, seg_stride, n_seg):
        """Copy this rank's segments of the payload into every peer's payload and close the step with the flag barrier.
        The epoch lives in device memory (epoch argument 0), so the call can sit inside a replayed CUDA graph; do not mix
        with barrier() on the same buffer."""
        if self._payloads is None:
            self._payloads = (C.c_void_p * self.world)(*[C.c_void_p(b + self.FLAG_BYTES) for b in self.peer_base])
        self.N.check(self.N.lib().ggb_peer_push_barrier(self._payloads, self._flags, self.base + 128, self.rank, self.world,
                                                        seg_offset, seg_bytes, seg_stride, n_seg, 0, stream))

    def close(self):
        L = self.N.lib()
        for r, b in enumerate(self.peer_base):
            if r != self.rank:
                L.ggb_ipc_close(b)
        L.ggb_dev_free(self.base)
