"""Peer-memory arena for the SyncBN statistics exchange (csrc/peer.cu): one cudaMalloc'ed buffer per rank, exported
with CUDA IPC, mapped by every peer.  torch.distributed is used once, to exchange the 64-byte IPC handles."""
import ctypes

import torch

from . import _C


class _RawCuda:
    """Zero-copy torch view of a raw device pointer through __cuda_array_interface__."""

    def __init__(self, ptr, count, typestr):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (ptr, False), "version": 2}


class PeerArena:
    @staticmethod
    def _view(ptr, count, typestr, device):
        """torch tensor over the arena (the CPU simulator's test harness substitutes a host view)"""
        raw = _RawCuda(ptr, count, typestr)
        t = torch.as_tensor(raw, device=device)
        t._keep_raw = raw
        return t

    def __init__(self, nfloats, dist, device):
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        assert self.world <= 8
        self.nbytes = (nfloats * 4 + 255) // 256 * 256
        L = _C.lib()
        local = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * 64)()
        _C.check(L.sseg_peer_alloc(self.nbytes, ctypes.byref(local), handle))
        self.local_ptr = local.value
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle))
        self.bases = (ctypes.c_void_p * self.world)()
        err = None
        for r, h in enumerate(handles):
            if r == self.rank:
                self.bases[r] = self.local_ptr
            else:
                p = ctypes.c_void_p()
                buf = (ctypes.c_ubyte * 64).from_buffer_copy(h)
                try:
                    _C.check(L.sseg_peer_open(buf, ctypes.byref(p)))
                    self.bases[r] = p.value
                except _C.SsegError as exc:  # keep going: every rank must reach the barrier below
                    err = exc
        self.floats = self._view(self.local_ptr, self.nbytes // 4, "<f4", device)
        self.ints = self._view(self.local_ptr, self.nbytes // 4, "<i4", device)
        dist.barrier()  # every rank has mapped every arena before anyone starts signalling
        if err is not None:
            raise err
