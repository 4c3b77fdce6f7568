"""Peer-visible device memory (include/gse.h: gse_peer_alloc / gse_peer_open): plain cudaMalloc
allocations with a CUDA IPC handle, wrapped as torch tensors through ``__cuda_array_interface__``.
torch only views the memory; allocation, export and mapping go through the C ABI."""
import ctypes

import torch

from gpu_se_b200 import _lib


class PeerBuffer:
    """Memory owned by this process that other ranks of the node may map."""

    def __init__(self, device, nbytes):
        self.device = torch.device(device)
        self.nbytes = int(nbytes)
        ptr = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(_lib.GSE_IPC_HANDLE_BYTES)
        _lib.check(_lib.lib.gse_peer_alloc(self.device.index, self.nbytes, ctypes.byref(ptr), handle))
        self.ptr = int(ptr.value)
        self.handle = handle.raw
        self.__cuda_array_interface__ = {"shape": (self.nbytes,), "typestr": "|u1", "data": (self.ptr, False),
                                         "version": 2, "strides": None}

    def tensor(self, dtype, shape):
        with torch.cuda.device(self.device):
            t = torch.as_tensor(self, device=self.device)
        return t.view(dtype).reshape(shape)

    def close(self):
        if getattr(self, "ptr", 0):
            _lib.lib.gse_peer_free(self.device.index, ctypes.c_void_p(self.ptr))
            self.ptr = 0


class PeerMapping:
    """Another rank's PeerBuffer mapped into this process."""

    def __init__(self, device, handle):
        self.device = torch.device(device)
        ptr = ctypes.c_void_p()
        _lib.check(_lib.lib.gse_peer_open(self.device.index, handle, ctypes.byref(ptr)))
        self.ptr = int(ptr.value)

    def close(self):
        if getattr(self, "ptr", 0):
            _lib.lib.gse_peer_close(self.device.index, ctypes.c_void_p(self.ptr))
            self.ptr = 0


def peer_zeros(device, shape, dtype):
    """(tensor, PeerBuffer) of zeros in peer-visible memory."""
    n = 1
    for d in shape:
        n *= int(d)
    buf = PeerBuffer(device, max(n * torch.empty((), dtype=dtype).element_size(), 16))
    return buf.tensor(dtype, shape), buf
