"""One process per GPU: the gather step of the multi-GPU path.  Rendering needs no
collective; finished shard buffers are gathered to rank 0 (NCCL over NVLink on the GPU box, gloo in the
CPU tests) where the frame is assembled."""
import torch
import torch.distributed as dist


def shard_of_rank():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def gather_tiles(local_buf, max_elems, gather_list=None, dst=0):
    """Gathers every rank's tile buffer (padded to max_elems so that all contributions have one size) to
    `dst`.  local_buf: 1-D tensor of >= max_elems elements on the rank's device.  Returns the list of
    per-rank tensors on dst, None elsewhere."""
    rank, world = shard_of_rank()
    if world == 1:
        return [local_buf]
    send = local_buf[:max_elems]
    if rank == dst:
        if gather_list is None:
            gather_list = [torch.empty_like(send) for _ in range(world)]
        dist.gather(send, gather_list, dst=dst)
        return gather_list
    dist.gather(send, None, dst=dst)
    return None


# ---- peer-memory frame: the render kernel's fold writes finished pixels straight into rank 0's memory -----------------
# Finished tiles are the only global stores of the render kernel.  When every rank can map rank 0's tile arena (CUDA IPC
# + NVLink peer access), each rank's kernel writes its shard into its slice of that arena directly: the "gather" is the
# kernel's own stores travelling over NVLink while it renders, and the only collective left is a barrier before rank 0
# assembles the frame.  Falls back to the NCCL gather above when IPC is unavailable.
import ctypes as _C

_cudart = None


class _IpcHandle(_C.Structure):  # cudaIpcMemHandle_t: 64 opaque bytes, passed BY VALUE to cudaIpcOpenMemHandle
    _fields_ = [("reserved", _C.c_char * 64)]


def _rt():
    global _cudart
    if _cudart is None:
        for name in ("libcudart.so.12", "libcudart.so"):
            try:
                _cudart = _C.CDLL(name)
                break
            except OSError:
                continue
        if _cudart is None:
            raise RuntimeError("libcudart not found")
        _cudart.cudaMalloc.argtypes = [_C.POINTER(_C.c_void_p), _C.c_size_t]
        _cudart.cudaFree.argtypes = [_C.c_void_p]
        _cudart.cudaIpcGetMemHandle.argtypes = [_C.POINTER(_IpcHandle), _C.c_void_p]
        _cudart.cudaIpcOpenMemHandle.argtypes = [_C.POINTER(_C.c_void_p), _IpcHandle, _C.c_uint]
        _cudart.cudaIpcCloseMemHandle.argtypes = [_C.c_void_p]
    return _cudart


class PeerArena:
    """world_size slices of `slice_bytes` on rank 0's GPU, mapped into every rank.  .ptr(k) = device address of slice k
    valid in THIS process."""

    def __init__(self, slice_bytes):
        rank, world = shard_of_rank()
        self.rank, self.world, self.slice_bytes = rank, world, int(slice_bytes)
        rt = _rt()
        self._owner = rank == 0
        self._base = _C.c_void_p()
        handle = _IpcHandle()
        ok = True
        if rank == 0:
            ok = rt.cudaMalloc(_C.byref(self._base), self.slice_bytes * world) == 0
            ok = ok and rt.cudaIpcGetMemHandle(_C.byref(handle), self._base) == 0
        box = [_C.string_at(_C.addressof(handle), 64) if ok else None]  # raw 64 bytes (the field accessor would stop at a NUL)
        dist.broadcast_object_list(box, src=0)
        if box[0] is None:
            raise RuntimeError("rank 0 could not export the arena")
        if rank != 0:
            h = _IpcHandle.from_buffer_copy(box[0])
            rc = rt.cudaIpcOpenMemHandle(_C.byref(self._base), h, 1)  # cudaIpcMemLazyEnablePeerAccess
            ok = rc == 0 and bool(self._base.value)
        flags = [None] * world
        dist.all_gather_object(flags, bool(ok))
        if not all(flags):
            self.close()
            raise RuntimeError("CUDA IPC mapping failed on some rank")

    def ptr(self, k):
        return self._base.value + k * self.slice_bytes

    def close(self):
        if getattr(self, "_base", None) is not None and self._base.value:
            rt = _rt()
            if self._owner:
                rt.cudaFree(self._base)
            else:
                rt.cudaIpcCloseMemHandle(self._base)
            self._base = _C.c_void_p()
