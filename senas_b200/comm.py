"""NCCL communicator owned by libsenas_b200 (C ABI ``senas_comm_*``), one process per GPU.

Why not ``torch.distributed``'s process group for the hot exchange: its collectives cannot be captured into the CUDA
graph of the search step on this stack (the watchdog thread polls CUDA events during capture -- round 1 measured a
deadlock), which left two blocking all-reduces and two graph-launch gaps on the step's critical path.  A communicator
without a watchdog can: ``ncclAllReduce`` on the capture stream becomes a node of the graph.  ``torch.distributed``
(any backend) is still used once, as the side channel that hands rank 0's NCCL unique id to the other ranks.
"""
import ctypes as C

import torch

from . import _lib


class Comm:
    def __init__(self, rank=None, world=None, group=None, device=None, lib=None):
        import torch.distributed as dist
        self.lib = lib if lib is not None else _lib.get()
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        ident = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            buf = (C.c_char * 128)()
            _lib.check(self.lib, self.lib.senas_comm_unique_id(C.cast(buf, C.c_void_p)))
            ident = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        backend = dist.get_backend(group)
        t = ident.to(self.device) if backend == 'nccl' else ident
        dist.broadcast(t, 0, group=group)
        raw = bytes(t.cpu().numpy().tobytes())
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib, self.lib.senas_comm_init(C.c_char_p(raw), self.rank, self.world, C.byref(handle)))
        self.handle = handle

    def all_reduce_(self, flat, stream=None):
        """In-place SUM of a contiguous fp32 tensor over all ranks, enqueued on ``stream`` (default: the current one)."""
        if flat.dtype != torch.float32 or not flat.is_contiguous():
            raise ValueError('senas_b200.Comm.all_reduce_: contiguous fp32 tensor expected')
        st = (stream or torch.cuda.current_stream(flat.device)).cuda_stream
        _lib.check(self.lib, self.lib.senas_comm_allreduce(self.handle, flat.data_ptr(), flat.numel(), st))
        return flat

    def destroy(self):
        """Release the communicator.  Call it explicitly, on every rank, AFTER every CUDA graph that captured one of its
        all-reduces has been destroyed and the device is idle (NCCL requirement); nothing is done at interpreter exit
        (a communicator torn down in the middle of process shutdown can block forever)."""
        if getattr(self, 'handle', None):
            torch.cuda.synchronize(self.device)
            self.lib.senas_comm_destroy(self.handle)
            self.handle = None
