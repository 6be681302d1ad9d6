"""Multi-GPU plumbing for the registration path (one process per GPU, torch.distributed).

The path shards across independent sequences only (frames of one sequence are serially dependent,
ref: src/front_end/front_end.cpp:329-334), so there is NO data-path collective.  The one exchange that exists is
the broadcast of a SHARED local map (BASELINE config 4): the built map is a single contiguous device blob whose
layout is a pure function of the configuration and the four point counts, so it moves with ONE collective per map
epoch -- no size handshake, no export / import copies, no host synchronisation -- straight from the builder's blob
into a second blob of every receiving handle, on a side stream, while frames keep registering against the active
map (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
import contextlib
import ctypes

import numpy as np
import torch
import torch.distributed as dist

SEQUENCES = ("00", "02", "05", "08", "01", "06", "07", "09")


def sequence_for_rank(rank):
    """Independent KITTI-shaped stream driven by this rank."""
    return SEQUENCES[rank % len(SEQUENCES)]


def aggregate_frames_per_sec(frames_this_rank, ms_this_rank, device=None):
    """Whole-job throughput: all ranks' frames / the slowest rank's device time (never wall clock)."""
    t = torch.tensor([float(frames_this_rank), float(ms_this_rank)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        frames = t[0:1].clone()
        ms = t[1:2].clone()
        dist.all_reduce(frames, op=dist.ReduceOp.SUM)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(frames[0]) / (float(ms[0]) * 1e-3), float(ms[0])
    return float(t[0]) / (float(t[1]) * 1e-3), float(t[1])


class _DevView:
    """A raw device allocation of the library seen as a uint8 torch tensor (zero copy)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def _view(ptr, nbytes, device):
    if device is not None and torch.device(device).type == "cuda":
        return torch.as_tensor(_DevView(ptr, nbytes), device=device)
    return torch.from_numpy(np.ctypeslib.as_array((ctypes.c_uint8 * int(nbytes)).from_address(int(ptr))))


class SharedMapChannel:
    """One-collective transport of a shared local map.

    reg      this rank's registration handle (receives; needs map_recv_buffer / map_adopt)
    builder  on the source rank: the handle that BUILT the map (map_send_buffer / signal_stream); default `reg` itself,
             in which case the source registers against its own blob and receives nothing
    The point counts n_map are known to every rank (they define the blob layout), so no size is exchanged.
    """

    def __init__(self, reg, src=0, device=None, builder=None):
        self.reg, self.src, self.device = reg, src, device
        self.builder = builder if builder is not None else reg
        self.cuda = device is not None and torch.device(device).type == "cuda"
        self.side = torch.cuda.Stream(device) if self.cuda else None
        self.bytes = 0

    def _side_handle(self):
        return self.side.cuda_stream if self.cuda else 0

    def broadcast(self, n_map):
        """Enqueue the collective for the NEXT map (counts n_map) on the side stream.  Returns the bytes moved."""
        rank = dist.get_rank()
        own_blob = rank == self.src and self.builder is self.reg
        recv = None
        if not own_blob:
            ptr, nbytes = self.reg.map_recv_buffer(n_map)
            recv = _view(ptr, nbytes, self.device)
        ctx = torch.cuda.stream(self.side) if self.cuda else contextlib.nullcontext()
        with ctx:
            if rank == self.src:
                sptr, sbytes = self.builder.map_send_buffer()
                send = _view(sptr, sbytes, self.device)
                self.builder.signal_stream(self._side_handle())        # the collective starts after the build
                dist.broadcast(send, src=self.src)
                if recv is not None:
                    assert recv.numel() == send.numel()
                    recv.copy_(send, non_blocking=True)                # the source's own registration handle: local copy
                self.bytes = sbytes
            else:
                dist.broadcast(recv, src=self.src)
                self.bytes = recv.numel()
        return self.bytes

    def adopt(self):
        """Switch this rank's registration handle to the received map (device-side wait + pointer swap)."""
        if dist.get_rank() == self.src and self.builder is self.reg:
            return
        self.reg.map_adopt(self._side_handle())
