"""Multi-GPU plumbing for the registration path (one process per GPU, torch.distributed).

The path shards across independent sequences only (frames of one sequence are serially dependent,
ref: src/front_end/front_end.cpp:329-334), so there is NO data-path collective.  The one exchange that exists is
the broadcast of a SHARED local map (BASELINE config 4): the built map is a single contiguous device blob
(tloam_b200_map_export / _import), moved with one broadcast (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
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


def broadcast_shared_map(reg, src=0, device=None):
    """Rank `src` has built a map (set_input_target); every other rank adopts it without running the build.
    Two collectives: the blob size (8 B), then the blob.  Returns the number of bytes moved."""
    rank = dist.get_rank()
    size = torch.zeros(1, dtype=torch.int64, device=device)
    if rank == src:
        size[0] = reg.map_blob_size()
    dist.broadcast(size, src=src)
    n = int(size[0])
    buf = torch.empty(n, dtype=torch.uint8, device=device)
    if rank == src:
        reg.map_export(buf.data_ptr(), n)
    dist.broadcast(buf, src=src)
    if rank != src:
        if buf.is_cuda:
            torch.cuda.current_stream().synchronize()
        reg.map_import(buf.data_ptr(), n)
    return n
