"""world_size-2 tests of the multi-GPU plumbing.  On CPU (gloo) the registration handle is replaced by a
host-memory stand-in with the same zero-copy blob contract (map_send_buffer / map_recv_buffer / map_adopt), so the
protocol (ONE collective per map epoch, no size handshake, max-over-ranks timing, per-rank sequence assignment) is exercised without a
GPU; with >= 2 GPUs the same code runs over NCCL against real handles (marked gpu)."""
import ctypes
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class HostMapStandIn:
    """The zero-copy blob contract of LocalRegistration (map_send_buffer / map_recv_buffer / map_adopt / signal_stream)
    on host memory: the layout size is a pure function of the point counts (here: 16 bytes per point + 256)."""

    def __init__(self, payload=None):
        self.blob = (ctypes.c_uint8 * len(payload)).from_buffer_copy(payload) if payload is not None else None
        self.incoming = None
        self.active = None
        self.adopted = 0

    @staticmethod
    def layout_bytes(n_map):
        return 256 + 16 * int(sum(n_map))

    def map_send_buffer(self):
        return ctypes.addressof(self.blob), len(self.blob)

    def map_recv_buffer(self, n_map):
        self.incoming = (ctypes.c_uint8 * self.layout_bytes(n_map))()
        return ctypes.addressof(self.incoming), len(self.incoming)

    def map_adopt(self, stream):
        self.active, self.incoming = bytes(self.incoming), None
        self.adopted += 1

    def signal_stream(self, stream):
        pass


def _worker_cpu(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tloam_b200 import multi
    n_map = (1000, 200, 2100, 1700)
    ok = True
    for epoch in range(3):                                                   # one collective per map epoch, no handshake
        payload = np.random.default_rng(7 + epoch).integers(0, 256, HostMapStandIn.layout_bytes(n_map), dtype=np.uint8).tobytes()
        if epoch == 0:
            reg = HostMapStandIn(payload if rank == 0 else None)
            builder = HostMapStandIn(payload) if rank == 0 else None
            chan = multi.SharedMapChannel(reg, src=0, builder=builder)       # separate builder: the source receives a local copy
        elif rank == 0:
            builder.blob = (ctypes.c_uint8 * len(payload)).from_buffer_copy(payload)
        n = chan.broadcast(n_map)
        chan.adopt()
        ok = ok and n == len(payload) and reg.active == payload and reg.adopted == epoch + 1
    # source registers against its own blob: nothing to receive, adopt is a no-op
    own = HostMapStandIn(payload if rank == 0 else None)
    chan2 = multi.SharedMapChannel(own, src=0)
    chan2.broadcast(n_map)
    chan2.adopt()
    ok = ok and ((rank == 0 and own.adopted == 0) or (rank != 0 and own.active == payload))
    fps, ms = multi.aggregate_frames_per_sec(30, 10.0 + 5.0 * rank)          # rank 1 is slower
    ok = ok and abs(ms - 15.0) < 1e-12 and abs(fps - 60 / 15e-3) < 1e-6
    ok = ok and multi.sequence_for_rank(rank) == ("00", "02")[rank]
    out.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_shared_map_broadcast_and_aggregation_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_worker_cpu, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == {0: True, 1: True}


def test_single_process_aggregation():
    from tloam_b200 import multi
    fps, ms = multi.aggregate_frames_per_sec(20, 10.0)
    assert fps == 2000.0 and ms == 10.0
    assert [multi.sequence_for_rank(r) for r in range(9)] == ["00", "02", "05", "08", "01", "06", "07", "09", "00"]


def _worker_gpu(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import tloam_b200
    from tloam_b200 import multi, synth
    cfg = synth.scaled(0.03, seed=11)
    T_gt = synth.se3_exp([1.0, 2.0, 0.0, 0.0, 0.0, 0.2])
    scan = synth.make_scan(cfg, T_gt, 0)
    predict = T_gt @ synth.se3_exp(synth.CONFIG1_PERTURB)
    reg = tloam_b200.LocalRegistration(device=rank)
    mp_ = synth.make_map(cfg, T_gt)
    if rank == 0:
        reg.set_input_target(mp_)                                # only the source rank builds the map
    chan = multi.SharedMapChannel(reg, src=0, device=torch.device("cuda", rank))
    chan.broadcast([len(c) for c in mp_])                        # ONE collective; the counts define the layout on every rank
    chan.adopt()
    reg.set_input_source(scan)
    T = reg.scan_matching(predict)
    gathered = [torch.zeros(16, dtype=torch.float64, device="cuda") for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(T.reshape(-1)).cuda())
    same = all(torch.equal(gathered[0], g) for g in gathered)      # bit-identical poses on every GPU
    out.put((rank, bool(same)))
    reg.close()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_shared_map_broadcast_nccl_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_worker_gpu, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == {0: True, 1: True}
