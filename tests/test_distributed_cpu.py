"""CPU, world_size 2 over gloo: the multi-GPU host logic (lighthouse2_b200/distributed.py) with a CPU stand-in core
backed by the oracle. Two ranks render their sample shards, reduce, and rank 0 must get the image a single process
renders with all samples - i.e. sample sharding reproduces the single-GPU random sequence."""
import os
import sys
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H, SPP = 48, 32, 2


def _scene():
    from lighthouse2_b200 import scenes
    sd = scenes.config2_scene(16, 12, n_materials=3, light_quads=1, floaters=40)
    view = scenes.view_pyramid((0, 30, -80), (0, 0, 0), 40, W, H)
    return sd, view


class OracleCore:
    """The surface ShardedRenderer uses, implemented with the CPU oracle."""

    def __init__(self, sd):
        self.sd, self.width, self.height = sd, W, H
        self.first, self.total = 0, 0
        self.oracle = None
        self.acc = torch.zeros((H, W, 4), dtype=torch.float32)
        self.pixels = None

    def SetSampleShard(self, first, total):
        from oracle import binding as orc
        self.first, self.total = first, total
        self.oracle = orc.FrameOracle(self.sd, W, H, SPP, 1e-3, 10.0, 3, 1, threads=2, sample_base=first, total_spp=total)

    def Render(self, view, converge):
        self.oracle.render(view, converge)
        self.acc.copy_(torch.from_numpy(self.oracle.accum))

    def SamplesTaken(self):
        return self.oracle.samples_taken

    def FinalizeExternal(self, acc, samples):
        self.pixels = (acc / samples).numpy()


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lighthouse2_b200.distributed import ShardedRenderer, sample_shard
    sd, view = _scene()
    core = OracleCore(sd)
    r = ShardedRenderer(core, SPP, rank, world, core.acc)
    assert (core.first, core.total) == sample_shard(rank, world, SPP) == (rank * SPP, world * SPP)
    total = 0
    for conv in (1, 0, 0):        # Restart, first converging frame (restarts too, rendercore.cpp:827-833), Converge
        total = r.render(view, conv)
    assert total == 2 * world * SPP
    r.finalize()
    if rank == 0:
        np.save(out, core.pixels)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sample_sharding_matches_single_process(tmp_path):
    from oracle import binding as orc
    out = str(tmp_path / "sharded.npy")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    sd, view = _scene()
    single = orc.FrameOracle(sd, W, H, 2 * SPP, 1e-3, 10.0, 3, 1, threads=2)
    single.render(view, 1)
    single.render(view, 0)
    want = single.render(view, 0)
    assert single.samples_taken == 4 * SPP
    assert got.shape == want.shape
    np.testing.assert_allclose(got[..., :3], want[..., :3], rtol=2e-5, atol=2e-6)


def test_sample_shard_partition():
    from lighthouse2_b200.distributed import sample_shard
    for world in (1, 2, 4, 8):
        for spp in (1, 2, 16):
            covered = []
            for r in range(world):
                first, total = sample_shard(r, world, spp)
                assert total == world * spp
                covered += list(range(first, first + spp))
            assert covered == list(range(world * spp))
