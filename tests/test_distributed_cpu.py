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


# ---- tile sharding (TileShardedRenderer): the host-side protocol over gloo with a stand-in core -----------------------------------------
class _TileStubCore:
    """Records what TileShardedRenderer asks of a core; handles are per-rank byte strings so that the exchange order can be checked."""

    def __init__(self, rank):
        self.rank, self.calls, self.settings, self.imported = rank, [], {}, None

    def Setting(self, name, value):
        self.settings[name] = value
        self.calls.append(("Setting", name, value))

    def TileCreate(self, rank, world):
        self.calls.append(("TileCreate", rank, world))
        return "gatherer"

    def TileExport(self, g):
        return bytes([0xA0 + self.rank]) * 8

    def TileImport(self, g, blob):
        self.imported = blob

    def TileRows(self, g):
        return (4 * self.rank, 240, 2)

    def Render(self, view, converge, async_):
        self.calls.append(("Render", converge, async_))

    def TileFrame(self, g):
        self.calls.append(("TileFrame",))

    def ReadPixelsAsync(self, out):
        self.calls.append(("ReadPixelsAsync",))

    def WaitForRender(self):
        self.calls.append(("WaitForRender",))

    def TileWait(self, g):
        self.calls.append(("TileWait",))

    def WaitReadPixels(self):
        self.calls.append(("WaitReadPixels",))

    def TileDestroy(self, g):
        self.calls.append(("TileDestroy",))


def _tile_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lighthouse2_b200.distributed import TileShardedRenderer
    core = _TileStubCore(rank)
    r = TileShardedRenderer(core, rank, world, filter_shard=1, interleave=1)
    # the settings reach the core before the gatherer is created (lh2b_tile_create reads them), pipelined frames are switched on
    names = [c[1] for c in core.calls if c[0] == "Setting"]
    assert set(names) >= {"pipeline", "tileFilterShard", "tileInterleave"}
    assert [c[0] for c in core.calls].index("TileCreate") > max(i for i, c in enumerate(core.calls) if c[0] == "Setting")
    # every rank imports every rank's handles, concatenated in rank order
    assert core.imported == b"".join(bytes([0xA0 + q]) * 8 for q in range(world))
    assert r.rows == (4 * rank, 240, 2)
    host = np.zeros((4, 4, 4), np.float32)
    for conv in (1, 0):
        r.frame("view", conv, host)
    frame_calls = [c[0] for c in core.calls if c[0] in ("Render", "TileFrame", "ReadPixelsAsync")]
    assert frame_calls == (["Render", "TileFrame", "ReadPixelsAsync"] * 2 if rank == 0 else ["Render", "TileFrame"] * 2)    # only rank 0 holds the frame
    assert [c for c in core.calls if c[0] == "Render"] == [("Render", 1, True), ("Render", 0, True)]
    r.close()
    tail = [c[0] for c in core.calls][-4:] if rank == 0 else [c[0] for c in core.calls][-3:]
    assert tail == (["WaitForRender", "TileWait", "WaitReadPixels", "TileDestroy"] if rank == 0 else ["WaitForRender", "TileWait", "TileDestroy"])
    if rank == 0:
        open(out, "w").write("ok")
    dist.destroy_process_group()


def test_two_rank_tile_sharding_host_protocol(tmp_path):
    out = str(tmp_path / "tile_ok.txt")
    port = 31500 + os.getpid() % 2000
    mp.spawn(_tile_worker, args=(2, port, out), nprocs=2, join=True)
    assert open(out).read() == "ok"
