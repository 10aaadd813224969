"""Worker for tests/test_multigpu_gpu.py and for manual runs under torchrun (world_size >= 2, one GPU per rank):
sample-sharded frames gathered on rank 0 by (a) the core's peer-memory collective in both layouts (root gather, reduce-scatter) and
(b) the NCCL reduce must all equal one GPU rendering all samples (float sums in a different order: relative 1e-5)."""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lighthouse2_b200 import RenderCore, scenes  # noqa: E402
from lighthouse2_b200.distributed import PeerGatherRenderer, PipelinedShardedRenderer, TileShardedRenderer  # noqa: E402

W, H, SPP = 160, 90, 2


def make_core(dev, sd, spp):
    core = RenderCore(dev)
    core.SetTarget(W, H, spp)
    core.Setting("epsilon", 1e-3)
    core.Setting("maxPathLength", 3)
    sd.upload(core)
    return core


def tile_sharding(rank, world, local, sd):
    """Tile (row-band) sharded frames, plain and with the SVGF / TAA chain, moving camera: rank 0's image must equal the
    single-GPU frame bit for bit (1 spp). Filter mode three ways: tail on rank 0; filter chain sharded too (every rank filters a
    band, history rows read from their owners over NVLink) with interleaved and with contiguous rendered rows - those two also with
    converging frames in the sequence."""
    TW, TH = 192, 240
    views = [scenes.view_pyramid((3 * k, 30, -80 + k), (0, 0, 0), 40, TW, TH) for k in range(6)]
    ok = True
    #        filter, filter_shard, interleave, converge flags
    cases = [(0, 0, 1, [1] * 6), (1, 0, 1, [1] * 6), (1, 1, 1, [1] * 6), (1, 1, 0, [1] * 6), (1, 1, 1, [1, 1, 0, 0, 1, 0])]
    if os.environ.get("LH2B_WORKER_CASE"):
        cases = [cases[int(os.environ["LH2B_WORKER_CASE"])]]
    for filt, shard, inter, conv in cases:
        seq = []
        for k, c in enumerate(conv):
            seq.append(views[k] if c == 1 else seq[-1])     # a converging frame keeps the view of the frame before
        def make(spp=1):
            c = RenderCore(local)
            c.SetTarget(TW, TH, spp)
            c.Setting("epsilon", 1e-3)
            c.Setting("maxPathLength", 3)
            c.Setting("filter", filt)
            c.Setting("TAA", filt)
            sd.upload(c)
            return c
        want = []
        if rank == 0:
            single = make()
            for v, c in zip(seq, conv):
                single.Render(v, c)
                want.append(single.ReadPixels().copy())
            single.Shutdown()
        core = make()
        r = TileShardedRenderer(core, rank, world, filter_shard=shard, interleave=inter)
        outs = [torch.zeros((TH, TW, 4), dtype=torch.float32).pin_memory() for _ in views]
        for k, (v, c) in enumerate(zip(seq, conv)):
            r.frame(v, c, outs[k])
        r.finish()
        name = "plain" if not filt else ("filter, tail on rank 0" if not shard else f"filter, chain sharded, {'interleaved' if inter else 'contiguous'} rows, converge {conv}")
        if rank == 0:
            for k in range(len(views)):
                same = np.array_equal(outs[k].numpy(), want[k])
                err = np.abs(outs[k].numpy() - want[k]).max()
                bad_rows = np.nonzero((outs[k].numpy() != want[k]).any(axis=(1, 2)))[0]
                print(f"tile [{name}] frame {k}: rows {r.rows} identical={same} max abs err {err:.2e}" + (f" differing rows {bad_rows[:8]}..{bad_rows[-1]} ({len(bad_rows)})" if len(bad_rows) else ""), flush=True)
                ok &= bool(same)
                if not same and os.environ.get("LH2B_WORKER_VERBOSE"):
                    ys, xs = np.nonzero((outs[k].numpy() != want[k]).any(axis=2))
                    print(f"   {len(ys)} pixels differ; first: " + ", ".join(f"(y{y} x{x}: {outs[k].numpy()[y, x, :3]} vs {want[k][y, x, :3]})" for y, x in list(zip(ys, xs))[:6]), flush=True)
                    print("   per row: " + " ".join(f"{y}:{int((ys == y).sum())}" for y in np.unique(ys)[:40]), flush=True)
        r.close()
        core.Shutdown()
        dist.barrier()
    return ok


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    sd = scenes.config2_scene(48, 32, n_materials=4, light_quads=2, floaters=200)
    views = [scenes.view_pyramid((5 * k, 30, -80), (0, 0, 0), 40, W, H) for k in range(6)]
    conv = [1, 0, 0, 1, 0, 0]
    want = []
    if rank == 0:
        single = make_core(local, sd, SPP * world)      # the same samples on one GPU
        for v, c in zip(views, conv):
            single.Render(v, c)
            want.append(single.ReadPixels().copy())
        single.Shutdown()
    ok = True
    for kind in (() if os.environ.get("LH2B_WORKER_TILE_ONLY") else ("peer", "peer-reduce-scatter", "nccl")):
        core = make_core(local, sd, SPP)
        core.Setting("gatherMode", 1 if kind == "peer-reduce-scatter" else 0)
        r = PeerGatherRenderer(core, SPP, rank, world) if kind.startswith("peer") else PipelinedShardedRenderer(core, SPP, rank, world, f"cuda:{local}")
        outs = [torch.zeros((H, W, 4), dtype=torch.float32).pin_memory() for _ in views]
        for k, (v, c) in enumerate(zip(views, conv)):
            r.frame(v, c, outs[k])
        r.finish()
        if rank == 0:
            for k in range(len(views)):
                got = outs[k].numpy()
                err = np.abs(got - want[k]).max() / max(1e-6, np.abs(want[k]).max())
                print(f"{kind} frame {k}: max rel err {err:.2e}", flush=True)
                ok &= bool(err < 1e-5)
        if kind.startswith("peer"):
            r.close()
        core.Shutdown()
        dist.barrier()
    ok &= tile_sharding(rank, world, local, sd)
    flag = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    if rank == 0:
        print("MULTIGPU_OK" if ok else "MULTIGPU_FAIL", flush=True)
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
