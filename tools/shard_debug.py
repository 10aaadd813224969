"""Debug aid for the sharded filter chain (csrc/tile_gather.cu): under torchrun, frame by frame, every rank compares the rows of its
band in every filter buffer with one GPU rendering the whole frame, and reports the first buffers / pixels that differ.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 tools/shard_debug.py [interleave]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from lighthouse2_b200 import RenderCore, scenes
from lighthouse2_b200.distributed import TileShardedRenderer

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
inter = int(sys.argv[1]) if len(sys.argv) > 1 else 1
if os.environ.get("SHARD_DEBUG_4K"):       # the C5 workload itself
    TW, TH = 3840, 2160
    sd = scenes.config2_scene(1000, 500, n_materials=64, light_quads=8)
    views = [scenes.view_pyramid((0.2 * k, 30, -80 + 0.1 * k), (0, 0, 0), 40, TW, TH) for k in range(int(os.environ.get("SHARD_DEBUG_FRAMES", "5")))]
else:
    TW, TH = 192, 240
    sd = scenes.config2_scene(48, 32, n_materials=4, light_quads=2, floaters=200)
    views = [scenes.view_pyramid((3 * k, 30, -80 + k), (0, 0, 0), 40, TW, TH) for k in range(4)]


def make():
    c = RenderCore(local)
    c.SetTarget(TW, TH, 1)
    c.Setting("epsilon", 1e-3), c.Setting("maxPathLength", 3), c.Setting("filter", 1), c.Setting("TAA", 1)
    for key, val in os.environ.items():         # LH2B_SET_<setting>=<value>, e.g. bvhBuilder=1: the deterministic host builder gives every core the same tree
        if key.startswith("LH2B_SET_"):
            c.Setting(key[9:], float(val))
    sd.upload(c)
    return c


single, core = make(), make()           # every rank renders the whole frame itself as the reference
r = TileShardedRenderer(core, rank, world, filter_shard=1, interleave=inter)
rpb = ((TH + world - 1) // world + 15) // 16 * 16
b0, b1 = rank * rpb, min(TH, (rank + 1) * rpb)
for k, v in enumerate(views):
    single.Render(v, 1)
    r.frame(v, 1, None)
    r.finish()
    dist.barrier()
    want = dict(zip(("features", "worldPos", "deltaDepth", "accumulator"), single.ReadFilterBuffers()))
    want.update(single.ReadFilterHistory())
    got = dict(zip(("features", "worldPos", "deltaDepth", "accumulator"), core.ReadFilterBuffers()))
    got.update(core.ReadFilterHistory())
    lines = []
    img_single, img = single.ReadPixels(), None
    if rank == 0:
        img = core.ReadPixels()
        d = (img.view(np.uint32) != img_single.view(np.uint32)).any(axis=2)
        if d.any():
            ys, xs = np.nonzero(d)
            print(f"frame {k} FINAL IMAGE differs in {len(ys)} px, rows {ys.min()}..{ys.max()}; first (y{ys[0]} x{xs[0]}): {img[ys[0], xs[0]]} vs {img_single[ys[0], xs[0]]}", flush=True)
        else:
            print(f"frame {k} final image identical", flush=True)
    if k > 0 and (got["taa"][b0:b1].view(np.uint32) != want["taa"][b0:b1].view(np.uint32)).any():
        ys, xs = np.nonzero((got["taa"][b0:b1].view(np.uint32) != want["taa"][b0:b1].view(np.uint32)).any(axis=2))
        y, x = b0 + ys[0], xs[0]
        mv = got["motion"][y, x]
        pu, pv = mv[0] - 0.5, mv[1] - 0.5
        x1, y1 = int(pu - 2.0), int(pv - 2.0)
        fp_got, fp_want = prev_got["taa"][max(y1, 0):y1 + 4, max(x1, 0):x1 + 4], prev_want["taa"][max(y1, 0):y1 + 4, max(x1, 0):x1 + 4]
        print(f"[rank {rank} frame {k}] taa ({y},{x}): got {got['taa'][y, x]} want {want['taa'][y, x]} motion {mv} (single {want['motion'][y, x]}) footprint rows {y1}..{y1 + 3} cols {x1}..{x1 + 3} "
              f"prev-taa footprint identical: {np.array_equal(fp_got.view(np.uint32), fp_want.view(np.uint32))}; phase3 3x3 identical: "
              f"{np.array_equal(got['phase3'][y - 1:y + 2, max(x - 1, 0):x + 2].view(np.uint32), want['phase3'][y - 1:y + 2, max(x - 1, 0):x + 2].view(np.uint32))}", flush=True)
        if not np.array_equal(fp_got.view(np.uint32), fp_want.view(np.uint32)):
            print(f"   got footprint {fp_got[..., 0]}\n   want footprint {fp_want[..., 0]}", flush=True)
    prev_got, prev_want = got, want
    for name in ("accumulator", "features", "worldPos", "deltaDepth", "motion", "moments", "phase1", "phase3", "taa"):
        a, b = got[name], want[name]
        if name == "accumulator":
            a, b = a[:, b0:b1], b[:, b0:b1]
            diff = (a.view(np.uint32) != b.view(np.uint32)).any(axis=(0, 3))
        else:
            a, b = a[b0:b1], b[b0:b1]
            diff = (a.view(np.uint32) != b.view(np.uint32)).any(axis=2)
        if diff.any():
            ys, xs = np.nonzero(diff)
            lines.append(f"{name}: {len(ys)} px, rows {b0 + ys.min()}..{b0 + ys.max()}, x {xs.min()}..{xs.max()}; first (y{b0 + ys[0]} x{xs[0]})")
    for q in range(world):
        if q == rank:
            print(f"frame {k} rank {rank} band [{b0},{b1}): " + ("all identical" if not lines else " | ".join(lines)), flush=True)
        dist.barrier()
r.close()
dist.destroy_process_group()
