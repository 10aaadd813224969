"""BASELINE.json configs[4]: 4K, 1 spp, SVGF filter + TAA, moving camera, real-time frame-time mode at N GPUs (tile-sharded frames,
csrc/tile_gather.cu). Run under torchrun (N >= 2) or plainly (N = 1):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/run_c5_scaling.py [out.json] [frames]
Prints / writes ms per frame (device time on rank 0's stream around K pipelined frames, every frame read back to pinned host
memory) and frames per second."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from lighthouse2_b200 import RenderCore, scenes
from lighthouse2_b200.distributed import TileShardedRenderer

out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/c5_scaling.json"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 40
readback = (sys.argv[3] if len(sys.argv) > 3 else "readback") == "readback"   # "device": the finished frame stays in rank 0's pixel buffer
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
W, H = 3840, 2160
sd = scenes.config2_scene(1000, 500, n_materials=64, light_quads=8)
core = RenderCore(local)
core.SetTarget(W, H, 1)
core.Setting("epsilon", 1e-3), core.Setting("filter", 1), core.Setting("TAA", 1)
for k, v in os.environ.items():
    if k.startswith("LH2B_SET_"):
        core.Setting(k[9:], float(v))
sd.upload(core)
host = [torch.empty((H, W, 4), dtype=torch.float32, pin_memory=True) for _ in range(2)]
view_of = lambda f: scenes.view_pyramid((0.2 * f, 30, -80 + 0.1 * f), (0, 0, 0), 40, W, H)
r = TileShardedRenderer(core, rank, world) if world > 1 else None
if r is None:
    core.Setting("pipeline", 1)


def run(n, f0):
    for f in range(n):
        if r is not None:
            r.frame(view_of(f0 + f), 1, host[f & 1] if (rank == 0 and readback) else None)
        else:
            core.Render(view_of(f0 + f), 1, True)
            if readback:
                core.ReadPixelsAsync(host[f & 1].numpy())
    if r is not None:
        r.finish()
    else:
        core.WaitForRender(), core.WaitReadPixels()


run(6, 0)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
run(frames, 6)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
dt = time.perf_counter() - t0
if rank == 0:
    res = {"config": "C5: 3840x2160, 1 spp, path length 3, filter + TAA, moving camera, 1,000,016 triangles, 64 materials", "n_gpus": world,
           "frames": frames, "ms_per_frame": dt / frames * 1e3, "fps": frames / dt,
           "frame_ends": "in pinned host memory (133 MB RGBA32F over PCIe per frame)" if readback else "in rank 0's device pixel buffer (what a GL / display path consumes)",
           "sharding": "tile (row bands), bands gathered on rank 0 over NVLink peer memory, filter chain on rank 0" if world > 1 else "none",
           "rows_rank0": list(r.rows) if r is not None else [0, H], "image_mean": float(host[(frames - 1) & 1].numpy()[..., :3].mean()) if readback else float(core.ReadPixels()[..., :3].mean())}
    print(json.dumps(res), flush=True)
    os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
    json.dump(res, open(out_path, "w"), indent=1)
if r is not None:
    r.close()
core.Shutdown()
if world > 1:
    dist.destroy_process_group()
