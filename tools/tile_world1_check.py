"""Sanity check of the tile-sharding path with a single rank (torchrun --nproc-per-node 1): create / frame / deferred tail, plain and
filter + TAA, against the ordinary frames of the same core."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.distributed as dist
import multigpu_worker as w
from lighthouse2_b200 import scenes
torch.cuda.set_device(0)
dist.init_process_group("nccl", device_id=torch.device("cuda:0"))
sd = scenes.config2_scene(48, 32, n_materials=4, light_quads=2, floaters=200)
ok = w.tile_sharding(0, 1, 0, sd)
print("TILE1_OK" if ok else "TILE1_FAIL")
dist.destroy_process_group()
