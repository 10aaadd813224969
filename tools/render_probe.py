"""Render sanity probe: config-2/3 style frame through the CoreAPI mirror, prints stats."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lighthouse2_b200 import RenderCore, scenes

nx, nz, W, H, spp, maxlen = [int(a) for a in (sys.argv[1:7] if len(sys.argv) > 6 else (200, 100, 640, 360, 1, 3))]
sd = scenes.config2_scene(nx, nz, n_materials=6, light_quads=2)
core = RenderCore(0)
core.SetTarget(W, H, spp)
core.Setting("epsilon", 1e-3); core.Setting("maxPathLength", maxlen)
t0 = time.time(); sd.upload(core); print("upload+build %.2fs" % (time.time() - t0))
view = scenes.view_pyramid((0, 30, -80), (0, 0, 0), 40, W, H)
for i in range(3):
    core.Render(view, 1)
    fs = core.GetFrameStats(); st = core.GetCoreStats()
    print(i, {k: fs[k].item() for k in fs.dtype.names if k != "reserved"})
img = core.ReadPixels()
print("mean rgb", img[..., :3].mean(axis=(0, 1)), "max", img[..., :3].max(), "nan", np.isnan(img).sum())
print("probe", st["probedInstid"], st["probedTriid"], st["probedDist"])
np.save("gpurun_out/render_probe.npy", img) if os.path.isdir("gpurun_out") else None
