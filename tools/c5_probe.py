"""C5 frames (4K, filter + TAA, moving camera) for profiling: python tools/c5_probe.py [frames]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lighthouse2_b200 import RenderCore, scenes
W, H = 3840, 2160
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
sd = scenes.config2_scene(1000, 500, n_materials=64, light_quads=8)
core = RenderCore(0); core.SetTarget(W, H, 1); core.Setting("epsilon", 1e-3); core.Setting("filter", 1); core.Setting("TAA", 1)
for k, v in os.environ.items():
    if k.startswith("LH2B_SET_"):
        core.Setting(k[9:], float(v))
sd.upload(core)
acc = {}
for f in range(n):
    view = scenes.view_pyramid((0.2 * f, 30, -80 + 0.1 * f), (0, 0, 0), 40, W, H)
    core.Render(view, 1)
    fs = core.GetFrameStats()
    if f >= 2:
        for k in fs.dtype.names:
            if k != "reserved":
                acc[k] = acc.get(k, 0.0) + float(fs[k]) / (n - 2)
print({k: round(v, 4) for k, v in acc.items()})
core.Shutdown()
