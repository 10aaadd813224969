"""compute-sanitizer target: a filtered frame or two of the C5 workload (or a small version of it).
    compute-sanitizer --tool initcheck python tools/initcheck_probe.py [4k]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lighthouse2_b200 import RenderCore, scenes
big = len(sys.argv) > 1 and sys.argv[1] == "4k"
W, H = (3840, 2160) if big else (320, 184)
sd = scenes.config2_scene(1000, 500, n_materials=64, light_quads=8) if big else scenes.config2_scene(100, 50, n_materials=64, light_quads=8)
c = RenderCore(0)
c.SetTarget(W, H, 1)
c.Setting("epsilon", 1e-3), c.Setting("filter", 1), c.Setting("TAA", 1)
sd.upload(c)
for k in range(2):
    c.Render(scenes.view_pyramid((0.2 * k, 30, -80 + 0.1 * k), (0, 0, 0), 40, W, H), 1)
print("done", c.ReadPixels().mean())
