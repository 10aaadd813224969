"""Isolate the cost of the top level and of fused generation: same 1M-triangle terrain frame (path length 1) with
(a) a point light -> single identity instance, (b) the light quad as second instance -> two-level."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lighthouse2_b200 import RenderCore, scenes, abi

W, H = 1920, 1080
view = scenes.view_pyramid((0, 30, -80), (0, 0, 0), 40, W, H)
for mode in ("point-light/single-level", "quad-light/two-level"):
    sd = scenes.config2_scene(1000, 500, n_materials=1, light_quads=1)
    if mode.startswith("point"):
        sd.meshes, sd.instances = sd.meshes[:1], sd.instances[:1]
        sd.tri_lights = sd.tri_lights[:0]
        pl = np.zeros(1, abi.CorePointLight); pl["position"] = (0, 26, 0); pl["radiance"] = (4000, 4000, 3200); pl["energy"] = 11200
        sd.point_lights = pl
    core = RenderCore(0)
    core.SetTarget(W, H, 1); core.Setting("epsilon", 1e-3); core.Setting("maxPathLength", 1)
    for k, v in os.environ.items():
        if k.startswith("LH2B_SET_"):
            core.Setting(k[9:], float(v))
    sd.upload(core)
    acc = {}
    for i in range(12):
        core.Render(view, 1)
        fs = core.GetFrameStats()
        if i >= 2:
            for k in ("generateExtendMs", "shadeMs", "connectMs", "shadowRays"):
                acc[k] = acc.get(k, 0) + float(fs[k]) / 10
    print(mode, {k: round(v, 4) for k, v in acc.items()})
    core.Shutdown()
