"""Product in filter mode vs the CPU FilteredFrameOracle over a frame sequence (moving camera, TAA): prints per-frame agreement."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lighthouse2_b200 import RenderCore, scenes
from oracle import binding as orc
W, H = 160, 96
for taa in (1, 0):
    sd = scenes.config2_scene(48, 32, n_materials=6, light_quads=2, floaters=300)
    core = RenderCore(); core.SetTarget(W, H, 1); core.Setting("epsilon", 1e-3); core.Setting("filter", 1); core.Setting("TAA", taa)
    core.Setting("clampDirect", 15.0); core.Setting("clampIndirect", 15.0)
    for k_, v_ in os.environ.items():
        if k_.startswith("LH2B_SET_"):
            core.Setting(k_[9:], float(v_))
    sd.upload(core)
    views = [scenes.view_pyramid((0.4 * k, 30 + 0.1 * k, -80 + 0.3 * k), (0, 0, 0), 40, W, H) for k in range(5)] + [scenes.view_pyramid((1.6, 30.4, -78.8), (0, 0, 0), 40, W, H)] * 2
    with orc.accel(1):
        fo = orc.FilteredFrameOracle(sd, W, H, taa=bool(taa))
        for k, v in enumerate(views):
            core.Render(v, 1)
            got = core.ReadPixels()[..., :3]
            want = fo.render(v, 1)[..., :3]
            inner = (slice(16, H - 16), slice(16, W - 16))
            d = np.abs(got[inner] - want[inner])
            rel = np.sqrt(((got[inner] - want[inner]) ** 2).mean()) / np.sqrt((want[inner] ** 2).mean())
            print(f"taa {taa} frame {k}: >1e-2 {float((d > 1e-2).any(-1).mean()):.4f} >3e-2 {float((d > 3e-2).any(-1).mean()):.4f} >1e-1 {float((d > 1e-1).any(-1).mean()):.4f} relRMSE {rel:.4f} "
                  f"means {got[inner].mean():.4f} {want[inner].mean():.4f} border>3e-2 {float((np.abs(got - want) > 3e-2).any(-1).mean()):.4f}", flush=True)
    core.Shutdown()
