"""Stage breakdown of the literal tinyapp scene (recorded from the reference RenderSystem) at several resolutions."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_rendersystem_dropin as rs
from oracle import binding as orc
from lighthouse2_b200 import RenderCore
for (W, H) in ((640, 360), (1920, 1080)):
    rs.run_host(rs.RECORDER, "/tmp/none.bin", frames=1, w=W, h=H, record="/tmp/c1.rec")
    sd, info = orc.load_recording("/tmp/c1.rec")
    core = RenderCore(0); core.SetTarget(W, H, 1); core.Setting("epsilon", 1e-3); core.Setting("clampValue", 10)
    for k, v in os.environ.items():
        if k.startswith("LH2B_SET_"):
            core.Setting(k[9:], float(v))
    sd.upload(core)
    acc = {}
    N = 12
    for i in range(N + 2):
        core.Render(info["view"], 1)
        fs = core.GetFrameStats()
        if i >= 2:
            for k in fs.dtype.names:
                if k != "reserved":
                    acc[k] = acc.get(k, 0.0) + float(fs[k]) / N
    print(W, H, {k: round(v, 4) for k, v in acc.items()}, "Mrays/s", (acc["extensionRays"] + acc["shadowRays"]) / acc["totalMs"] / 1e3, flush=True)
    for m in range(0, 172, 60):
        print("  bvh", m, core.GetBvhStats(m))
    core.Shutdown()
