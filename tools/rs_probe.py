"""Debug helper: tinyapp through the reference RenderSystem on our core vs the CPU oracle; dumps both frames."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_rendersystem_dropin as t
from oracle import binding as orc
W, H = 640, 360
os.makedirs("gpurun_out", exist_ok=True)
st = t.run_host(t.CORE, "/tmp/frame.bin", frames=1, w=W, h=H)
got = np.fromfile("/tmp/frame.bin", np.float32).reshape(H, W, 4)
t.run_host(t.RECORDER, "/tmp/none.bin", frames=1, w=W, h=H, record="/tmp/scene.rec")
sd, info = orc.load_recording("/tmp/scene.rec")
with orc.accel(1):
    o = orc.FrameOracle(sd, W, H, 1, 1e-3, 10.0, 3, 1)
    want, rec = o.render(info["view"], 1, records=True)
d = np.abs(got[..., :3] - want[..., :3]).max(axis=2)
bad = np.argwhere(d > 1e-3 * np.maximum(1.0, np.abs(want[..., :3]).max(axis=2)))
print(st, "bad pixels", len(bad))
for y, x in bad[:40]:
    r = rec[x + y * W]
    print(x, y, got[y, x, :3], want[y, x, :3], "hit", r["hit"][1:3], )
np.savez_compressed("gpurun_out/rs_probe.npz", got=got, want=want)
