"""Is a 4K filter-mode frame of the C5 workload reproducible bit for bit (a) on the same core, twice, (b) on a second core that built
its own acceleration structure, (c) with another BVH builder? Prints the number of differing accumulator pixels.
    python tools/determinism_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lighthouse2_b200 import RenderCore, scenes

W, H = 3840, 2160
sd = scenes.config2_scene(1000, 500, n_materials=64, light_quads=8)
views = [scenes.view_pyramid((0.2 * k, 30, -80 + 0.1 * k), (0, 0, 0), 40, W, H) for k in range(3)]


def make(builder=None):
    c = RenderCore(0)
    c.SetTarget(W, H, 1)
    c.Setting("epsilon", 1e-3), c.Setting("filter", 1), c.Setting("TAA", 1)
    if builder is not None:
        c.Setting("bvhBuilder", builder)
    for k, v in os.environ.items():
        if k.startswith("LH2B_SET_"):
            c.Setting(k[9:], float(v))
    sd.upload(c)
    return c


def frames(c):
    out = []
    for v in views:
        c.Render(v, 1)
        out.append(c.ReadFilterBuffers()[3].copy())
    return out


def diff(a, b, what):
    for k, (x, y) in enumerate(zip(a, b)):
        d = (x.view(np.uint32) != y.view(np.uint32)).any(axis=(0, 3))
        ys, xs = np.nonzero(d)
        print(f"{what}: frame {k}: {len(ys)} accumulator pixels differ" + (f", e.g. (y{ys[0]} x{xs[0]}): {x[:, ys[0], xs[0]].ravel()} vs {y[:, ys[0], xs[0]].ravel()}" if len(ys) else ""), flush=True)


a = make()
fa1 = frames(a)
fa2 = frames(a)
diff(fa1, fa2, "same core, run twice")
b = make()
diff(fa1, frames(b), "second core, own BVH build")
