"""Where do two core instances of one process stop agreeing bit for bit? Runs a 4K frame of the C5 scene stage by stage through the parity
hooks (ray queries, the shade stage, shadow-ray queries) on two cores with IDENTICAL host inputs and compares every output bitwise.
    python tools/stage_probe.py [filter]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lighthouse2_b200 import RenderCore, scenes

W, H = 3840, 2160
filt = 1 if (len(sys.argv) > 1 and sys.argv[1] == "filter") else 0
sd = scenes.config2_scene(1000, 500, n_materials=64, light_quads=8)
view = scenes.view_pyramid((0, 30, -80), (0, 0, 0), 40, W, H)


def make():
    c = RenderCore(0)
    c.SetTarget(W, H, 1)
    c.Setting("epsilon", 1e-3), c.Setting("filter", filt), c.Setting("TAA", filt)
    for k, v in os.environ.items():
        if k.startswith("LH2B_SET_"):
            c.Setting(k[9:], float(v))
    sd.upload(c)
    c.Render(view, 1)        # fixes the view-dependent parameters of the shade hook
    return c


def bits(x):
    return np.ascontiguousarray(x).view(np.uint32)


def report(what, x, y):
    same = x.shape == y.shape and np.array_equal(bits(x), bits(y))
    extra = ""
    if not same and x.shape == y.shape:
        rows = np.nonzero((bits(x) != bits(y)).reshape(x.shape[0], -1).any(axis=1))[0]
        extra = f": {len(rows)} of {x.shape[0]} entries differ, first {rows[0]}: {x.reshape(x.shape[0], -1)[rows[0]]} vs {y.reshape(y.shape[0], -1)[rows[0]]}"
    elif not same:
        extra = f": shapes {x.shape} vs {y.shape}"
    print(f"{what}: identical {same}{extra}", flush=True)
    return same


a, b = make(), make()
if not filt:
    report("whole frame, accumulator", a.ReadAccumulator().reshape(-1, 4), b.ReadAccumulator().reshape(-1, 4))
O, D = scenes.camera_rays(view, W, H)
n = W * H
O4, D4 = O.copy(), D.copy()
O4[:, 3] = ((np.arange(n, dtype=np.uint32) << 6) | 1).view(np.float32)
T4 = np.ones((n, 4), np.float32)
shift = 0x5A17C3E1
for L in (1, 2, 3):
    ha, hb = a.TraceRays(O4, D4), b.TraceRays(O4, D4)
    report(f"L{L} closest hits ({len(O4)} rays)", ha, hb)
    R0 = (0x9E3779B9 * L + L * 91771) & 0xFFFFFFFF
    acc0 = np.zeros((H, W, 4), np.float32)
    (ea, sa, aa), (eb, sb, ab) = a.ShadePaths(L, O4, D4, T4, ha, R0, shift, 0, acc0), b.ShadePaths(L, O4, D4, T4, ha, R0, shift, 0, acc0)
    ia, ib = np.argsort(ea["O"][:, 3].view(np.uint32) >> 6, kind="stable"), np.argsort(eb["O"][:, 3].view(np.uint32) >> 6, kind="stable")
    for f in ("O", "D", "T"):
        report(f"L{L} shade: extension rays {f}", ea[f][ia], eb[f][ib])
    ja, jb = np.argsort(sa["E"][:, 3].view(np.uint32), kind="stable"), np.argsort(sb["E"][:, 3].view(np.uint32), kind="stable")
    for f in ("O", "D", "E"):
        report(f"L{L} shade: shadow rays {f}", sa[f][ja], sb[f][jb])
    report(f"L{L} shade: accumulator (first half)", aa.reshape(-1, 4), ab.reshape(-1, 4))
    if len(sa["O"]):
        report(f"L{L} shadow-ray occlusion ({len(sa['O'])} rays)", a.TraceShadowRays(sa["O"][ja], sa["D"][ja]).reshape(-1, 1).astype(np.uint32), b.TraceShadowRays(sa["O"][ja], sa["D"][ja]).reshape(-1, 1).astype(np.uint32))
    O4, D4, T4 = ea["O"][ia], ea["D"][ia], ea["T"][ia]
    if len(O4) == 0:
        break
