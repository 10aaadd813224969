"""A few frames of one BASELINE.json configuration, for ncu: python tools/profile_target.py c2|c3|c4|c5|build [frames]
(c3: 4 spp instead of 16 so that a replayed kernel stays short; c4: 4 x 1M-triangle meshes, 400 instances + per-frame refit)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lighthouse2_b200 import RenderCore, scenes

which = sys.argv[1] if len(sys.argv) > 1 else "c2"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 3
W, H = 1920, 1080
core = RenderCore(0)
for k, v in os.environ.items():
    if k.startswith("LH2B_SET_"):
        core.Setting(k[9:], float(v))
if which in ("c2", "c3", "c5", "build"):
    if which == "c5":
        W, H = 3840, 2160
    sd = scenes.config2_scene(1000, 500, n_materials=1 if which in ("c2", "build") else 64, light_quads=1 if which in ("c2", "build") else 8)
    core.SetTarget(W, H, 4 if which == "c3" else 1)
    core.Setting("epsilon", 1e-3)
    if which in ("c2", "build"):
        core.Setting("maxPathLength", 1)
    if which == "c3":
        core.Setting("maxPathLength", 8), core.Setting("maxDiffuseBounces", 2)
    if which == "c5":
        core.Setting("filter", 1), core.Setting("TAA", 1)
    sd.upload(core)
    if which == "build":
        # a second full build (new triangle count) and a refit (same count) of the 1M-triangle mesh
        v, t = sd.meshes[0]
        core.SetGeometry(0, v[:-3], t[:-1]); core.FinalizeInstances()
        v2 = v[:-3].copy(); v2[:, 1] += 0.01
        core.SetGeometry(0, v2, None); core.FinalizeInstances()
    for f in range(frames):
        view = scenes.view_pyramid((0.2 * f, 30, -80 + 0.1 * f), (0, 0, 0), 40, W, H)
        core.Render(view, 1)
elif which == "c4":
    core.SetTarget(W, H, 1); core.Setting("epsilon", 1e-3)
    base = scenes.terrain(1000, 500, extent=6.0, seed=5)
    base[:, 1] *= 0.3
    mats = scenes.make_materials([dict(color=(0.7, 0.7, 0.7)), dict(color=(80, 80, 64))])
    core.SetSkyData(*scenes.gradient_sky())
    core.SetMaterials(mats)
    tris = scenes.core_tris_from_verts(base)
    NM, NI = 4, 400
    for m in range(NM):
        core.SetGeometry(m, base, tris)
    lq = scenes.quad((0, 60, 0), (0, -1, 0), 30, 30); lt = scenes.core_tris_from_verts(lq, material=1)
    core.SetGeometry(NM, lq, lt)
    core.SetLights(scenes.tri_lights(lq, lt, mats, inst_idx=NI))
    rng = np.random.default_rng(3)
    pos = (rng.random((NI, 3)) * 2 - 1) * np.array([90, 25, 90])
    view = scenes.view_pyramid((0, 60, -200), (0, 0, 0), 45, W, H)
    for f in range(frames):
        for i in range(NI):
            a = 0.01 * f * (1 + i % 7)
            m = np.eye(4, dtype=np.float32)
            m[0, 0], m[0, 2], m[2, 0], m[2, 2] = np.cos(a), np.sin(a), -np.sin(a), np.cos(a)
            m[:3, 3] = pos[i]
            core.SetInstance(i, i % NM, m)
        core.SetInstance(NI, NM); core.SetInstance(NI + 1, -1)
        if f > 0:
            moved = base.copy(); moved[:, 1] += (0.2 * np.sin(base[:, 0] * 2 + f)).astype(np.float32)
            for m in range(NM):
                core.SetGeometry(m, moved, None)
        core.FinalizeInstances()
        core.Render(view, 1)
print(which, "done", core.GetFrameStats()["totalMs"])
core.Shutdown()
