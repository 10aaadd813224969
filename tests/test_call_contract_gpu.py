"""The call-order contract of the boundary (from lib/rendercore_optix7/rendercore.cpp): what the reference does at the
edges of its API, checked on the core through the C ABI and compared with the CPU oracle where an image results.

  - Render before the first FinalizeInstances is a silent no-op (gpuHasSceneData, rendercore.cpp:431,821)
  - SetTarget may come at any time, re-allocates and clears accumulation (rendercore.cpp:284-326)
  - SetInstance( idx, -1 ) truncates the instance list (rendercore.cpp:346-376)
  - SetGeometry on an existing mesh index with a different triangle count replaces the mesh (core_mesh.cpp:34-129: full rebuild)
  - unknown Setting names are ignored (rendercore.cpp:746-760)
  - a scene without lights and without sky renders black; counts stay consistent"""
import numpy as np
import pytest

from lighthouse2_b200 import RenderCore, scenes
from oracle import binding as orc
from util import rel_rmse, pixel_mismatch_fraction

pytestmark = pytest.mark.gpu
W, H = 128, 72


def _agree(core, sd, view, w=W, h=H, frames=1):
    o = orc.FrameOracle(sd, w, h, 1, 1e-3, 10.0, 3, 1)
    for _ in range(frames):
        want = o.render(view, 1)
    got = core.ReadPixels()
    assert np.isfinite(got).all() and rel_rmse(got, want) < 0.02 and pixel_mismatch_fraction(got, want) < 0.005
    return got


def test_render_before_finalize_is_a_noop_and_settarget_any_time():
    sd = scenes.config2_scene(32, 24, n_materials=3, light_quads=1, floaters=100)
    view = scenes.view_pyramid((0, 30, -80), (0, 0, 0), 40, W, H)
    core = RenderCore()
    core.SetTarget(W, H, 1)
    core.Setting("epsilon", 1e-3)
    core.Setting("definitelyNotASetting", 3.0)          # ignored
    core.Render(view, 1)                                 # nothing uploaded yet: must not fail, must not produce rays
    st = core.GetCoreStats()
    assert int(st["totalRays"]) == 0 and not core.ReadPixels().any()
    sd.upload(core)
    core.Render(view, 1)
    _agree(core, sd, view)                               # first real frame = the oracle's first frame (the no-op consumed no seeds)
    # new target size in the middle of a run: buffers follow, accumulation restarts
    w2, h2 = 96, 48
    view2 = scenes.view_pyramid((0, 30, -80), (0, 0, 0), 40, w2, h2)
    core.SetTarget(w2, h2, 1)
    core.Render(view2, 0)                                # Converge on a fresh target behaves like Restart (samplesTaken = 0)
    got = core.ReadPixels()
    assert got.shape[:2] == (h2, w2) and np.isfinite(got).all() and got[..., :3].mean() > 0
    assert int(core.GetCoreStats()["primaryRayCount"]) == w2 * h2
    core.Shutdown()


def test_instance_list_truncation_and_mesh_replacement():
    sd = scenes.config2_scene(32, 24, n_materials=3, light_quads=2, floaters=100)
    view = scenes.view_pyramid((0, 30, -80), (0, 0, 0), 40, W, H)
    core = RenderCore()
    core.SetTarget(W, H, 1)
    core.Setting("epsilon", 1e-3)
    sd.upload(core)
    core.Render(view, 1)
    _agree(core, sd, view)
    # replace the terrain mesh by one with a different triangle count (rebuild, not refit)
    tv = scenes.terrain(20, 14, 50.0, 99, 40)
    tt = scenes.core_tris_from_verts(tv)
    tt["material"] = 1
    sd.meshes[0] = (tv, tt)
    core.SetGeometry(0, tv, tt)
    core.FinalizeInstances()
    core.Render(view, 1)
    _agree(core, sd, view, frames=2)
    # drop the last instance (a light quad): its triangles no longer occlude or get hit; the light list is the host's business
    n = len(sd.instances)
    core.SetInstance(n - 1, -1)
    core.FinalizeInstances()
    sd.instances = sd.instances[:n - 1]
    core.Render(view, 1)
    _agree(core, sd, view, frames=3)
    core.Shutdown()


def test_no_lights_no_sky_is_black():
    sd = scenes.config2_scene(16, 12, n_materials=2, light_quads=1)
    sd.tri_lights = sd.tri_lights[:0]
    sd.meshes, sd.instances = sd.meshes[:1], sd.instances[:1]
    sd.sky = None
    view = scenes.view_pyramid((0, 30, -80), (0, 0, 0), 40, W, H)
    core = RenderCore()
    core.SetTarget(W, H, 1)
    sd.upload(core)
    core.Render(view, 1)
    st = core.GetCoreStats()
    assert not core.ReadPixels()[..., :3].any() and int(st["totalShadowRays"]) == 0 and int(st["primaryRayCount"]) == W * H
    core.Shutdown()


def test_present_gl_without_a_gl_context_fails_cleanly():
    """lh2b_present_gl (CUDA-GL interop, the reference's InteropTexture path) in a process without an OpenGL context: an error code and a
    message, no crash, and the core keeps rendering - the C++ class then falls back to the host upload."""
    from lighthouse2_b200 import capi
    lib = capi.load_library()
    sd = scenes.config2_scene(12, 8, n_materials=1, light_quads=1)
    core = RenderCore()
    core.SetTarget(64, 48, 1)
    sd.upload(core)
    view = scenes.view_pyramid((0, 30, -80), (0, 0, 0), 40, 64, 48)
    core.Render(view, 1)
    rc = lib.lh2b_present_gl(core._h, 7)
    assert rc != 0 and b"present_gl" in lib.lh2b_last_error()
    core.Render(view, 1)
    assert np.isfinite(core.ReadPixels()).all()
    core.Shutdown()
