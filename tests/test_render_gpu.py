"""GPU radiance parity of the full wavefront loop (generate, extend, shade, connect, finalize) against the
CPU oracle frame renderer, same seeds, same spp. Tolerances: the device code uses fast-math intrinsics, the
oracle libm; a handful of paths flip a discrete decision, everything else agrees to ~1e-5."""
import numpy as np
import pytest

from lighthouse2_b200 import RenderCore, scenes
from oracle import binding as orc
from util import rel_rmse, pixel_mismatch_fraction

pytestmark = pytest.mark.gpu
W, H = 160, 90
REL_RMSE_TOL = 0.02          # relative RMSE bound on the frame (stated tolerance for radiance parity)
MISMATCH_TOL = 0.005         # at most 0.5 % of the pixels may differ by more than 1e-3 relative


def _scene(n_mat=6, lights=2):
    return scenes.config2_scene(48, 32, n_materials=n_mat, light_quads=lights, floaters=300)


def _core(sd, spp=1, maxlen=3, bounces=1, eps=1e-3):
    core = RenderCore()
    core.SetTarget(W, H, spp)
    core.Setting("epsilon", eps)
    core.Setting("clampValue", 10.0)
    core.Setting("maxPathLength", maxlen)
    core.Setting("maxDiffuseBounces", bounces)
    sd.upload(core)
    return core


def _compare(core, oracle, view, converge=1):
    core.Render(view, converge)
    got = core.ReadPixels()
    want = oracle.render(view, converge)
    st = core.GetCoreStats()
    assert np.isfinite(got).all()
    r, f = rel_rmse(got, want), pixel_mismatch_fraction(got, want)
    assert r < REL_RMSE_TOL and f < MISMATCH_TOL, (r, f)
    # ray counts: identical up to flipped decisions
    ext, shd = oracle.ray_counts
    assert abs(int(st["totalExtensionRays"]) - ext) <= max(8, ext // 2000), (st["totalExtensionRays"], ext)
    assert abs(int(st["totalShadowRays"]) - shd) <= max(8, shd // 2000), (st["totalShadowRays"], shd)
    return r, f


def test_frame_default_settings():
    sd = _scene()
    view = scenes.view_pyramid((0, 30, -80), (0, 0, 0), 40, W, H)
    core, oracle = _core(sd), orc.FrameOracle(sd, W, H, 1, 1e-3, 10.0, 3, 1)
    _compare(core, oracle, view)
    core.Shutdown()


def test_converge_sequence_and_restart():
    """Restart, Converge, Converge, Restart: samplesTaken, blue-noise shift and camRNGseed must evolve as in
    rendercore.cpp:827-833,855,900."""
    sd = _scene(3, 1)
    view = scenes.view_pyramid((10, 25, -70), (0, 2, 0), 40, W, H, aperture=0.05, distortion=0.05)
    core, oracle = _core(sd), orc.FrameOracle(sd, W, H, 1, 1e-3, 10.0, 3, 1)
    for conv in (1, 0, 0, 1, 0):
        _compare(core, oracle, view, conv)
    core.Shutdown()


def test_long_paths_multi_spp():
    sd = _scene(6, 3)
    view = scenes.view_pyramid((0, 12, -60), (0, 0, 0), 50, W, H)
    core, oracle = _core(sd, spp=4, maxlen=8, bounces=2), orc.FrameOracle(sd, W, H, 4, 1e-3, 10.0, 8, 2)
    _compare(core, oracle, view)
    core.Shutdown()


def test_no_lights_sky_only():
    sd = _scene(2, 1)
    sd.tri_lights = sd.tri_lights[:0]
    sd.meshes, sd.instances = sd.meshes[:1], sd.instances[:1]
    view = scenes.view_pyramid((0, 30, -80), (0, 0, 0), 40, W, H)
    core, oracle = _core(sd), orc.FrameOracle(sd, W, H, 1, 1e-3, 10.0, 3, 1)
    _compare(core, oracle, view)
    st = core.GetCoreStats()
    assert st["totalShadowRays"] == 0
    core.Shutdown()


def test_probe_and_stats():
    sd = _scene(1, 1)
    view = scenes.view_pyramid((0, 30, -80), (0, 0, 0), 40, W, H)
    core = _core(sd)
    core.SetProbePos(W // 2, H // 2)
    core.Render(view, 1)
    st = core.GetCoreStats()
    oracle = orc.FrameOracle(sd, W, H, 1, 1e-3, 10.0, 3, 1)
    _, rec = oracle.render(view, 1, records=True)
    want = rec[W // 2 + (H // 2) * W]["hit"]       # primary hit record of the probed pixel (same jittered ray)
    assert (int(st["probedInstid"]), int(st["probedTriid"])) == (int(want[1]), int(want[2]))
    assert abs(float(st["probedDist"]) / float(want[3:4].view(np.float32)[0]) - 1) < 1e-4
    assert st["primaryRayCount"] == W * H and st["totalRays"] == st["totalExtensionRays"] + st["totalShadowRays"]
    core.Shutdown()


def test_pipelined_readback_matches_blocking():
    """lh2b_read_pixels_async: the copy of frame k overlaps frame k+1 (which presents into the second pixel buffer); every
    frame read that way must equal the blocking read of an identically seeded core. Six frames cover both buffers, the
    device-side wait on a still-pending copy, and asynchronous Render (async=True) in front of the read."""
    import torch
    sd = _scene(3, 1)
    views = [scenes.view_pyramid((10 + 3 * k, 25, -70), (0, 2, 0), 40, W, H) for k in range(6)]
    a, b = _core(sd), _core(sd)
    want = []
    for v in views:
        a.Render(v, 1)
        want.append(a.ReadPixels().copy())
    pinned = [torch.empty((H, W, 4), dtype=torch.float32, pin_memory=True) for _ in views]
    for k, v in enumerate(views):
        b.Render(v, 1, k % 2 == 1)               # odd frames: asynchronous render, read enqueued while the frame is in flight
        b.ReadPixelsAsync(pinned[k].numpy())
    b.WaitReadPixels()
    for k in range(len(views)):
        assert np.array_equal(pinned[k].numpy(), want[k]), k
    # the blocking read still sees the last frame
    assert np.array_equal(b.ReadPixels(), want[-1])
    a.Shutdown(), b.Shutdown()


def test_frame_principled_materials():
    """Setting("bsdf", 1): the principled model the stock reference cores compile (disney.h), 16 materials covering every lobe,
    longer paths and two diffuse bounces so sampled lobes feed further vertices."""
    sd = scenes.config2_scene(48, 32, light_quads=2, floaters=300, material_specs=scenes.principled_specs(16))
    view = scenes.view_pyramid((0, 30, -80), (0, 0, 0), 40, W, H)
    core = _core(sd, spp=2, maxlen=5, bounces=2)
    core.Setting("bsdf", 1)
    oracle = orc.FrameOracle(sd, W, H, 2, 1e-3, 10.0, 5, 2, bsdf=1)
    for conv in (1, 0):
        _compare(core, oracle, view, conv)
    core.Shutdown()


def test_pipelined_frames_match_blocking_frames():
    """Setting("pipeline", 1) + asynchronous Render: frame k+1 is enqueued while frame k is in flight, frame k is harvested
    afterwards. Pixels (read with the pipelined read-back) and ray counts must equal the blocking sequence, the statistics
    lag one frame until WaitForRender."""
    import torch
    sd = _scene(3, 1)
    views = [scenes.view_pyramid((10 + 3 * k, 25, -70), (0, 2, 0), 40, W, H) for k in range(5)]
    conv = [1, 0, 0, 1, 0]
    a, b = _core(sd), _core(sd)
    want_px, want_rays = [], []
    for v, c in zip(views, conv):
        a.Render(v, c)
        want_px.append(a.ReadPixels().copy())
        st = a.GetCoreStats()
        want_rays.append((int(st["totalExtensionRays"]), int(st["totalShadowRays"])))
    b.Setting("pipeline", 1)
    pinned = [torch.empty((H, W, 4), dtype=torch.float32, pin_memory=True) for _ in views]
    got_rays = []
    for k, (v, c) in enumerate(zip(views, conv)):
        b.Render(v, c, True)
        b.ReadPixelsAsync(pinned[k].numpy())
        if k > 0:
            st = b.GetCoreStats()              # describes frame k-1
            got_rays.append((int(st["totalExtensionRays"]), int(st["totalShadowRays"])))
    b.WaitForRender()
    st = b.GetCoreStats()
    got_rays.append((int(st["totalExtensionRays"]), int(st["totalShadowRays"])))
    b.WaitReadPixels()
    assert got_rays == want_rays
    for k in range(len(views)):
        assert np.array_equal(pinned[k].numpy(), want_px[k]), k
    a.Shutdown(), b.Shutdown()


def _compare_full(sd, w, h, spp, maxlen, bounces, view):
    core = RenderCore()
    core.SetTarget(w, h, spp)
    core.Setting("epsilon", 1e-3)
    core.Setting("clampValue", 10.0)
    core.Setting("maxPathLength", maxlen)
    core.Setting("maxDiffuseBounces", bounces)
    sd.upload(core)
    with orc.accel(1):                       # BVH-pruned oracle: identical to its exhaustive search (tests/test_oracle_cpu.py)
        oracle = orc.FrameOracle(sd, w, h, spp, 1e-3, 10.0, maxlen, bounces)
        core.Render(view, 1)
        got, want = core.ReadPixels(), oracle.render(view, 1)
    st = core.GetCoreStats()
    r, f = rel_rmse(got, want), pixel_mismatch_fraction(got, want)
    ext, shd = oracle.ray_counts
    core.Shutdown()
    assert np.isfinite(got).all() and r < REL_RMSE_TOL and f < MISMATCH_TOL, (r, f)
    assert abs(int(st["totalExtensionRays"]) - ext) <= max(8, ext // 2000), (st["totalExtensionRays"], ext)
    assert abs(int(st["totalShadowRays"]) - shd) <= max(8, shd // 2000), (st["totalShadowRays"], shd)
    return r, f


def test_full_size_c2_frame():
    """BASELINE.json configs[1] at full size - the bench workload itself (1,000,002 triangles, 1920x1080, path length 1):
    every pixel of the frame against the CPU oracle frame, same seeds."""
    import bench
    sd, view = bench.build_scene()
    _compare_full(sd, bench.W, bench.H, 1, 1, 1, view)


def test_full_size_c3_frame():
    """configs[2] shape at full resolution: 1 M triangles, 64 materials, 8 emissive quads, 1920x1080, path length 8 with two
    diffuse bounces, NEE + MIS; 2 spp per Render here (16 in the config) so that the CPU oracle finishes in seconds."""
    sd = scenes.config2_scene(1000, 500, n_materials=64, light_quads=8)
    view = scenes.view_pyramid((0, 30, -80), (0, 0, 0), 40, 1920, 1080)
    _compare_full(sd, 1920, 1080, 2, 8, 2, view)


def test_more_than_64_lights_are_picked_uniformly():
    """80 emissive triangles: beyond MAXISLIGHTS the reference's importance sampling overruns its table; the core (and the oracle) take the
    reference's uniform-pick branch instead of rejecting the scene. Frame parity against the oracle as usual."""
    sd = scenes.config2_scene(24, 16, n_materials=3, light_quads=40, floaters=60)
    assert len(sd.tri_lights) == 80
    Wl, Hl = 96, 54
    view = scenes.view_pyramid((0, 30, -80), (0, 0, 0), 40, Wl, Hl)
    core = RenderCore()
    core.SetTarget(Wl, Hl, 1)
    core.Setting("epsilon", 1e-3)
    sd.upload(core)
    core.Render(view, 1)
    img = core.ReadPixels()
    st = core.GetCoreStats()
    want = orc.FrameOracle(sd, Wl, Hl, 1, 1e-3, 10.0, 3, 1).render(view, 1)
    assert int(st["totalShadowRays"]) > 0
    assert np.isfinite(img).all() and rel_rmse(img, want) < 0.02 and pixel_mismatch_fraction(img, want) < 0.005
    core.Shutdown()


def test_precise_math_frames_agree_tightly():
    """Setting "preciseMath" 1 (IEEE division / sqrt, accurate transcendentals, no FMA contraction in the shade stage): what separates the
    shipping fast-math frame from the libm oracle is then gone except for isolated discrete flips - at most 0.1 % of the pixels differ by
    more than 1e-3 relative (fast-math bound: 0.5 %) and the relative RMSE stays below 1 % over a Restart / Converge sequence with long
    paths and several samples."""
    sd = _scene(6, 3)
    view = scenes.view_pyramid((0, 12, -60), (0, 0, 0), 50, W, H)
    core, oracle = _core(sd, spp=2, maxlen=8, bounces=2), orc.FrameOracle(sd, W, H, 2, 1e-3, 10.0, 8, 2)
    core.Setting("preciseMath", 1)
    worst = (0.0, 0.0)
    for conv in (1, 0, 1):
        core.Render(view, conv)
        got, want = core.ReadPixels(), oracle.render(view, conv)
        worst = max(worst, (pixel_mismatch_fraction(got, want), rel_rmse(got, want)))
    assert worst[0] < 0.001 and worst[1] < 0.01, worst
    core.Shutdown()
