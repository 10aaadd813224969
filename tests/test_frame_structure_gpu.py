"""Round-2 mechanics of the core that change how a frame is executed but must not change what it computes:
  - connect( L ) running next to extend( L + 1 ) on a second stream (Setting "overlapConnect") - frames stay bit-identical at 1 spp;
  - the SAH-optimal collapse to 8-wide (Setting "bvhCollapse") - another tree, the same hit records as the exhaustive search;
  - the traversal work counters (lh2b_trace_stats) - counting instantiations return the same hits and plausible counts;
  - the pipelined frame mode with both of the above."""
import numpy as np
import pytest

from lighthouse2_b200 import RenderCore, scenes
from oracle import binding as orc

pytestmark = pytest.mark.gpu
W, H = 160, 96


def _frames(settings, n=4, spp=1):
    sd = scenes.config2_scene(48, 32, n_materials=6, light_quads=2, floaters=300)
    core = RenderCore()
    core.SetTarget(W, H, spp)
    core.Setting("epsilon", 1e-3)
    core.Setting("maxPathLength", 5)
    core.Setting("maxDiffuseBounces", 2)
    for k, v in settings.items():
        core.Setting(k, v)
    sd.upload(core)
    out = []
    for f in range(n):
        view = scenes.view_pyramid((2 * f, 30, -80 + f), (0, 0, 0), 40, W, H)
        core.Render(view, 1 if f != 2 else 0)
        out.append(core.ReadPixels().copy())
    st = core.GetCoreStats()
    core.Shutdown()
    return out, (int(st["totalExtensionRays"]), int(st["totalShadowRays"]))


def test_connect_overlap_keeps_frames_bit_identical():
    a, ca = _frames({"overlapConnect": 0})
    b, cb = _frames({"overlapConnect": 1})
    assert ca == cb
    for x, y in zip(a, b):
        assert np.array_equal(x.view(np.uint32), y.view(np.uint32))


@pytest.mark.parametrize("builder", [0, 2], ids=["ploc", "lbvh"])
def test_optimal_and_greedy_collapse_trace_the_same_hits(builder):
    mesh = scenes.terrain(60, 40, extent=50, seed=7, floaters=300)
    O, D = scenes.random_rays(20000, extent=60, seed=5)
    want = orc.closest_hits([mesh], [(0, None)], O, D)
    nodes = []
    for collapse in (0, 1):
        core = RenderCore()
        core.Setting("bvhBuilder", builder)
        core.Setting("bvhCollapse", collapse)
        core.SetGeometry(0, mesh)
        core.SetInstance(0, 0)
        core.SetInstance(1, -1)
        core.FinalizeInstances()
        assert np.array_equal(core.TraceRays(O, D), want)
        nodes.append(int(core.GetBvhStats(0)["nodes"]))
        core.Shutdown()
    assert nodes[1] < 0.8 * nodes[0]          # fuller nodes: the optimal collapse needs far fewer of them


def test_trace_stats_count_the_work_and_change_nothing():
    mesh = scenes.terrain(60, 40, extent=50, seed=7, floaters=300)
    O, D = scenes.random_rays(8192, extent=60, seed=9)
    core = RenderCore()
    core.SetGeometry(0, mesh)
    core.SetInstance(0, 0)
    core.SetInstance(1, -1)
    core.FinalizeInstances()
    plain = core.TraceRays(O, D)
    core.TraceStatsEnable(True)
    counted = core.TraceRays(O, D)
    st = core.TraceStatsRead()
    core.TraceStatsEnable(False)
    assert np.array_equal(plain, counted)
    hits = int((plain[:, 2] != 0xFFFFFFFF).sum())
    assert st["rays"] == 8192 and st["nodeSteps"] >= 8192 and st["triTests"] >= hits > 0
    assert st["nodePhases"] <= st["iterations"] and st["nodeLanes"] == st["nodeSteps"] and st["triLanes"] == st["triTests"]
    assert st["instanceEntries"] == 0          # one identity instance: flat scene
    again = core.TraceStatsRead()
    assert again["rays"] == 0                  # reset by the first read; nothing counted while disabled
    core.TraceRays(O, D)
    assert core.TraceStatsRead()["rays"] == 0
    core.Shutdown()


def test_pipelined_frames_equal_synchronous_frames():
    sd = scenes.config2_scene(48, 32, n_materials=6, light_quads=2, floaters=300)
    views = [scenes.view_pyramid((2 * f, 30, -80 + f), (0, 0, 0), 40, W, H) for f in range(4)]

    def run(pipeline):
        core = RenderCore()
        core.SetTarget(W, H, 1)
        core.Setting("epsilon", 1e-3)
        sd.upload(core)
        core.Setting("pipeline", pipeline)
        imgs = []
        for v in views:
            core.Render(v, 1, bool(pipeline))
            if pipeline:
                core.WaitForRender()
            imgs.append(core.ReadPixels().copy())
        core.Shutdown()
        return imgs

    for x, y in zip(run(0), run(1)):
        assert np.array_equal(x.view(np.uint32), y.view(np.uint32))


def test_core_instances_of_one_process_agree_bit_for_bit():
    """Two cores of one process, each with its own allocations, must render the same frame bit for bit - also where the blue-noise sampler
    leaves its table (ranking tile indexed with the unwrapped dimension: tile pixel (127, 127) at path lengths >= 3 reads the 16 zero
    words of padding behind it; before the padding existed it read the neighbouring allocation and the instances differed in 1-4 pixels of a
    4K frame). A frame larger than one 128 x 128 tile, so that the pixel is in it, and three cores alive at once (different neighbours)."""
    Wb, Hb = 512, 288
    sd = scenes.config2_scene(100, 50, n_materials=16, light_quads=4)
    cores = []
    for _ in range(3):
        c = RenderCore()
        c.SetTarget(Wb, Hb, 1)
        c.Setting("epsilon", 1e-3), c.Setting("maxPathLength", 4)
        sd.upload(c)
        cores.append(c)
    for f in range(3):
        view = scenes.view_pyramid((0.5 * f, 30, -80 + 0.2 * f), (0, 0, 0), 40, Wb, Hb)
        acc = []
        for c in cores:
            c.Render(view, 1)
            acc.append(c.ReadAccumulator().copy())
        for other in acc[1:]:
            assert np.array_equal(acc[0].view(np.uint32), other.view(np.uint32)), f"frame {f}: core instances differ"
    for c in cores:
        c.Shutdown()
