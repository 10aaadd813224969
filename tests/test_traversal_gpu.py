"""GPU parity of extend / connect (hand-written CWBVH traversal) against the brute-force CPU oracle.
Bit-exact: hit record = (u16|v16<<16, instance, primitive, t bits); occlusion flag."""
import numpy as np
import pytest

from lighthouse2_b200 import RenderCore, scenes
from oracle import binding as orc

pytestmark = pytest.mark.gpu


def _rot(axis, angle, t=(0, 0, 0), s=1.0):
    axis = np.asarray(axis, np.float64) / np.linalg.norm(axis)
    x, y, z = axis
    c, si = np.cos(angle), np.sin(angle)
    r = np.array([[c + x * x * (1 - c), x * y * (1 - c) - z * si, x * z * (1 - c) + y * si],
                  [y * x * (1 - c) + z * si, c + y * y * (1 - c), y * z * (1 - c) - x * si],
                  [z * x * (1 - c) - y * si, z * y * (1 - c) + x * si, c + z * z * (1 - c)]])
    m = np.eye(4)
    m[:3, :3] = r * s
    m[:3, 3] = t
    return m.astype(np.float32)


BUILDER = {"gpu-ploc": 0, "host-sah": 1, "gpu-lbvh": 2}


@pytest.fixture(params=list(BUILDER), autouse=True)
def builder(request):
    """Every parity case runs with each builder: GPU PLOC (default), host binned SAH, GPU LBVH (radix tree)."""
    global _builder
    _builder = BUILDER[request.param]
    yield request.param


_builder = 0


def _check(meshes, instances, O, D, shadow_tmax=None):
    core = RenderCore()
    core.Setting("bvhBuilder", _builder)
    for i, m in enumerate(meshes):
        core.SetGeometry(i, m)
    for i, (mi, xf) in enumerate(instances):
        core.SetInstance(i, mi, xf)
    core.SetInstance(len(instances), -1)
    core.FinalizeInstances()
    got = core.TraceRays(O, D)
    want = orc.closest_hits(meshes, instances, O, D)
    miss_g, miss_w = got[:, 2] == 0xFFFFFFFF, want[:, 2] == 0xFFFFFFFF
    assert np.array_equal(miss_g, miss_w), f"{(miss_g != miss_w).sum()} rays disagree on hit/miss"
    h = ~miss_w
    bad = np.nonzero((got[h] != want[h]).any(axis=1))[0]
    assert bad.size == 0, f"{bad.size} of {h.sum()} hit records differ, first: got {got[h][bad[:3]]} want {want[h][bad[:3]]}"
    if shadow_tmax is not None:
        D2 = D.copy()
        D2[:, 3] = shadow_tmax
        og = core.TraceShadowRays(O, D2)
        ow = orc.occluded(meshes, instances, O, D2)
        assert np.array_equal(og, ow), f"{(og != ow).sum()} occlusion flags differ"
    core.Shutdown()
    return int(h.sum())


def test_single_mesh_soup():
    mesh = scenes.random_soup(20000, extent=10, size=1.5, seed=3)
    O, D = scenes.random_rays(20000, extent=14, seed=4)
    hits = _check([mesh], [(0, None)], O, D, shadow_tmax=9.0)
    assert hits > 2000


def test_terrain_primary_rays():
    mesh = scenes.terrain(60, 40, extent=50, seed=7, floaters=500)
    view = scenes.view_pyramid((0, 30, -80), (0, 0, 0), 40, 160, 90)
    O, D = scenes.camera_rays(view, 160, 90)
    hits = _check([mesh], [(0, None)], O, D, shadow_tmax=60.0)
    assert hits > 160 * 90 // 4


def test_instances_two_level():
    m0 = scenes.random_soup(3000, extent=3, size=0.8, seed=11)
    m1 = scenes.terrain(20, 20, extent=4, seed=12)
    inst = [(0, _rot((0, 1, 0), 0.3, (5, 0, 0))), (1, _rot((1, 0, 0), -0.7, (-4, 1, 2), 1.5)),
            (0, _rot((1, 1, 0), 1.1, (0, -3, -5), 0.5)), (1, None), (0, _rot((0, 0, 1), 2.0, (1, 6, 1)))]
    O, D = scenes.random_rays(20000, extent=12, seed=13)
    hits = _check([m0, m1], inst, O, D, shadow_tmax=7.5)
    assert hits > 2000


def test_axis_aligned_and_degenerate():
    # floor quads (zero-thickness boxes), a sliver, a zero-area triangle, axis-parallel rays
    q = [scenes.quad((0, 0, 0), (0, 1, 0), 20, 20), scenes.quad((0, 5, 0), (0, -1, 0), 4, 4),
         scenes.quad((3, 2, 0), (1, 0, 0), 6, 6)]
    deg = np.zeros((6, 4), np.float32)
    deg[0:3, :3] = [[1, 1, 1], [1, 1, 1], [1, 1, 1]]
    deg[3:6, :3] = [[-2, 1, -2], [2, 1.000001, 2], [0, 1.0000005, 0]]
    mesh = np.concatenate(q + [deg])
    rng = np.random.default_rng(5)
    n = 4096
    O = np.zeros((n, 4), np.float32); D = np.zeros((n, 4), np.float32)
    O[:, :3] = (rng.random((n, 3)) * 2 - 1) * [9, 0, 9] + [0, 8, 0]
    D[:, 1] = -1  # straight down, dx = dz = 0
    O2, D2 = scenes.random_rays(4096, extent=9, seed=6)
    O = np.concatenate([O, O2]); D = np.concatenate([D, D2])
    _check([mesh], [(0, None)], O, D, shadow_tmax=6.0)


def test_empty_and_tiny():
    tri = np.zeros((3, 4), np.float32)
    tri[:, :3] = [[-1, 0, -1], [1, 0, -1], [0, 0, 1]]
    O, D = scenes.random_rays(512, extent=3, seed=8)
    _check([tri], [(0, None)], O, D, shadow_tmax=2.0)
    _check([tri], [(0, _rot((1, 0, 0), 0.5, (0, 1, 0)))], O, D, shadow_tmax=2.0)


@pytest.mark.parametrize("mode", [1, 2], ids=["in-place", "recollapse"])
@pytest.mark.parametrize("xf", [False, True], ids=["flat", "two-level"])
def test_refit_after_deformation(mode, xf):
    """SetGeometry with the same triangle count: the GPU builder keeps the topology - bvhRefit 1 refits the binary tree and
    requantises the wide tree in place, 2 collapses the refitted binary tree again. Results must still be exact, in a flat
    scene (triangle records keep their instance tag) and behind an instance transform."""
    base = scenes.terrain(50, 40, extent=20, seed=21)
    O, D = scenes.random_rays(20000, extent=24, seed=22)
    inst = [(0, _rot((0, 1, 0), 0.4, (1, 0.5, -2)) if xf else None)]
    core = RenderCore()
    core.Setting("bvhBuilder", _builder)
    core.Setting("bvhRefit", mode)
    core.SetGeometry(0, base)
    core.SetInstance(0, 0, inst[0][1])
    core.SetInstance(1, -1)
    core.FinalizeInstances()
    rng = np.random.default_rng(23)
    for step in range(3):
        mesh = base.copy()
        mesh[:, 1] += (np.sin(mesh[:, 0] * 0.7 + step) * 1.5 + rng.standard_normal(len(mesh)) * 0.2 * step).astype(np.float32)
        core.SetGeometry(0, mesh)
        core.FinalizeInstances()
        got = core.TraceRays(O, D)
        want = orc.closest_hits([mesh], inst, O, D)
        assert np.array_equal(got, want), f"refit step {step}: {(got != want).any(axis=1).sum()} records differ"
    core.Shutdown()


def test_many_instances():
    m0 = scenes.random_soup(400, extent=1.0, size=0.5, seed=31)
    m1 = scenes.terrain(8, 8, extent=1.5, seed=32)
    rng = np.random.default_rng(33)
    inst = []
    for i in range(150):
        t = (rng.random(3) * 2 - 1) * 12
        inst.append((i % 2, _rot(rng.standard_normal(3), rng.random() * 6.28, t, 0.5 + rng.random())))
    O, D = scenes.random_rays(8000, extent=16, seed=34)
    hits = _check([m0, m1], inst, O, D, shadow_tmax=10.0)
    assert hits > 500


def test_full_size_c2_frame_rays():
    """BASELINE.json configs[1] at full size: the 1,000,002-triangle bench scene, all 2,073,600 primary rays of the 1920x1080
    frame and as many shadow-type queries, bit-exact against the CPU oracle (exhaustive search pruned by the oracle's own BVH,
    which tests/test_oracle_cpu.py proves identical to the unpruned search)."""
    import bench
    sd, view = bench.build_scene()
    meshes = [m[0] for m in sd.meshes]
    O, D = scenes.camera_rays(view, bench.W, bench.H)
    assert O.shape[0] == bench.W * bench.H
    with orc.accel(1):
        hits = _check(meshes, list(sd.instances), O, D, shadow_tmax=45.0)
    assert hits > bench.W * bench.H // 3


def test_full_size_instanced_two_level():
    """configs[3] shape: 1000 transformed instances of 10 meshes (100 k triangles each here), 400 k random rays, two-level
    traversal bit-exact against the BVH-pruned oracle."""
    rng = np.random.default_rng(41)
    meshes = [scenes.terrain(250, 200, extent=3.0, seed=50 + i) for i in range(10)]
    inst = []
    for i in range(1000):
        t = (rng.random(3) * 2 - 1) * [60, 10, 60]
        inst.append((i % 10, _rot(rng.standard_normal(3), rng.random() * 6.28, t, 0.5 + rng.random())))
    O, D = scenes.random_rays(400000, extent=70, seed=42)
    with orc.accel(1):
        hits = _check(meshes, inst, O, D, shadow_tmax=50.0)
    assert hits > 20000
