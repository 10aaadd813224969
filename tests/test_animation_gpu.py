"""Device-side mesh animation against the CPU restatement of HostMesh::SetPose
(lib/RenderSystem/host_mesh.cpp:711-741,748-906; oracle/lh2_oracle_anim.h), and the path behind it: after a pose the BVH is
refitted on the device and ray queries must be bit-exact against brute force over the posed triangles."""
import numpy as np
import pytest

from lighthouse2_b200 import RenderCore, scenes
from oracle import binding as orc

pytestmark = pytest.mark.gpu


def _rot(axis, a):
    c, s = np.cos(a), np.sin(a)
    m = np.eye(4, dtype=np.float32)
    i, j = [(1, 2), (0, 2), (0, 1)][axis]
    m[i, i], m[i, j], m[j, i], m[j, j] = c, -s, s, c
    return m


def _mesh(seed=3):
    v = scenes.terrain(24, 16, 10.0, seed, 40)
    t = scenes.core_tris_from_verts(v)
    # smooth-ish vertex normals so that normal skinning is exercised with non-face normals
    rng = np.random.default_rng(seed)
    tf = t.view(np.float32).reshape(-1, 52)
    for k in range(3):
        n = tf[:, (2 + k) * 4:(2 + k) * 4 + 3] + 0.2 * rng.standard_normal((len(t), 3)).astype(np.float32)
        tf[:, (2 + k) * 4:(2 + k) * 4 + 3] = n / np.linalg.norm(n, axis=1, keepdims=True)
    return v, t


def _skin(v, joints=5, seed=5):
    rng = np.random.default_rng(seed)
    n = v.reshape(-1, 4).shape[0]
    j = rng.integers(0, joints, (n, 4)).astype(np.uint32)
    w = rng.random((n, 4)).astype(np.float32)
    w[:, 2:] *= (rng.random((n, 2)) > 0.5)            # many vertices with two influences only
    w /= w.sum(axis=1, keepdims=True)
    return j, w


def _pose(joints, k):
    mats = []
    for q in range(joints):
        m = _rot(q % 3, 0.15 * (k + 1) * (q + 1) / joints) @ _rot((q + 1) % 3, -0.1 * k)
        m[:3, 3] = (0.3 * q * (k + 1), 0.2 * k, -0.1 * q)
        mats.append(m)
    return np.stack(mats).astype(np.float32)


def _core(v, t, builder):
    core = RenderCore()
    core.Setting("bvhBuilder", builder)
    core.SetGeometry(0, v, t)
    core.SetInstance(0, 0)
    core.SetInstance(1, -1)
    core.FinalizeInstances()
    return core


def _check_queries(core, verts, seed):
    O, D = scenes.random_rays(4000, 14.0, seed)
    hits = core.TraceRays(O, D)
    want = orc.closest_hits([verts], [(0, None)], O, D)
    assert np.array_equal(np.ascontiguousarray(hits).view(np.uint32), np.ascontiguousarray(want).view(np.uint32))


@pytest.mark.parametrize("builder", [0, 2, 1], ids=["ploc", "lbvh", "host-sah"])
def test_skinning_matches_host_restatement_and_refit_stays_exact(builder):
    v, t = _mesh()
    j, w = _skin(v)
    core = _core(v, t, builder)
    core.SetSkin(0, j, w)
    for k in range(3):
        mats = _pose(5, k)
        core.SetPose(0, mats)
        core.FinalizeInstances()                       # refit (same triangle count)
        gv, gt = core.ReadGeometry(0, len(t))
        wv, wt = orc.skin_mesh(v, t, j, w, mats)
        assert np.allclose(gv, wv, rtol=1e-5, atol=1e-5) and (gv[:, 3] == 1).all()
        gf, wf = gt.view(np.float32).reshape(-1, 52), wt.view(np.float32).reshape(-1, 52)
        assert np.allclose(gf[:, 8:20], wf[:, 8:20], atol=2e-5)          # vN0..2 + Nx/Ny/Nz
        assert np.allclose(gf[:, 32:44].reshape(-1, 4)[:, :3], wf[:, 32:44].reshape(-1, 4)[:, :3], rtol=1e-5, atol=1e-5)   # vertex0..2
        untouched = np.r_[0:8, 20:32, 44:52]
        assert np.array_equal(gf[:, untouched].view(np.uint32), t.view(np.float32).reshape(-1, 52)[:, untouched].view(np.uint32))
        _check_queries(core, gv, 100 + k)              # traversal over the refitted BVH vs brute force over the posed vertices
    core.Shutdown()


def test_morph_targets():
    v, t = _mesh(9)
    rng = np.random.default_rng(1)
    n = v.reshape(-1, 4).shape[0]
    deltas = np.zeros((2, n, 4), np.float32); deltas[..., :3] = 0.4 * rng.standard_normal((2, n, 3))
    normals = np.zeros((2, n, 4), np.float32); normals[..., :3] = 0.3 * rng.standard_normal((2, n, 3))
    core = _core(v, t, 0)
    core.SetMorphTargets(0, deltas, normals)
    for wts in ((0.0, 0.0), (0.7, 0.1), (0.2, 1.0)):
        core.SetMorphWeights(0, wts)
        core.FinalizeInstances()
        gv, gt = core.ReadGeometry(0, len(t))
        wv, wt = orc.morph_mesh(v, t, deltas, normals, np.array(wts, np.float32))
        assert np.allclose(gv, wv, rtol=1e-5, atol=1e-5)
        gf, wf = gt.view(np.float32).reshape(-1, 52), wt.view(np.float32).reshape(-1, 52)
        assert np.allclose(gf, wf, rtol=1e-5, atol=2e-5, equal_nan=True)
        _check_queries(core, gv, 7)
    core.Shutdown()


def test_set_geometry_from_device_pointers():
    import torch
    v, t = _mesh(4)
    core = _core(v, t, 0)
    v2 = v.copy().reshape(-1, 4); v2[:, 1] += 0.5 * np.sin(v2[:, 0])
    dv = torch.from_numpy(v2).cuda()
    dt = torch.from_numpy(t.view(np.uint8).reshape(len(t), -1).copy()).cuda()
    core.SetGeometryDevice(0, dv.data_ptr(), len(t), dt.data_ptr())
    core.FinalizeInstances()
    gv, gt = core.ReadGeometry(0, len(t))
    assert np.array_equal(gv, v2) and np.array_equal(gt.view(np.uint8), t.view(np.uint8))
    _check_queries(core, v2, 11)
    core.Shutdown()
