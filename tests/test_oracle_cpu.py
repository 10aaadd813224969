"""CPU: known-answer tests of the oracle itself (closed-form cases), independent of any GPU result."""
import numpy as np
from lighthouse2_b200 import scenes
from oracle import binding as orc


def _tri(a, b, c):
    v = np.zeros((3, 4), np.float32)
    v[:, :3] = [a, b, c]
    return v


def test_closest_hit_known_answer():
    tri = _tri((0, 0, 0), (4, 0, 0), (0, 4, 0))
    O = np.array([[1, 1, -3, 0], [1, 1, 3, 0], [3, 3, -3, 0], [1, 1, -3, 0]], np.float32)
    D = np.array([[0, 0, 1, 0], [0, 0, -1, 0], [0, 0, 1, 0], [0, 0, -1, 0]], np.float32)
    h = orc.closest_hits([tri], [(0, None)], O, D)
    # hit from the front and from the back (no culling), u = v = 0.25, t = 3
    for i in (0, 1):
        assert h[i, 1] == 0 and h[i, 2] == 0 and h[i, 3:4].view(np.float32)[0] == 3.0
        assert (h[i, 0] & 65535) == int(65535 * 0.25) and (h[i, 0] >> 16) == int(65535 * 0.25)
    assert h[2, 2] == 0xFFFFFFFF and h[3, 2] == 0xFFFFFFFF       # outside (u + v > 1) and pointing away (t < 0)


def test_instance_transform_and_tie_break():
    tri = _tri((0, 0, 0), (4, 0, 0), (0, 4, 0))
    m_far, m_near = np.eye(4, dtype=np.float32), np.eye(4, dtype=np.float32)
    m_far[2, 3], m_near[2, 3] = 5.0, 2.0
    O = np.array([[1, 1, -1, 0]], np.float32); D = np.array([[0, 0, 1, 0]], np.float32)
    h = orc.closest_hits([tri], [(0, m_far), (0, m_near)], O, D)
    assert h[0, 1] == 1 and h[0, 3:4].view(np.float32)[0] == 3.0
    # two coincident instances: equal t, the smaller instance index wins
    h = orc.closest_hits([tri], [(0, m_near), (0, m_near)], O, D)
    assert h[0, 1] == 0
    # duplicated triangle inside one mesh: the smaller primitive index wins
    h = orc.closest_hits([np.concatenate([tri, tri])], [(0, None)], O + np.float32([0, 0, 0, 0]), D)
    assert h[0, 2] == 0


def test_occlusion_interval_is_open():
    tri = _tri((0, 0, 0), (4, 0, 0), (0, 4, 0))
    O = np.array([[1, 1, -3, 0]] * 3, np.float32)
    D = np.array([[0, 0, 1, 2.9], [0, 0, 1, 3.0], [0, 0, 1, 3.1]], np.float32)
    assert orc.occluded([tri], [(0, None)], O, D).tolist() == [0, 0, 1]


def test_white_furnace_floor_under_constant_sky():
    """A diffuse floor under a constant sky, no lights: every floor pixel is exactly albedo * sky for any sample,
    because cosine-weighted sampling cancels the BSDF (throughput = albedo) and the bounce always escapes."""
    sd = scenes.SceneDesc()
    sd.materials = scenes.make_materials([dict(color=(0.5, 0.25, 0.75))])
    q = scenes.quad((0, 0, 0), (0, 1, 0), 400, 400)
    sd.meshes = [(q, scenes.core_tris_from_verts(q))]
    sd.instances = [(0, None)]
    sky = np.full((64, 128, 3), (0.8, 0.6, 0.4), np.float32)
    sd.sky = (sky, 128, 64)
    W, H = 32, 16
    view = scenes.view_pyramid((0, 5, 0), (0, 0, 3), 40, W, H)
    o = orc.FrameOracle(sd, W, H, 2, 1e-3, 10.0, 3, 1, threads=2)
    img = o.render(view, 1)
    want = np.float32([0.5, 0.25, 0.75]) * np.float32([0.8, 0.6, 0.4])     # the camera looks down: every pixel sees the floor
    np.testing.assert_allclose(img[..., :3], np.broadcast_to(want, img[..., :3].shape), rtol=2e-4)
    assert o.ray_counts[0] == 2 * W * H * 2   # every path: primary + one bounce that escapes
    assert o.ray_counts[1] == 0                # no lights: no shadow rays


def test_skinning_restatement_known_answers():
    """host_mesh.cpp:884-904 on cases with closed-form answers: identity joints leave the mesh unchanged; one rigid joint
    moves positions by M and normals by its rotation; two joints at half weight blend the matrices (not the results)."""
    v = scenes.quad((0, 0, 0), (0, 1, 0), 2.0, 2.0)
    t = scenes.core_tris_from_verts(v)
    n = v.reshape(-1, 4).shape[0]
    ident = np.eye(4, dtype=np.float32)[None]
    j0, w0 = np.zeros((n, 4), np.uint32), np.tile(np.array([1, 0, 0, 0], np.float32), (n, 1))
    gv, gt = orc.skin_mesh(v, t, j0, w0, ident)
    assert np.allclose(gv[:, :3], v.reshape(-1, 4)[:, :3], atol=1e-7) and np.allclose(gt.view(np.float32), t.view(np.float32), atol=1e-6, equal_nan=True)
    a = 0.5
    R = np.eye(4, dtype=np.float32); R[0, 0], R[0, 1], R[1, 0], R[1, 1] = np.cos(a), -np.sin(a), np.sin(a), np.cos(a); R[:3, 3] = (1, 2, 3)
    gv, gt = orc.skin_mesh(v, t, j0, w0, R[None])
    pts = v.reshape(-1, 4).copy(); pts[:, 3] = 1
    want = (R @ pts.T).T
    assert np.allclose(gv[:, :3], want[:, :3], atol=1e-6)
    f = gt.view(np.float32).reshape(-1, 52)
    assert np.allclose(f[:, 8:11], np.tile(R[:3, :3] @ np.array([0, 1, 0], np.float32), (len(t), 1)), atol=1e-6)      # vN0
    assert np.allclose(np.stack([f[:, 11], f[:, 15], f[:, 19]], 1), np.tile(R[:3, :3] @ np.array([0, 1, 0], np.float32), (len(t), 1)), atol=1e-6)  # N
    T2 = np.eye(4, dtype=np.float32); T2[:3, 3] = (0, 4, 0)
    j1 = np.tile(np.array([0, 1, 0, 0], np.uint32), (n, 1)); w1 = np.tile(np.array([0.5, 0.5, 0, 0], np.float32), (n, 1))
    gv, _ = orc.skin_mesh(v, t, j1, w1, np.stack([np.eye(4, dtype=np.float32), T2]))
    assert np.allclose(gv[:, 1], v.reshape(-1, 4)[:, 1] + 2.0, atol=1e-6)
