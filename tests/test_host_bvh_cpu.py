"""CPU: the product's host BVH builder (Setting "bvhBuilder" 1: binned-SAH binary tree, greedy collapse to 8-wide, node encoding -
lighthouse2_b200/csrc/bvh_build_cpu.cpp) checked without a GPU. lh2b_host_bvh_build hands out the encoded nodes and triangle
records; the oracle's own reader of the format (oracle/lh2_oracle_cwbvh.h, written from the layout documented in csrc/bvh.h)
  - verifies the structure: every node / record reachable exactly once, records are the mesh's triangles in Moeller-Trumbore
    form, every encoded child box contains everything below it, slot masks well-formed;
  - traverses it with plain float slab tests and the oracle's triangle test: the hits must equal the exhaustive search bit for bit."""
import ctypes
import numpy as np
import pytest

from lighthouse2_b200 import capi, scenes
from oracle import binding as orc


def host_bvh(verts):
    lib = capi.load_library()
    v = np.ascontiguousarray(verts, np.float32).reshape(-1, 4)
    n = v.shape[0] // 3
    nodes, tris = np.zeros((2 * n + 8, 80), np.uint8), np.zeros((n + 8, 12), np.float32)
    counts = (ctypes.c_int * 2)()
    rc = lib.lh2b_host_bvh_build(ctypes.c_void_p(v.ctypes.data), n, ctypes.c_void_p(nodes.ctypes.data), nodes.shape[0],
                                 ctypes.c_void_p(tris.ctypes.data), tris.shape[0], counts)
    assert rc == 0, rc
    return nodes[:counts[0]].copy(), tris[:counts[1]].copy()


MESHES = {
    "terrain": lambda: scenes.terrain(60, 40, extent=50, seed=7, floaters=300),
    "soup": lambda: scenes.random_soup(6000, extent=10, size=1.5, seed=3),
    "quads-and-degenerates": lambda: np.concatenate([scenes.quad((0, 0, 0), (0, 1, 0), 20, 20), scenes.quad((3, 2, 0), (1, 0, 0), 6, 6),
                                                     np.zeros((3, 4), np.float32), np.float32([[-2, 1, -2, 0], [2, 1.000001, 2, 0], [0, 1.0000005, 0, 0]])]),
    "single-triangle": lambda: np.float32([[-1, 0, -1, 0], [1, 0, -1, 0], [0, 0, 1, 0]]),
}


@pytest.mark.parametrize("name", list(MESHES))
def test_host_builder_emits_a_correct_cwbvh(name):
    verts = MESHES[name]().reshape(-1, 4)
    n = verts.shape[0] // 3
    nodes, tris = host_bvh(verts)
    rep = orc.cwbvh_check(nodes, tris, verts)
    assert rep["errors"] == 0, rep
    assert rep["nodesVisited"] == nodes.shape[0] and rep["trisVisited"] == tris.shape[0] == n
    assert rep["leafSlots"] + rep["innerSlots"] + rep["emptySlots"] == 8 * nodes.shape[0]
    if n > 64:
        assert nodes.shape[0] < n and rep["maxDepth"] <= 2 + int(np.ceil(np.log2(n)))       # 8-wide: far fewer nodes than triangles
    # traversal of the encoded structure == exhaustive search
    ext = float(np.abs(verts[:, :3]).max()) + 4
    O, D = scenes.random_rays(6000, extent=ext, seed=11)
    want = orc.closest_hits([verts], [(0, None)], O, D)
    got = orc.cwbvh_closest_hits(nodes, tris, O, D)
    assert np.array_equal(got, want)
    if n > 64:
        assert (want[:, 2] != 0xFFFFFFFF).sum() > 200


def test_reader_detects_a_broken_structure():
    """The checker is not vacuous: shrinking one quantised box, dropping a triangle or duplicating a child is reported."""
    verts = scenes.terrain(20, 16, extent=20, seed=5).reshape(-1, 4)
    nodes, tris = host_bvh(verts)
    assert orc.cwbvh_check(nodes, tris, verts)["errors"] == 0
    masks = int(nodes.view(np.uint32).reshape(-1, 20)[0, 6])
    slot = int(np.nonzero([(masks >> s) & 1 or (masks >> (8 + s)) & 1 for s in range(8)])[0][0])
    bad = nodes.copy()
    bad[0, 56 + slot] = bad[0, 32 + slot]                 # qhi.x := qlo.x: the child's box collapses in x
    assert orc.cwbvh_check(bad, tris, verts)["errors"] > 0
    bad = nodes.copy()
    bad.view(np.uint32).reshape(-1, 20)[0, 6] &= ~np.uint32((1 << slot) | (1 << (8 + slot)))   # the slot is declared empty: what hangs below it is lost
    assert orc.cwbvh_check(bad, tris, verts)["errors"] > 0
    t2 = tris.copy()
    t2[5, 0] += 0.25                                       # a record that is no longer the mesh's triangle
    assert orc.cwbvh_check(nodes, t2, verts)["errors"] > 0
