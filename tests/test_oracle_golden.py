"""CPU: pins the oracle's shading restatement (oracle/lh2_oracle_render.h ShadeStep) to the REFERENCE itself.
tests/golden/shade_reference_vectors.npz (lambert.h) and shade_disney_reference_vectors.npz (disney.h, the stock build's
BSDF, 16 principled materials) hold inputs and outputs of the reference's unmodified shadeKernel
(lib/rendercore_optix7/kernels/pathtracer.h:54-238), compiled for sm_100a and run on a B200 by
tools/make_golden_shade.py through oracle/_ref/libref_shade_gpu.so. The oracle must reproduce them path by path:
same paths emit extension / shadow rays, values within 2e-3 relative (the reference build is -use_fast_math, the
oracle uses libm), at most 0.3 % of paths may take a different discrete branch."""
import os
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from make_golden_shade import golden_scene, W, H  # noqa: E402
from oracle import binding as orc  # noqa: E402

GOLDEN = {0: os.path.join(ROOT, "tests", "golden", "shade_reference_vectors.npz"),            # reference kernel + lambert.h
          1: os.path.join(ROOT, "tests", "golden", "shade_disney_reference_vectors.npz")}     # reference kernel + disney.h (stock build)
FLIP_TOL, VAL_TOL = 0.003, 2e-3
PEAK_PDF_TOL = 5e-2     # sampling densities above 100 (near-singular microfacet lobes): see tests/test_shade_stage_gpu.py


@pytest.fixture(scope="module", params=[0, 1], ids=["lambert", "disney"])
def setup(request):
    bsdf = request.param
    g = np.load(GOLDEN[bsdf])
    sd, view = golden_scene(bsdf)
    chk = float(np.sum(sd.meshes[0][0][:, :3].astype(np.float64)))
    assert abs(chk - float(g["scene_checksum"][0])) < 1e-6 * max(1.0, abs(chk)), "scene generator changed: regenerate the golden vectors"
    return g, sd, view, orc.FrameOracle(sd, W, H, 1, 1e-3, 10.0, 3, 1, bsdf=bsdf)


def _match(a, b, ka, kb, n_total, what):
    ia, ib = np.argsort(ka, kind="stable"), np.argsort(kb, kind="stable")
    common, pa, pb = np.intersect1d(ka[ia], kb[ib], return_indices=True)
    flips = (len(ka) - len(common)) + (len(kb) - len(common))
    assert flips <= max(3, FLIP_TOL * n_total), f"{what}: {flips} paths emitted by only one side ({len(ka)} vs {len(kb)})"
    for name in a:
        xa, xb = a[name][ia][pa][:, :3].astype(np.float64), b[name][ib][pb][:, :3].astype(np.float64)
        if name == "T":   # peaked lobes: compare throughput / density (what the next vertex uses); the density is checked by the caller
            wa, wb = a[name][ia][pa][:, 3].astype(np.float64), b[name][ib][pb][:, 3].astype(np.float64)
            peaked = np.maximum(wa, wb) > 100.0
            xa, xb = np.where(peaked[:, None], xa / wa[:, None], xa), np.where(peaked[:, None], xb / wb[:, None], xb)
        bad = (np.abs(xa - xb) / (1e-3 + np.abs(xb)) > VAL_TOL).any(axis=1)
        assert bad.sum() <= max(3, FLIP_TOL * n_total), f"{what}.{name}: {bad.sum()} of {len(bad)} matched rays differ"
    return ia[pa], ib[pb]


@pytest.mark.parametrize("L", [1, 2, 3])
def test_oracle_shade_step_matches_reference_kernel(setup, L):
    g, sd, view, oracle = setup
    O4, D4, T4, hits = g[f"L{L}_O"], g[f"L{L}_D"], g[f"L{L}_T"], g[f"L{L}_hits"]
    n = len(O4)
    out = oracle.shade_paths(view, L, O4, D4, T4, hits, int(g[f"L{L}_R0"][0]), int(g["shift"][0]), 0)
    fl = out["flags"]
    # extension rays, matched by the path index in O.w
    mine = {"O": out["extO"][(fl & 1) > 0], "D": out["extD"][(fl & 1) > 0], "T": out["extT"][(fl & 1) > 0]}
    ref = {"O": g[f"L{L}_extO"], "D": g[f"L{L}_extD"], "T": g[f"L{L}_extT"]}
    ia, ib = _match(mine, ref, mine["O"][:, 3].view(np.uint32) >> 6, ref["O"][:, 3].view(np.uint32) >> 6, n, f"L{L} extension")
    if len(ia):
        # packed words: path flags (O.w), packed normal (D.w) and the postponed pdf (T.w)
        same = mine["O"][ia][:, 3].view(np.uint32) == ref["O"][ib][:, 3].view(np.uint32)
        assert same.mean() > 1 - FLIP_TOL
        pn_m, pn_r = mine["D"][ia][:, 3].view(np.uint32), ref["D"][ib][:, 3].view(np.uint32)
        close = (np.abs((pn_m & 65535).astype(np.int64) - (pn_r & 65535)) <= 2) & (np.abs((pn_m >> 16).astype(np.int64) - (pn_r >> 16)) <= 2)
        assert close.mean() > 1 - FLIP_TOL
        pdf = np.abs(mine["T"][ia][:, 3] - ref["T"][ib][:, 3]) / (1e-3 + np.abs(ref["T"][ib][:, 3]))
        assert (pdf > np.where(ref["T"][ib][:, 3] > 100.0, PEAK_PDF_TOL, VAL_TOL)).mean() <= FLIP_TOL
    # shadow rays, matched by pixel index in E.w
    mine = {"O": out["shO"][(fl & 2) > 0], "D": out["shD"][(fl & 2) > 0], "E": out["shE"][(fl & 2) > 0]}
    ref = {"O": g[f"L{L}_shO"], "D": g[f"L{L}_shD"], "E": g[f"L{L}_shE"]}
    ia, ib = _match(mine, ref, mine["E"][:, 3].view(np.uint32), ref["E"][:, 3].view(np.uint32), n, f"L{L} shadow")
    if len(ia):
        tmax = np.abs(mine["D"][ia][:, 3] - ref["D"][ib][:, 3]) / (1e-3 + np.abs(ref["D"][ib][:, 3]))
        assert (tmax > VAL_TOL).mean() <= FLIP_TOL
    # direct deposits (sky, emissive surfaces)
    dep = np.zeros((H * W, 3), np.float64)
    d = out["deposit"][(fl & 4) > 0]
    np.add.at(dep, d[:, 3].view(np.uint32), d[:, :3].astype(np.float64))
    acc = g[f"L{L}_acc"].reshape(-1, 4)[:, :3].astype(np.float64)
    bad = (np.abs(dep - acc) / (1e-3 + np.abs(acc)) > VAL_TOL).any(axis=1)
    assert bad.sum() <= max(3, FLIP_TOL * n), f"L{L} accumulator: {bad.sum()} pixels differ"


def test_golden_covers_the_branches(setup):
    g = setup[0]
    assert len(g["L1_extO"]) > 1000 and len(g["L1_shO"]) > 1000 and len(g["L2_extO"]) > 20 and len(g["L3_shO"]) > 3
    flags = g["L1_extO"][:, 3].view(np.uint32) & 63
    assert (flags & 1).any() and (flags & 2).any() and (flags & 4).any()      # specular, bounced and via-specular paths all occur
