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


def _frac_bad(a, b, tol):
    return float((np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)) > tol).any(axis=-1).mean())


def _uncombine(x):
    """5.11 fixed-point pairs (tools_shared.h:237-262) -> (direct, indirect) float triples"""
    u = np.ascontiguousarray(x).view(np.uint32)
    d = np.stack([(u[..., 0] >> 16), (u[..., 0] & 65535), u[..., 1]], -1) / 2048.0
    i = np.stack([(u[..., 2] >> 16), (u[..., 2] & 65535), u[..., 3]], -1) / 2048.0
    return d, i


def test_filter_chain_against_reference_kernel_vectors():
    """The oracle's CPU restatement of the SVGF / TAA chain (oracle/lh2_oracle_filter.h) against outputs of the REFERENCE's own
    filter kernels (finalize_shared.h, compiled unmodified for sm_100a and run on the B200 by tools/make_golden_filter.py).
    The reference kernels are fast-math builds, the oracle uses libm, and two stages are sensitive to that: the diamond search of
    specular pixels accepts a step on `d < bestDist`, which is a tie on planar regions (one search step of 5 * 0.45^k pixels
    more or less - 34 % of the specular pixels here, none of the diffuse ones), and the reference's TAA pass is racy (stored as a
    per-pixel [min, max] over 12 executions). Tolerances below are the measured disagreement with head-room."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "filter_reference_vectors.npz"))
    inputs = {k[3:]: g[k] for k in g.files if k.startswith("in_")}
    H, W = inputs["features"].shape[:2]

    def settings(taa, stationary):
        return dict(w=W, h=H, samplesTaken=1, camIsStationary=stationary, taa=taa, directClamp=15.0, indirectClamp=15.0, j0=0.0, j1=0.0,
                    prevj0=0.0, prevj1=0.0, prevView=g["prevView"])

    # B: stationary camera, no TAA - no search, no race: everything agrees closely
    b = orc.filter_chain_cpu(inputs, settings(0, 1))
    gd, gi = _uncombine(b["shadingAfterPrepare"]); wd, wi = _uncombine(g["B_shadingAfterPrepare"])
    assert _frac_bad(gd, wd, 1.5 / 2048) == 0 and _frac_bad(gi, wi, 1.5 / 2048) == 0
    assert _frac_bad(b["motion"], g["B_motion"], 1e-3) == 0 and _frac_bad(b["moments"], g["B_moments"], 2e-3) < 0.002
    assert (b["featuresOut"] != g["B_featuresOut"]).any(axis=-1).mean() < 0.002
    assert _frac_bad(b["phase3"][..., :3], g["B_phase3"][..., :3], 5e-3) < 0.005
    assert _frac_bad(b["target"][..., :3], g["B_target"][..., :3], 5e-3) < 0.005
    assert not b["target"][0].any() and not b["target"][:, 0].any()          # border pixels are never written (finalize_shared.h:594)
    # A: moving camera, TAA
    a = orc.filter_chain_cpu(inputs, settings(1, 0))
    gd, gi = _uncombine(a["shadingAfterPrepare"]); wd, wi = _uncombine(g["A_shadingAfterPrepare"])
    assert _frac_bad(gd, wd, 1.5 / 2048) == 0 and _frac_bad(gi, wi, 1.5 / 2048) == 0
    spec = ((inputs["features"][..., 3] >> 4) & 3) != 0
    d = np.abs(a["motion"] - g["A_motion"]).max(axis=-1)
    assert (d[~spec] > 1e-3).mean() == 0                                      # analytic reprojection of diffuse pixels
    assert 0.05 < spec.mean() < 0.5 and (d[spec] > 2e-2).mean() < 0.5 and (d[spec] > 1.1).mean() < 0.01   # searched pixels: within one early step
    assert _frac_bad(a["moments"], g["A_moments"], 2e-3) < 0.03
    assert (a["featuresOut"] != g["A_featuresOut"]).any(axis=-1).mean() < 0.01
    for k in ("phase1", "phase2"):
        gd, gi = _uncombine(a[k]); wd, wi = _uncombine(g["A_" + k])
        assert _frac_bad(gd, wd, 4.0 / 2048) < 0.02 and _frac_bad(gi, wi, 4.0 / 2048) < 0.03, k
    assert _frac_bad(a["phase3"][..., :3], g["A_phase3"][..., :3], 5e-3) < 0.015
    for k, tol, bound in (("taaPixels", 1e-2, 0.06), ("target", 3e-2, 0.05)):
        lo, hi, x = g["A_" + k + "_min"][..., :3], g["A_" + k + "_max"][..., :3], a[k][..., :3]
        assert float(((x < lo - tol) | (x > hi + tol)).any(axis=-1).mean()) < bound, k
    assert np.isfinite(a["target"]).all() and a["target"][1:-1, 1:-1, :3].mean() > 0.05
