"""Stage-level parity of the SVGF / TAA chain (prepare, a-trous x3, TAA, unsharp) against the REFERENCE's own kernels
compiled for sm_100a (oracle/_ref/libref_filter_gpu.so, from lib/CUDA/shared_kernel_code/finalize_shared.h), on identical
feature / history buffers derived from two camera positions over a real scene. Both sides are fast-math builds; packed
5.11 fixed-point shading may differ by an LSB where an operation was reassociated, hence small absolute tolerances.
The reference's TAA pass updates `pixels` in place while neighbouring threads still read it (a race); ours reads a tile
of the input and writes a separate buffer, so a small fraction of pixels may legitimately differ there."""
import numpy as np
import pytest

from lighthouse2_b200 import RenderCore, scenes
from oracle import binding as orc

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not orc.have_ref_filter_gpu(), reason="oracle/_ref/libref_filter_gpu.so is built only where /root/reference exists")]
W, H = 160, 96


def pack_normal2(n):
    q = np.clip(((n + 1) * 511).astype(np.int64), 0, 1023).astype(np.uint32)
    return (q[..., 0] << 2) + (q[..., 1] << 12) + (q[..., 2] << 22)


def hdr_to_rgb32(c):
    r = (1023.0 * np.minimum(1.0, c[..., 0])).astype(np.uint32)
    g = (2047.0 * np.minimum(1.0, c[..., 1])).astype(np.uint32)
    b = (2047.0 * np.minimum(1.0, c[..., 2])).astype(np.uint32)
    return (r << 22) + (g << 11) + b


def combine(a, b):
    q = lambda x: (np.minimum(x, 31.999) * 2048.0).astype(np.uint32)
    out = np.zeros(a.shape[:-1] + (4,), np.uint32)
    out[..., 0] = (q(a[..., 0]) << 16) + q(a[..., 1]); out[..., 1] = q(a[..., 2])
    out[..., 2] = (q(b[..., 0]) << 16) + q(b[..., 1]); out[..., 3] = q(b[..., 2])
    return out.view(np.float32)


def uncombine(x):
    u = np.ascontiguousarray(x).view(np.uint32).astype(np.float64)
    ui = np.ascontiguousarray(x).view(np.uint32)
    d = np.stack([(ui[..., 0] >> 16), (ui[..., 0] & 65535), ui[..., 1]], -1) / 2048.0
    i = np.stack([(ui[..., 2] >> 16), (ui[..., 2] & 65535), ui[..., 3]], -1) / 2048.0
    return d, i


def gbuffer(core, sd, view, rng, spec_mat):
    O, D = scenes.camera_rays(view, W, H)
    hits = core.TraceRays(O, D)
    t = hits[:, 3].view(np.float32).copy()
    miss = hits[:, 2] == 0xFFFFFFFF
    t[miss] = 1e34
    inst, prim = hits[:, 1].astype(np.int64), hits[:, 2].astype(np.int64)
    inst[miss], prim[miss] = 0, 0
    N = np.zeros((W * H, 3), np.float32); mat = np.zeros(W * H, np.int64)
    for i, (mi, _) in enumerate(sd.instances):
        tri = sd.meshes[mi][1]
        sel = (inst == i) & ~miss
        N[sel] = np.stack([tri["Nx"][prim[sel]], tri["Ny"][prim[sel]], tri["Nz"][prim[sel]]], -1)
        mat[sel] = tri["material"][prim[sel]]
    flip = (N * D[:, :3]).sum(-1) > 0
    N[flip] *= -1
    N[miss] = -D[miss, :3]
    P = O[:, :3] + D[:, :3] * np.minimum(t, 50000)[:, None]
    albedo = np.clip(sd.materials["color"]["value"][mat], 0.05, 1.0).astype(np.float32)
    albedo[miss] = 0.6
    spec = ((mat == spec_mat) & ~miss).astype(np.uint32)
    pn = pack_normal2(N) + spec
    feat = np.zeros((W * H, 4), np.uint32)
    feat[:, 0], feat[:, 1], feat[:, 2] = hdr_to_rgb32(albedo), pn, t.view(np.uint32)
    feat[:, 3] = (spec << 4) + (mat.astype(np.uint32) << 6) + rng.integers(0, 16, W * H).astype(np.uint32)
    wp = np.zeros((W * H, 4), np.float32); wp[:, :3] = P; wp[:, 3] = pn.view(np.float32)
    depth = t.reshape(H, W)
    dd = np.zeros((H, W, 4), np.float32)
    dd[:, :-1, 2] = np.clip(depth[:, 1:] - depth[:, :-1], -1, 1); dd[:-1, :, 3] = np.clip(depth[1:, :] - depth[:-1, :], -1, 1)
    return feat.reshape(H, W, 4), wp.reshape(H, W, 4), dd, albedo.reshape(H, W, 3)


@pytest.fixture(scope="module")
def case():
    rng = np.random.default_rng(77)
    sd = scenes.config2_scene(48, 32, n_materials=6, light_quads=2, floaters=300)
    core = RenderCore()
    for i, (v, t) in enumerate(sd.meshes):
        core.SetGeometry(i, v, t)
    for i, (m, xf) in enumerate(sd.instances):
        core.SetInstance(i, m, xf)
    core.SetInstance(len(sd.instances), -1)
    core.FinalizeInstances()
    prev_view = scenes.view_pyramid((0.0, 30, -80), (0, 0, 0), 40, W, H)
    view = scenes.view_pyramid((0.6, 30.2, -79.5), (0.1, 0, 0), 40, W, H)
    feat, wp, dd, albedo = gbuffer(core, sd, view, rng, spec_mat=1)
    _, pwp, _, _ = gbuffer(core, sd, prev_view, rng, spec_mat=1)
    smooth = lambda c: (0.4 + 0.3 * np.sin(np.linspace(0, 9, W))[None, :, None] * np.cos(np.linspace(0, 7, H))[:, None, None] + 0 * c).astype(np.float32)
    direct = albedo * (smooth(albedo) + 0.5 * rng.random((H, W, 1)).astype(np.float32))
    indirect = albedo * (0.3 * rng.random((H, W, 3)).astype(np.float32))
    acc = np.zeros((2, H, W, 4), np.float32); acc[0, ..., :3] = direct; acc[1, ..., :3] = indirect
    pm = np.zeros((H, W, 4), np.float32)
    pm[..., 0] = 0.5 + 0.1 * rng.random((H, W)); pm[..., 1] = pm[..., 0] ** 2 + 0.02 * rng.random((H, W))
    pm[..., 2] = 0.2 + 0.1 * rng.random((H, W)); pm[..., 3] = pm[..., 2] ** 2 + 0.01 * rng.random((H, W))
    fin = combine(smooth(albedo) + 0.05 * rng.random((H, W, 3)).astype(np.float32), 0.15 + 0.05 * rng.random((H, W, 3)).astype(np.float32))
    pp = np.zeros((H, W, 4), np.float32); pp[..., :3] = np.sqrt(albedo * 0.6) + 0.02 * rng.random((H, W, 3)).astype(np.float32)
    inputs = dict(accumulator=acc, features=feat, worldPos=wp, prevWorldPos=pwp, deltaDepth=dd, prevMoments=pm, filteredIN=fin, prevPixels=pp)
    yield core, inputs, prev_view
    core.Shutdown()


def _run_both(core, inputs, prev_view, taa, stationary):
    st = dict(w=W, h=H, samplesTaken=1, camIsStationary=stationary, taa=taa, directClamp=15.0, indirectClamp=15.0, j0=0.0, j1=0.0, prevj0=0.0, prevj1=0.0,
              prevView=prev_view)
    want = orc.ref_filter_gpu(inputs, st)
    io, got, keep = orc.make_filter_io(inputs, st)
    core.FilterChain(io)
    return got, want


def _frac_bad(a, b, tol):
    return float((np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)) > tol).any(axis=-1).mean())


@pytest.mark.parametrize("taa,stationary", [(1, 0), (0, 1)])
def test_filter_chain_matches_reference_kernels(case, taa, stationary):
    core, inputs, prev_view = case
    got, want = _run_both(core, inputs, prev_view, taa, stationary)
    # prepare: demodulated shading (5.11 fixed point), motion vectors, moments, history counters
    gd, gi = uncombine(got["shadingAfterPrepare"]); wd, wi = uncombine(want["shadingAfterPrepare"])
    assert _frac_bad(gd, wd, 1.5 / 2048) < 0.002 and _frac_bad(gi, wi, 1.5 / 2048) < 0.002
    assert _frac_bad(got["motion"], want["motion"], 2e-2) < 0.01, "motion vectors"
    assert _frac_bad(got["moments"], want["moments"], 2e-3) < 0.01, "luminance moments"
    assert (got["featuresOut"] != want["featuresOut"]).any(axis=-1).mean() < 0.01, "history counters"
    # a-trous phases
    for k in ("phase1", "phase2"):
        gd, gi = uncombine(got[k]); wd, wi = uncombine(want[k])
        assert _frac_bad(gd, wd, 4.0 / 2048) < 0.01 and _frac_bad(gi, wi, 4.0 / 2048) < 0.01, k
    assert _frac_bad(got["phase3"][..., :3], want["phase3"][..., :3], 5e-3) < 0.01, "phase 3 (remodulated, sqrt)"
    if taa:
        # The reference's TAApass reads and writes one buffer in place (a race): its own output changes from run to run (measured
        # with tools/filter_determinism_probe.py: every run differs, up to 2 % of the target pixels by more than 1e-2), ours is
        # deterministic (test_filter_chain_is_deterministic). Two checks that do not depend on how one run of the race falls:
        #  (1) against the per-pixel [min, max] envelope of eight runs of the reference kernels (as tests/test_oracle_golden.py does with
        #      the committed envelope of twelve runs);
        st = dict(w=W, h=H, samplesTaken=1, camIsStationary=stationary, taa=taa, directClamp=15.0, indirectClamp=15.0, j0=0.0, j1=0.0, prevj0=0.0, prevj1=0.0,
                  prevView=prev_view)
        runs = [want] + [orc.ref_filter_gpu(inputs, st) for _ in range(7)]
        for k, tol, bound in (("taaPixels", 1e-2, 0.06), ("target", 3e-2, 0.05)):
            stack = np.stack([np.asarray(r[k])[..., :3] for r in runs])
            lo, hi, x = stack.min(axis=0), stack.max(axis=0), got[k][..., :3]
            assert float(((x < lo - tol) | (x > hi + tol)).any(axis=-1).mean()) < bound, k
        #  (2) against the CPU restatement of the chain (race-free by construction like ours, pinned to the reference by the golden envelope):
        #      deterministic on both sides
        cpu = orc.filter_chain_cpu(inputs, st)
        assert _frac_bad(got["taaPixels"][..., :3], cpu["taaPixels"][..., :3], 1e-2) < 0.02, "TAA vs the CPU restatement"
        assert _frac_bad(got["target"][1:-1, 1:-1, :3], cpu["target"][1:-1, 1:-1, :3], 3e-2) < 0.02, "unsharp + un-gamma vs the CPU restatement"
    else:
        assert _frac_bad(got["target"][..., :3], want["target"][..., :3], 5e-3) < 0.01, "finalizeNoTAA"
    assert np.isfinite(got["target"]).all() and got["target"][1:-1, 1:-1, :3].mean() > 0.05


def test_filter_chain_is_deterministic(case):
    """Our chain (incl. the queued, persistent-thread diamond search of the prepare stage) gives bit-identical buffers run after run."""
    core, inputs, prev_view = case
    st = dict(w=W, h=H, samplesTaken=1, camIsStationary=0, taa=1, directClamp=15.0, indirectClamp=15.0, j0=0.0, j1=0.0, prevj0=0.0, prevj1=0.0,
              prevView=prev_view)
    runs = []
    for _ in range(3):
        io, got, keep = orc.make_filter_io(inputs, st)
        core.FilterChain(io)
        runs.append({k: np.array(v, copy=True) for k, v in got.items()})
    for r in runs[1:]:
        for k in r:
            a, b = r[k], runs[0][k]
            assert np.array_equal(a.view(np.uint32) if a.dtype == np.float32 else a, b.view(np.uint32) if b.dtype == np.float32 else b), k


def test_filter_reduces_noise(case):
    """Property: the filtered image is smoother than the noisy input it was given (the point of the stage)."""
    core, inputs, prev_view = case
    got, _ = _run_both(core, inputs, prev_view, 1, 0)
    d0, i0 = uncombine(got["shadingAfterPrepare"]); d2, i2 = uncombine(got["phase2"])
    rough = lambda x: np.abs(np.diff(x, axis=1)).mean()
    assert rough(d2) < 0.6 * rough(d0) and rough(i2) < 0.6 * rough(i0)
