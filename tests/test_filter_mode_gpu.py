"""Filter mode end to end (Setting("filter", 1)): the shade stage writes the per-pixel features of the first diffuse
vertex and splits direct / indirect light; FinalizeRender runs the SVGF / TAA chain.
 (1) features, world positions, depth derivatives and both accumulator halves against the CPU oracle;
 (2) the frame the core presents against the REFERENCE's filter kernels fed with the core's own buffers (pipeline wiring:
     buffer roles, history rotation, settings);
 (3) over a moving-camera sequence the output stays finite and is smoother than the unfiltered estimate."""
import numpy as np
import pytest

from lighthouse2_b200 import RenderCore, scenes
from oracle import binding as orc
from util import rel_rmse

pytestmark = pytest.mark.gpu
W, H = 160, 96


def _scene():
    sd = scenes.config2_scene(48, 32, n_materials=6, light_quads=2, floaters=300)
    sd.materials[1]["roughness"]["value"] = 0.0      # a mirror: exercises the via-specular feature path
    return sd


def _core(sd, taa):
    core = RenderCore()
    core.SetTarget(W, H, 1)
    core.Setting("epsilon", 1e-3)
    core.Setting("filter", 1)
    core.Setting("TAA", taa)
    core.Setting("clampDirect", 15.0)
    core.Setting("clampIndirect", 15.0)
    sd.upload(core)
    return core


def test_features_and_split_accumulator_match_oracle():
    sd = _scene()
    view = scenes.view_pyramid((0, 30, -80), (0, 0, 0), 40, W, H)
    core = _core(sd, 0)
    core.Render(view, 1)
    feat, wp, dd, acc = core.ReadFilterBuffers()
    o = orc.FrameOracle(sd, W, H, 1, 1e-3, 10.0, 3, 1, filter=True)
    o.render(view, 1)
    # packed words: albedo (10/11/11 bit), normal (3 x 10 bit + specular bit), depth bits, specular / material / history bits
    # (the history counter in the low 4 bits of .w belongs to prepareFilter, which the core has already run; depth is
    #  compared as a float below: the primary rays differ in the last bits between libm and the device intrinsics)
    assert (feat[..., 0] == o.features[..., 0]).mean() > 0.99, "albedo words"
    nd = lambda a, sh: ((a >> sh) & 1023).astype(np.int64)
    n_close = np.ones((H, W), bool)
    for sh in (2, 12, 22):
        n_close &= np.abs(nd(feat[..., 1], sh) - nd(o.features[..., 1], sh)) <= 1
    assert n_close.mean() > 0.99 and ((feat[..., 1] & 3) == (o.features[..., 1] & 3)).mean() > 0.995, "packed normals / specular bit"
    assert ((feat[..., 3] >> 4) == (o.features[..., 3] >> 4)).mean() > 0.995, "specular / material bits"
    close_depth = np.abs(feat[..., 2].view(np.float32) / np.maximum(o.features[..., 2].view(np.float32), 1e-6) - 1) < 1e-4
    assert close_depth.mean() > 0.995
    assert (np.abs(wp[..., :3] - o.world_pos[..., :3]) < 1e-2 + 1e-5 * np.abs(o.world_pos[..., :3])).all(axis=-1).mean() > 0.995
    assert ((wp[..., 3].view(np.uint32) & 3) == (o.world_pos[..., 3].view(np.uint32) & 3)).mean() > 0.995
    assert (np.abs(dd - o.delta_depth) < 1e-3 * (1 + np.abs(o.delta_depth))).all(axis=-1).mean() > 0.99
    assert rel_rmse(acc[0], o.accum[0]) < 0.02 and rel_rmse(acc[1], o.accum[1]) < 0.05
    assert acc[1, ..., :3].sum() > 0 and acc[0, ..., :3].sum() > 0
    core.Shutdown()


@pytest.mark.skipif(not orc.have_ref_filter_gpu(), reason="oracle/_ref/libref_filter_gpu.so is built only where /root/reference exists")
@pytest.mark.parametrize("taa", [0, 1])
def test_presented_frame_matches_reference_chain_on_same_buffers(taa):
    sd = _scene()
    views = [scenes.view_pyramid((0.0, 30, -80), (0, 0, 0), 40, W, H), scenes.view_pyramid((0.5, 30.1, -79.6), (0, 0, 0), 40, W, H)]
    core = _core(sd, taa)
    hist = dict(prevWorldPos=np.zeros((H, W, 4), np.float32), prevMoments=np.zeros((H, W, 4), np.float32),
                filteredIN=np.zeros((H, W, 4), np.float32), prevPixels=np.zeros((H, W, 4), np.float32))
    feat_hist = np.zeros((H, W, 4), np.uint32)
    prev_view = views[0]
    for k, view in enumerate(views):
        core.Render(view, 1)                       # every frame restarts the accumulator (camera moved)
        got = core.ReadPixels()
        feat, wp, dd, acc = core.ReadFilterBuffers()
        # history counter bits: the core's features buffer already holds the post-prepare counters; the reference starts
        # from last frame's counters and the fresh feature words
        feat_in = feat.copy()
        feat_in[..., 3] = (feat[..., 3] & ~np.uint32(15)) | (feat_hist[..., 3] & 15)
        st = dict(w=W, h=H, samplesTaken=1, camIsStationary=0, taa=taa, directClamp=15.0, indirectClamp=15.0, j0=0.0, j1=0.0, prevj0=0.0, prevj1=0.0,
                  prevView=prev_view if k else view)
        inputs = dict(accumulator=acc, features=feat_in, worldPos=wp, deltaDepth=dd, **hist)
        ref = orc.ref_filter_gpu(inputs, st)
        inner = (slice(1, H - 1), slice(1, W - 1))
        bad = (np.abs(got[inner][..., :3] - ref["target"][inner][..., :3]) > 3e-2).any(axis=-1).mean()
        # The reference TAA pass rewrites `pixels` while neighbouring threads still read it; with an empty history (frame 0)
        # that race moves many pixels, so the reference is only a loose bound there. The wiring itself is checked exactly
        # against the same chain run stand-alone (which tests/test_filter_gpu.py pins to the reference stage by stage).
        assert bad < (0.05 if not taa else (0.3 if k == 0 else 0.15)), f"frame {k}: {bad:.3f} of the pixels differ from the reference chain"
        io, own, keep = orc.make_filter_io(inputs, st)
        core.FilterChain(io)
        assert np.array_equal(got[inner], own["target"][inner]), f"frame {k}: presented frame differs from the stand-alone chain on the same buffers"
        # next frame's history = what the core itself now holds (equal to the stand-alone chain's outputs, just verified)
        hist = dict(prevWorldPos=wp, prevMoments=own["moments"], filteredIN=own["phase1"], prevPixels=own["taaPixels"] if taa else own["phase3"])
        feat_hist, prev_view = own["featuresOut"], view
    core.Shutdown()


def test_sequence_is_finite_and_smoother_than_unfiltered():
    sd = _scene()
    core = _core(sd, 1)
    raw = RenderCore()
    raw.SetTarget(W, H, 1); raw.Setting("epsilon", 1e-3); sd.upload(raw)
    rough = lambda img: float(np.abs(np.diff(img[..., :3], axis=1)).mean())
    for k in range(5):
        view = scenes.view_pyramid((0.3 * k, 30, -80 + 0.2 * k), (0, 0, 0), 40, W, H)
        core.Render(view, 1); raw.Render(view, 1)
        f, u = core.ReadPixels(), raw.ReadPixels()
        assert np.isfinite(f).all()
        if k >= 2:
            assert rough(np.sqrt(np.clip(f, 0, None))) < 0.8 * rough(np.sqrt(np.clip(u, 0, 10)))   # compare in the same (gamma) domain
    st = core.GetCoreStats()
    assert st["filterTime"] > 0
    core.Shutdown(); raw.Shutdown()


@pytest.mark.parametrize("precise", [0, 1], ids=["fast-math", "precise-math"])
@pytest.mark.parametrize("taa", [0, 1], ids=["svgf", "svgf+taa"])
def test_frame_sequence_matches_cpu_filtered_oracle(taa, precise):
    """BASELINE.json configs[4] at oracle size, end to end and against the CPU only: seven frames under a moving, then resting
    camera - path tracing with feature writes and the direct / indirect split, prepare incl. reprojection and the diamond search,
    three a-trous passes with the temporal blend, (TAA + unsharp), and the buffer rotation between frames - compared with
    orc.FilteredFrameOracle (frame oracle in filter mode + oracle/lh2_oracle_filter.h, which tests/test_oracle_golden.py pins to the
    reference's own kernels). The oracle uses libm without FMA contraction. Two builds of the device code are checked:
      fast-math (the shipping build, like the reference's -use_fast_math): the temporal feedback (0.9 history weight) and the
        unsharp mask (x 2.7) amplify the differences when TAA is on - without TAA at most 1 % of the interior pixels off by more
        than 3e-2 and relative RMSE below 4 % in every frame (measured <= 0.3 % / 2 %); with TAA at most 6 % off by more than 1e-1
        and relative RMSE below 12 % (measured <= 2.6 % / 6 %);
      precise-math (Setting "preciseMath" 1: IEEE division / sqrt, accurate transcendentals, no contraction in the shade and filter
        stages): without TAA at most 1 % off by more than 1e-2 and relative RMSE below 2.5 % (measured <= 0.6 % / 2 %), with TAA at
        most 2 % off by more than 1e-1 and relative RMSE below 6 % (measured <= 1.2 % / 4.5 %; what remains are isolated paths that
        take another discrete branch and are carried along by the history).
    Frame means within 0.5 % in all cases."""
    sd = scenes.config2_scene(48, 32, n_materials=6, light_quads=2, floaters=300)
    core = _core(sd, taa)
    core.Setting("preciseMath", precise)
    views = [scenes.view_pyramid((0.4 * k, 30 + 0.1 * k, -80 + 0.3 * k), (0, 0, 0), 40, W, H) for k in range(5)]
    views += [scenes.view_pyramid((1.6, 30.4, -78.8), (0, 0, 0), 40, W, H)] * 2
    inner = (slice(16, H - 16), slice(16, W - 16))      # the reference's border quirk spreads 14 pixels inwards (see tests/test_oracle_cpu.py)
    with orc.accel(1):
        fo = orc.FilteredFrameOracle(sd, W, H, taa=bool(taa))
        for k, v in enumerate(views):
            core.Render(v, 1)
            got, want = core.ReadPixels()[..., :3][inner], fo.render(v, 1)[..., :3][inner]
            assert np.isfinite(got).all()
            d = np.abs(got - want)
            rel = float(np.sqrt((d ** 2).mean()) / np.sqrt((want ** 2).mean()))
            off = lambda t: float((d > t).any(-1).mean())
            if taa and precise:
                assert off(1e-1) < 0.02 and rel < 0.06, (k, off(1e-1), rel)
            elif taa:
                assert off(1e-1) < 0.06 and rel < 0.12, (k, off(1e-1), rel)
            elif precise:
                assert off(1e-2) < 0.01 and rel < 0.025, (k, off(1e-2), rel)
            else:
                assert off(3e-2) < 0.01 and rel < 0.04, (k, off(3e-2), rel)
            assert abs(got.mean() / want.mean() - 1) < 0.005, k
    core.Shutdown()
