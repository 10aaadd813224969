"""Row-band rendering (lh2b_set_row_band), the building block of tile-sharded frames (csrc/tile_gather.cu, SURVEY.md 8e
partitioning 2): a core that renders rows [y0, y1) only must produce, for those rows, exactly what the whole-frame render
produces - same path indices, seeds, buffers - and must leave the other rows alone. 1 spp: bit-exact (one path per pixel, no
accumulation-order freedom); more samples per pixel: float sums in a different order."""
import numpy as np
import pytest

from lighthouse2_b200 import RenderCore, scenes

pytestmark = pytest.mark.gpu
W, H = 160, 96


def _core(sd, spp, filt):
    core = RenderCore()
    core.SetTarget(W, H, spp)
    core.Setting("epsilon", 1e-3)
    core.Setting("maxPathLength", 4)
    core.Setting("filter", 1 if filt else 0)
    sd.upload(core)
    return core


def _rows(band):
    y0, y1, step = band
    return np.array([y for y in range(y0, y1) if ((y // 4) - (y0 // 4)) % step == 0])


@pytest.mark.parametrize("filt", [False, True], ids=["plain", "filter-mode"])
@pytest.mark.parametrize("band", [(0, 32, 1), (20, 52, 1), (64, 96, 1), (16, 96, 3), (8, 96, 7)],
                         ids=["top", "middle-unaligned", "bottom", "every-3rd-tile-row", "every-7th-tile-row"])
def test_row_band_equals_rows_of_the_whole_frame(band, filt):
    sd = scenes.config2_scene(48, 32, n_materials=6, light_quads=2, floaters=300)
    views = [scenes.view_pyramid((2 * k, 30, -80), (0, 0, 0), 40, W, H) for k in range(3)]
    y0, y1, step = band
    rows = _rows(band)
    full, part = _core(sd, 1, filt), _core(sd, 1, filt)
    part.SetRowBand(y0, y1, step)
    for v in views:                      # Restart frames: seeds evolve, history bits of the features persist
        full.Render(v, 1), part.Render(v, 1)
        if filt:
            ff, fw, fd, fa = full.ReadFilterBuffers()
            pf, pw, pd, pa = part.ReadFilterBuffers()
            # the low 4 bits of features.w are the history counter, owned by the filter's prepare pass (which sees only this band here)
            ff, pf = ff.copy(), pf.copy()
            ff[..., 3] &= 0xFFFFFFF0
            pf[..., 3] &= 0xFFFFFFF0
            for a, b in ((ff, pf), (fw, pw), (fd, pd)):          # bit patterns: packed words and NaN-able floats live in these buffers
                assert np.array_equal(a[rows].view(np.uint32), b[rows].view(np.uint32))
            assert np.array_equal(fa[:, rows].view(np.uint32), pa[:, rows].view(np.uint32))
        else:
            assert np.array_equal(full.ReadAccumulator()[rows], part.ReadAccumulator()[rows])
            others = np.setdiff1d(np.arange(H), rows)
            assert not part.ReadAccumulator()[others].any()
    st_full, st_part = full.GetCoreStats(), part.GetCoreStats()
    assert int(st_part["primaryRayCount"]) == W * len(rows) and int(st_full["primaryRayCount"]) == W * H
    assert 0 < int(st_part["totalRays"]) < int(st_full["totalRays"])
    # back to the whole frame
    part.SetRowBand(0, 0)
    part.Render(views[0], 1)
    assert int(part.GetCoreStats()["primaryRayCount"]) == W * H
    full.Shutdown(), part.Shutdown()


def test_row_band_multi_spp_and_untouched_rows():
    sd = scenes.config2_scene(48, 32, n_materials=4, light_quads=2, floaters=200)
    view = scenes.view_pyramid((0, 30, -80), (0, 0, 0), 40, W, H)
    full, part = _core(sd, 4, False), _core(sd, 4, False)
    part.SetRowBand(24, 72)
    full.Render(view, 1), part.Render(view, 1)
    a, b = full.ReadAccumulator(), part.ReadAccumulator()
    np.testing.assert_allclose(b[24:72], a[24:72], rtol=1e-5, atol=1e-6)
    assert not b[:24].any() and not b[72:].any()
    full.Shutdown(), part.Shutdown()


def test_row_band_must_start_on_a_tile_row():
    """Bands are made of 4-row tile rows: a band starting inside one would shade rows it never generated."""
    sd = scenes.config2_scene(12, 8, n_materials=1, light_quads=1)
    core = _core(sd, 1, False)
    with pytest.raises(Exception, match="multiple of 4"):
        core.SetRowBand(2, 40)
    core.SetRowBand(4, 40)
    core.Shutdown()
