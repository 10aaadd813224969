"""CPU: the band layout of tile-sharded frames (lh2b_tile_layout, csrc/tile_gather.cu) is pure host arithmetic in the C-ABI library
and must partition the frame: every row belongs to exactly one rank, bands consist of whole 4-row tile rows (the generate kernel
enumerates 8x4-pixel tiles), rank 0's band is contiguous and scales with the root share, the peers are balanced to one tile row."""
import ctypes
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from lighthouse2_b200 import capi


def layout(height, world, share, rank):
    lib = capi.load_library()
    y0, y1, step = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    rc = lib.lh2b_tile_layout(height, world, share, rank, ctypes.byref(y0), ctypes.byref(y1), ctypes.byref(step))
    return rc, y0.value, y1.value, step.value


def rows_of(y0, y1, step):
    return [y for y in range(y0, y1) if ((y // 4) - (y0 // 4)) % step == 0]


@settings(max_examples=200, deadline=None)
@given(st.integers(1, 16), st.integers(1, 700), st.floats(0.0, 1.0))
def test_bands_partition_the_frame(world, quarter_rows, share):
    height = 4 * max(quarter_rows, world)
    owner = np.full(height, -1)
    counts = []
    for r in range(world):
        rc, y0, y1, step = layout(height, world, share, r)
        assert rc == 0 and y0 % 4 == 0 and 0 <= y0 < y1 <= height and step >= 1
        rows = rows_of(y0, y1, step)
        assert (owner[rows] == -1).all(), "a row is claimed twice"
        owner[rows] = r
        counts.append(len(rows))
    assert (owner >= 0).all(), "a row belongs to nobody"
    if world > 1:
        assert counts[0] >= 4 and counts[0] <= max(4, height // world)
        assert max(counts[1:]) - min(counts[1:]) <= 4            # peers: balanced to one tile row
    else:
        assert counts == [height]


def test_root_share_and_errors():
    assert layout(2160, 8, 1.0, 0)[1:] == (0, 268, 1)          # 2160 / 8 = 270 rows -> 268 (multiple of 4)
    assert layout(2160, 8, 0.1, 0)[1:] == (0, 24, 1)
    assert layout(2160, 8, 1.0, 3)[1:] == (268 + 8, 2160, 7)
    assert layout(2160, 2, 1.0, 1)[1:] == (1080, 2160, 1)       # one peer: a contiguous band
    assert layout(16, 8, 1.0, 0)[0] != 0                        # fewer than 4 rows per rank
    assert layout(1082, 4, 1.0, 0)[0] != 0                      # interleaving needs whole tile rows
    assert layout(1082, 2, 1.0, 1)[0] == 0                      # contiguous bands do not
    assert layout(1080, 4, 1.0, 4)[0] != 0 and layout(1080, 17, 1.0, 0)[0] != 0


# ---- sharded filter chain (Setting "tileFilterShard"): filter bands, halos, and which rendered rows travel where -----------------------
def shard_layout(height, world, interleave, rank):
    lib = capi.load_library()
    band, fb, halo, wp = (ctypes.c_int * 3)(), (ctypes.c_int * 2)(), (ctypes.c_int * 2)(), (ctypes.c_int * 2)()
    rc = lib.lh2b_tile_shard_layout(height, world, interleave, rank, band, fb, halo, wp)
    return rc, tuple(band), tuple(fb), tuple(halo), tuple(wp)


def rows_inside(band, e0, e1):
    lib = capi.load_library()
    first, count = ctypes.c_int(), ctypes.c_int()
    assert lib.lh2b_tile_rows_inside(band[0], band[1], band[2], e0, e1, ctypes.byref(first), ctypes.byref(count)) == 0
    return [4 * (first.value + j * band[2]) + k for j in range(count.value) for k in range(4)]


@settings(max_examples=150, deadline=None)
@given(st.integers(2, 8), st.integers(8, 700), st.integers(0, 1))
def test_shard_layout_partitions_rendering_and_filtering(world, quarter_rows, interleave):
    height = 4 * quarter_rows
    layouts = [shard_layout(height, world, interleave, r) for r in range(world)]
    if any(l[0] != 0 for l in layouts):
        assert all(l[0] != 0 for l in layouts)                  # too small a frame is refused for every rank alike
        rows_per_band = ((height + world - 1) // world + 15) // 16 * 16
        assert (world - 1) * rows_per_band >= height
        return
    rendered, filtered = np.full(height, -1), np.full(height, -1)
    for r, (_, band, fb, halo, wp) in enumerate(layouts):
        rows = [y for y in rows_of(*band) if y < height]
        assert (rendered[rows] == -1).all()
        rendered[rows] = r
        assert fb[0] % 16 == 0 and (fb[1] % 16 == 0 or fb[1] == height) and fb[0] < fb[1]
        assert (filtered[fb[0]:fb[1]] == -1).all()
        filtered[fb[0]:fb[1]] = r
        assert halo == (max(0, fb[0] - 16), min(height, fb[1] + 16))
        assert wp[0] <= halo[0] and wp[1] >= halo[1] and wp[0] % 4 == 0
    assert (rendered >= 0).all() and (filtered >= 0).all()
    # every row a rank needs (band + halo; world positions: the wider strip) is rendered by exactly one rank, and the per-sender
    # strided copies (lh2b_tile_rows_inside) deliver exactly those rows
    for d, (_, band_d, fb, halo, wp) in enumerate(layouts):
        for lo, hi in (halo, wp):
            got = np.zeros(height, int)
            for s, (_, band_s, *_rest) in enumerate(layouts):
                rows = [y for y in rows_inside(band_s, lo, hi) if y < height]
                assert all(lo <= y < hi for y in rows) and all(rendered[y] == s for y in rows)
                got[rows] += 1
            assert (got[lo:hi] == 1).all() and got.sum() == hi - lo


def test_shard_layout_values():
    assert shard_layout(2160, 8, 1, 3)[1:] == ((12, 2160, 8), (816, 1088), (800, 1104), (752, 1152))
    assert shard_layout(2160, 8, 0, 7)[1:] == ((1904, 2160, 1), (1904, 2160), (1888, 2160), (1840, 2160))
    assert shard_layout(2160, 2, 1, 0)[1:] == ((0, 2160, 2), (0, 1088), (0, 1104), (0, 1152))
    assert shard_layout(160, 8, 1, 0)[0] != 0                   # 8 bands of 32 rows do not fit 160 rows
    assert shard_layout(2162, 4, 1, 0)[0] != 0                  # whole tile rows only
    assert shard_layout(2160, 9, 1, 0)[0] != 0 and shard_layout(2160, 1, 1, 0)[0] != 0
