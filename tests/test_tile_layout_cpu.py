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
