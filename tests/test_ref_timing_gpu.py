"""Ours beside theirs (tools/ref_kernel_timing.py at a size that runs in seconds): the reference's own shadeKernel and filter
kernels (compiled for sm_100a from /root/reference into oracle/_ref) and the product's kernels are timed on identical buffers, kernel
launches only. The product must not be slower on any stage; the full-size numbers live in profiles/r2_reference_kernels.json."""
import os
import sys
import pytest

from oracle import binding as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not (orc.have_ref_shade_gpu(0) and orc.have_ref_filter_gpu()), reason="oracle/_ref is built only where /root/reference exists")]
SLACK = 1.15     # timing noise of sub-millisecond launches


def test_shade_kernel_not_slower_than_the_reference_kernel():
    import ref_kernel_timing as rk
    r = rk.shade_timing(960, 540, nx=400, nz=250)
    assert len(r["levels"]) == 3
    for lv in r["levels"]:
        assert lv["ours_ms"] > 0 and lv["reference_ms"] > 0
        # same work on both sides: the two kernels emit (nearly) the same rays
        assert abs(lv["extension_rays_out"] - lv["reference_extension_rays_out"]) <= 0.002 * lv["paths"] + 4
        assert lv["ours_ms"] <= SLACK * lv["reference_ms"], lv


def test_filter_chain_not_slower_than_the_reference_kernels():
    import ref_kernel_timing as rk
    r = rk.filter_timing(1920, 1080, nx=400, nz=250)
    for case in r["cases"]:
        ours, theirs = case["ours_ms"], case["reference_ms"]
        assert ours["chain"] > 0 and theirs["chain"] > 0
        assert ours["chain"] <= SLACK * theirs["chain"], case
