"""CPU: the boundary. (1) include/lh2_core_api.h must have the reference's layouts: compared against
tests/golden/abi_reference.txt, which is the output of the host compiled against the reference headers
(oracle/_ref/dropin_host_ref abi; regenerate with tools/make_golden_abi.sh where /root/reference exists).
(2) the shared library loads and exports every symbol include/lh2b.h declares. (3) numpy dtypes agree."""
import os
import re
import subprocess
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "abi_reference.txt")


def _table(text):
    return dict(line.split() for line in text.strip().splitlines())


def test_header_layout_matches_reference(tmp_path):
    exe = str(tmp_path / "host")
    subprocess.check_call(["g++", "-O0", "-std=c++17", "-I" + os.path.join(ROOT, "include"), "-o", exe,
                           os.path.join(ROOT, "tests", "host", "dropin_host.cpp"), "-ldl"])
    own = _table(subprocess.check_output([exe, "abi"], text=True))
    ref = _table(open(GOLDEN).read())
    assert own == ref
    assert len(ref) >= 60


@pytest.mark.skipif(not os.path.exists("/root/reference/lib/RenderSystem/core_api_base.h"), reason="reference tree not present")
def test_golden_abi_is_current():
    """Where the reference is available, the committed golden table must equal a fresh build against its headers."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "_ref/dropin_host_ref"])
    fresh = _table(subprocess.check_output([os.path.join(ROOT, "oracle", "_ref", "dropin_host_ref"), "abi"], text=True))
    assert fresh == _table(open(GOLDEN).read())


def test_library_exports_every_declared_symbol(core_lib):
    hdr = open(os.path.join(ROOT, "include", "lh2b.h")).read()
    declared = set(re.findall(r"LH2B_API\s+[\w\s\*]+?\b(lh2b_\w+)\s*\(", hdr))
    assert len(declared) >= 30
    for name in declared | {"CreateCore"}:
        assert hasattr(core_lib, name), name
    from lighthouse2_b200 import capi
    assert declared <= set(capi.SIGNATURES), declared - set(capi.SIGNATURES)


def test_numpy_dtypes_match_table():
    from lighthouse2_b200 import abi
    t = _table(open(GOLDEN).read())
    assert abi.CoreTri.itemsize == int(t["sizeof.CoreTri"])
    assert abi.CoreMaterial.itemsize == int(t["sizeof.CoreMaterial"])
    for f in ("color", "detailColor", "normals", "detailNormals", "flags", "absorption", "metallic", "roughness", "transmission",
              "eta", "ior", "urough", "Ks", "sigma", "specTrans", "flatness", "opacity"):
        assert abi.CoreMaterial.fields[f][1] == int(t["offsetof.CoreMaterial." + f]), f
    assert abi.Vec3Value.fields["uvscale"][1] == int(t["offsetof.Vec3Value.uvscale"])
    assert abi.ScalarValue.fields["uvscale"][1] == int(t["offsetof.ScalarValue.uvscale"])
    for f in ("ltriIdx", "material", "vN0", "Nx", "T", "area", "B", "alpha", "LOD", "vertex0", "vertex2"):
        assert abi.CoreTri.fields[f][1] == int(t["offsetof.CoreTri." + f]), f
    assert abi.CoreTri.fields["u1"][1] == int(t["offsetof.CoreTri.u1_0"])
    for f in ("energy", "radiance", "triIdx", "instIdx"):
        assert abi.CoreLightTri.fields[f][1] == int(t["offsetof.CoreLightTri." + f])
    for f in ("p1", "aperture", "spreadAngle", "distortion"):
        assert abi.ViewPyramid.fields[f][1] == int(t["offsetof.ViewPyramid." + f])
    for f in ("SMcount", "bvhBuildTime", "totalRays", "renderTime", "primaryRayCount", "traceTime0", "deepRayCount", "shadeTime",
              "probedInstid", "probedWorldPos"):
        assert abi.CoreStats.fields[f][1] == int(t["offsetof.CoreStats." + f]), f
    for f in ("pixelCount", "firstPixel", "storage"):
        assert abi.CoreTexDesc.fields[f][1] == int(t["offsetof.CoreTexDesc." + f])


def test_no_gpu_fails_loudly():
    """Without a CUDA device the core must refuse to initialise (no CPU fallback)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from lighthouse2_b200 import RenderCore, CoreError
    with pytest.raises(CoreError, match="no CPU fallback"):
        RenderCore(0)
