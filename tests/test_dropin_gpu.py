"""Drop-in proof on the GPU: a stand-in RenderSystem compiled against the REFERENCE's own interface headers
(oracle/_ref/dropin_host_ref, built here from /root/reference) dlopens libRenderCore_B200.so, resolves CreateCore and
drives a frame through the virtual interface; the same program compiled against include/lh2_core_api.h must print
the same numbers."""
import json
import os
import subprocess
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "lighthouse2_b200", "csrc", "libRenderCore_B200.so")
REF_HOST = os.path.join(ROOT, "oracle", "_ref", "dropin_host_ref")
OWN_HOST = os.path.join(ROOT, "tests", "host", "dropin_host_own")


def _run(host):
    out = subprocess.run([host, "render", LIB], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    return json.loads(out.stdout.strip().splitlines()[-1])


def _own_host():
    if not os.path.exists(OWN_HOST):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-I" + os.path.join(ROOT, "include"), "-o", OWN_HOST,
                               os.path.join(ROOT, "tests", "host", "dropin_host.cpp"), "-ldl"])
    return OWN_HOST


def test_own_header_host_renders():
    r = _run(_own_host())
    assert r["readback_rc"] == 0 and "B200" in r["device"] and r["sm"] == 148
    assert r["primary"] == 96 * 64 * 2 and r["shadow"] > 0 and r["total"] == r["extension"] + r["shadow"]
    assert all(0.05 < m < 5 for m in r["mean"])
    assert r["probe"][1] >= 0 and r["probe"][2] > 1.0


@pytest.mark.skipif(not os.path.exists(REF_HOST), reason="oracle/_ref/dropin_host_ref is built only where /root/reference exists")
def test_reference_header_host_matches():
    a, b = _run(REF_HOST), _run(_own_host())
    assert a["flavour"] == "reference-headers" and b["flavour"] == "own-header"
    for k in ("primary", "extension", "shadow", "total", "probe", "device", "sm"):
        assert a[k] == b[k], k
    # 2 spp: two paths add to one pixel with float atomics, so the sums may differ in the last bits from run to run
    assert all(abs(x - y) <= 1e-4 * abs(y) for x, y in zip(a["mean"], b["mean"])), (a["mean"], b["mean"])
