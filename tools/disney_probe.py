"""Debug aid: ours vs reference Disney shade kernel vs CPU oracle at path length 1, mismatch statistics per material."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import test_shade_stage_gpu as t
from oracle import binding as orc

sd, core, oracle, view = t._setup(bsdf=1)
O4, D4, T4 = t._primary_state(view)
hits = core.TraceRays(O4, D4)
L, R0, shift = 1, (0x9E3779B9 + 91771) & 0xFFFFFFFF, 0x5A17C3E1
(ext, sh, acc), want, ref = t._run_all(core, oracle, view, L, O4, D4, T4, hits, R0, shift)
rext, rsh, racc, cnt = ref
fl = want["flags"]
oe = dict(O=want["extO"][(fl & 1) > 0], D=want["extD"][(fl & 1) > 0], T=want["extT"][(fl & 1) > 0])
prim = hits[:, 2].view(np.int32) if hits.dtype != np.int32 else hits[:, 2]
hb = np.ascontiguousarray(hits).view(np.uint32).reshape(-1, 4)
mesh0_tris = sd.meshes[0][1]
def mat_of(pathidx):
    inst = hb[pathidx, 1].astype(np.int64); pr = hb[pathidx, 2].astype(np.int64)
    out = np.full(len(pathidx), -1)
    for k in range(len(pathidx)):
        m = sd.instances[inst[k]][0]
        out[k] = sd.meshes[m][1]["material"][pr[k]]
    return out
def cmp(a, b, name):
    ka, kb = t._key(a["O"]), t._key(b["O"])
    ia, ib = np.argsort(ka), np.argsort(kb)
    common, pa, pb = np.intersect1d(ka[ia], kb[ib], return_indices=True)
    print(f"== {name}: {len(ka)} vs {len(kb)}, common {len(common)}")
    mats = mat_of(common)
    for f in ("O", "D", "T"):
        xa, xb = a[f][ia][pa][:, :3].astype(np.float64), b[f][ib][pb][:, :3].astype(np.float64)
        err = (np.abs(xa - xb) / (1e-3 + np.abs(xb))).max(axis=1)
        bad = err > 2e-3
        print(f"  {f}: bad {bad.sum()} ({bad.mean():.4f}), by material kind:", {int(k): int(bad[mats % 8 == k].sum()) for k in range(8)}, "counts", {int(k): int((mats % 8 == k).sum()) for k in range(8)})
    pw_a, pw_b = a["T"][ia][pa][:, 3], b["T"][ib][pb][:, 3]
    e = np.abs(pw_a - pw_b) / (1e-3 + np.abs(pw_b)); bad = e > 2e-3
    print("  pdf: bad", bad.sum(), {int(k): int(bad[mats % 8 == k].sum()) for k in range(8)})
    if bad.any():
        i = np.nonzero(bad)[0][:5]
        for j in i: print("   e.g. path", common[j], "mat", mats[j], "pdf", pw_a[j], pw_b[j], "T", a["T"][ia][pa][j, :3], b["T"][ib][pb][j, :3])
cmp(ext, rext, "ours vs reference kernel")
cmp(oe, rext, "oracle vs reference kernel")
cmp(ext, oe, "ours vs oracle")
# shadow rays
def cmps(a, b, name):
    ka, kb = a["E"][:, 3].view(np.uint32), b["E"][:, 3].view(np.uint32)
    ia, ib = np.argsort(ka), np.argsort(kb)
    common, pa, pb = np.intersect1d(ka[ia], kb[ib], return_indices=True)
    xa, xb = a["E"][ia][pa][:, :3].astype(np.float64), b["E"][ib][pb][:, :3].astype(np.float64)
    err = (np.abs(xa - xb) / (1e-3 + np.abs(xb))).max(axis=1); bad = err > 2e-3
    mats = mat_of(common)
    print(f"== shadow {name}: {len(ka)} vs {len(kb)} common {len(common)} bad {bad.sum()}", {int(k): int(bad[mats % 8 == k].sum()) for k in range(8)})
os_ = dict(O=want["shO"][(fl & 2) > 0], D=want["shD"][(fl & 2) > 0], E=want["shE"][(fl & 2) > 0])
cmps(sh, rsh, "ours vs ref"); cmps(os_, rsh, "oracle vs ref")
