"""Is a 4K filter-mode frame of the C5 workload reproducible bit for bit on several core instances of one process (each builds its own
acceleration structure; LH2B_SET_bvhBuilder=1 selects the deterministic host builder)? Prints the number of differing accumulator pixels.
    python tools/determinism_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lighthouse2_b200 import RenderCore, scenes

W, H = 3840, 2160
sd = scenes.config2_scene(1000, 500, n_materials=64, light_quads=8)
views = [scenes.view_pyramid((0.2 * k, 30, -80 + 0.1 * k), (0, 0, 0), 40, W, H) for k in range(3)]


def make(builder=None):
    c = RenderCore(0)
    c.SetTarget(W, H, 1)
    c.Setting("epsilon", 1e-3), c.Setting("filter", 1), c.Setting("TAA", 1)
    if builder is not None:
        c.Setting("bvhBuilder", builder)
    for k, v in os.environ.items():
        if k.startswith("LH2B_SET_"):
            c.Setting(k[9:], float(v))
    sd.upload(c)
    return c


def frames(c):
    out = []
    for v in views:
        c.Render(v, 1)
        out.append(c.ReadFilterBuffers()[3].copy())
    return out


def diff(a, b, what):
    for k, (x, y) in enumerate(zip(a, b)):
        d = (x.view(np.uint32) != y.view(np.uint32)).any(axis=(0, 3))
        ys, xs = np.nonzero(d)
        print(f"{what}: frame {k}: {len(ys)} accumulator pixels differ" + (f", e.g. (y{ys[0]} x{xs[0]}): {x[:, ys[0], xs[0]].ravel()} vs {y[:, ys[0], xs[0]].ravel()}" if len(ys) else ""), flush=True)


mode = sys.argv[1] if len(sys.argv) > 1 else "order"
if mode == "order":
    a = make()
    fa = frames(a)
    b = make()
    fb = frames(b)
    c = make()
    fc = frames(c)
    diff(fa, fb, "first core vs second core")
    diff(fb, fc, "second core vs third core")
elif mode == "create-first":     # both cores exist before either renders
    a, b = make(), make()
    fa, fb = frames(a), frames(b)
    diff(fa, fb, "created a, b; rendered a, then b")
elif mode == "b-first":          # ... and the second one renders first
    a, b = make(), make()
    fb, fa = frames(b), frames(a)
    diff(fa, fb, "created a, b; rendered b, then a")
elif mode == "geometry":         # do the two cores hold the same vertex / shading-triangle data after the upload?
    a, b = make(), make()
    for i, (v, t) in enumerate(sd.meshes):
        n = np.asarray(v).reshape(-1, 4).shape[0] // 3
        va, ta = a.ReadGeometry(i, n)
        vb, tb = b.ReadGeometry(i, n)
        same_v = np.array_equal(va.view(np.uint32), vb.view(np.uint32))
        same_t = ta.tobytes() == tb.tobytes()
        same_in = np.array_equal(va.view(np.uint32), np.ascontiguousarray(v, np.float32).reshape(-1, 4).view(np.uint32))
        print(f"mesh {i}: {n} triangles, vertices identical {same_v} (and equal to the input: {same_in}), shading triangles identical {same_t}", flush=True)
        if not same_t:
            ba, bb = np.frombuffer(ta.tobytes(), np.uint32).reshape(n, -1), np.frombuffer(tb.tobytes(), np.uint32).reshape(n, -1)
            tri, word = np.nonzero(ba != bb)
            print(f"   {len(tri)} words differ; triangles {np.unique(tri)[:8]}, word offsets {np.unique(word)}", flush=True)
elif mode == "tables":           # ... and the same materials, lights, instance descriptors, sampler tables, sky, acceleration structure?
    a, b = make(), make()
    for name in ("materials", "triLights", "pointLights", "spotLights", "dirLights", "instDesc", "blueNoise", "sky", "argb32", "argb128", "nrm32", "instTrav", "nodes", "tris"):
        ta, tb = a.DebugReadTable(name), b.DebugReadTable(name)
        same = ta.shape == tb.shape and np.array_equal(ta, tb)
        where = "" if same or ta.shape != tb.shape else f"; first differing byte {int(np.nonzero(ta != tb)[0][0])}, {int((ta != tb).sum())} bytes differ"
        print(f"{name}: {ta.size} bytes, identical {same}{where}", flush=True)
elif mode == "frame-state":      # the wavefront buffers as one frame (maximum path length 3) leaves them: where do the two cores part?
    a, b = make(), make()
    a.Render(views[0], 1), b.Render(views[0], 1)
    ca, cb = a.DebugReadTable("counters").view(np.uint32), b.DebugReadTable("counters").view(np.uint32)
    print("extension rays per path length", ca[:5], cb[:5], "shadow rays", ca[18:23], cb[18:23], flush=True)
    n2, n3s = int(ca[2]), int(ca[18 + 3])           # rays shade( 2 ) emitted = paths of length 3; shadow rays of shade( 3 )

    def rows(core, name, n):
        return core.DebugReadTable(name).view(np.float32).reshape(-1, 4)[:n]

    def cmp(what, xa, xb, key_a, key_b):
        ia, ib = np.argsort(key_a, kind="stable"), np.argsort(key_b, kind="stable")
        same_keys = np.array_equal(key_a[ia], key_b[ib])
        d = (xa[ia].view(np.uint32) != xb[ib].view(np.uint32)).any(axis=1) if same_keys else None
        print(f"{what}: {len(xa)} entries, same path set {same_keys}" + ("" if d is None else f", {int(d.sum())} entries differ" + (f", first key {int(key_a[ia][np.nonzero(d)[0][0]])}: {xa[ia][np.nonzero(d)[0][0]]} vs {xb[ib][np.nonzero(d)[0][0]]}" if d.any() else "")), flush=True)

    # path length 3 reads set 0 (written by shade( 2 )): O.w carries the path index << 6 | flags
    Oa, Ob = rows(a, "path0O", n2), rows(b, "path0O", n2)
    ka, kb = Oa[:, 3].view(np.uint32) >> 6, Ob[:, 3].view(np.uint32) >> 6
    for f in ("O", "D", "T"):
        cmp(f"paths of length 3, {f} (output of shade 2)", rows(a, "path0" + f, n2), rows(b, "path0" + f, n2), ka, kb)
    cmp("hits of extend 3", rows(a, "hits", n2), rows(b, "hits", n2), ka, kb)
    bad_key = int(os.environ.get("PROBE_KEY", "5388903"))
    for tag, O_, k_, core in (("first core", Oa, ka, a), ("second core", Ob, kb, b)):
        i = np.nonzero(k_ == bad_key)[0]
        if len(i):
            h = rows(core, "hits", n2)[i[0]]
            print(f"{tag}: path {bad_key} at slot {int(i[0])}: O {O_[i[0]]} flags {int(O_[i[0], 3].view(np.uint32)) & 63} D {rows(core, 'path0D', n2)[i[0]]} T {rows(core, 'path0T', n2)[i[0]]} "
                  f"hit words {h.view(np.uint32)} (as float {h}, as int {h.view(np.int32)})", flush=True)
    Ea, Eb = rows(a, "connE", n3s), rows(b, "connE", n3s)
    for f in ("O", "D", "E"):
        cmp(f"shadow rays of shade 3, {f}", rows(a, "conn" + f, n3s), rows(b, "conn" + f, n3s), Ea[:, 3].view(np.uint32), Eb[:, 3].view(np.uint32))
elif mode == "tables-after":     # does a frame write into one of the read-only tables (a stray store into a neighbouring allocation)?
    a, b = make(), make()
    names = ("materials", "triLights", "pointLights", "spotLights", "dirLights", "instDesc", "blueNoise", "sky", "argb32", "argb128", "nrm32", "instTrav", "nodes", "tris")
    before = {(c, n): c.DebugReadTable(n).copy() for c in (a, b) for n in names}
    for v in views:
        a.Render(v, 1), b.Render(v, 1)
    for tag, c in (("first core", a), ("second core", b)):
        for n in names:
            now = c.DebugReadTable(n)
            if not np.array_equal(now, before[(c, n)]):
                idx = np.nonzero(now != before[(c, n)])[0]
                print(f"{tag}: table {n} CHANGED during rendering: {len(idx)} bytes, first at byte {int(idx[0])} (of {now.size}): {before[(c, n)][idx[0]:idx[0] + 16]} -> {now[idx[0]:idx[0] + 16]}", flush=True)
        print(f"{tag}: checked", flush=True)
