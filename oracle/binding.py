"""TEST INFRASTRUCTURE ONLY - ctypes binding of oracle/liblh2oracle.so (built by oracle/Makefile)."""
import ctypes
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liblh2oracle.so")
_lib = None


class OrcMesh(ctypes.Structure):
    _fields_ = [("verts4", ctypes.c_void_p), ("triCount", ctypes.c_int)]


class OrcInstance(ctypes.Structure):
    _fields_ = [("mesh", ctypes.c_int), ("xform", ctypes.c_float * 12)]


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        _lib = ctypes.CDLL(_LIB)
    return _lib


def set_accel(mode):
    """0: exhaustive search (the definition, default). 1: the same search pruned by a per-mesh BVH (lh2_oracle_bvh.h);
    identical results, proven by tests/test_oracle_cpu.py::test_bvh_oracle_equals_exhaustive_search. Returns the old mode."""
    old = lib().orc_get_accel()
    lib().orc_set_accel(int(mode))
    return old


class accel:
    """with orc.accel(1): ... - scoped set_accel."""

    def __init__(self, mode=1):
        self.mode = mode

    def __enter__(self):
        self.old = set_accel(self.mode)

    def __exit__(self, *a):
        set_accel(self.old)


def _scene(meshes, instances):
    """meshes: list of float32[3T,4]; instances: list of (meshIdx, 4x4 or None)."""
    keep = [np.ascontiguousarray(m, np.float32).reshape(-1, 4) for m in meshes]
    cm = (OrcMesh * len(keep))()
    for i, m in enumerate(keep):
        cm[i].verts4, cm[i].triCount = m.ctypes.data, m.shape[0] // 3
    ci = (OrcInstance * max(len(instances), 1))()
    for i, (mi, xf) in enumerate(instances):
        ci[i].mesh = mi
        x = np.eye(4, dtype=np.float32) if xf is None else np.asarray(xf, np.float32).reshape(4, 4)
        for k in range(12):
            ci[i].xform[k] = float(x.flat[k])
    return keep, cm, ci


def closest_hits(meshes, instances, O4, D4, threads=None):
    keep, cm, ci = _scene(meshes, instances)
    O4 = np.ascontiguousarray(O4, np.float32).reshape(-1, 4)
    D4 = np.ascontiguousarray(D4, np.float32).reshape(-1, 4)
    n = O4.shape[0]
    hits = np.empty((n, 4), np.uint32)
    lib().orc_closest_hits(cm, len(keep), ci, len(instances), ctypes.c_void_p(O4.ctypes.data), ctypes.c_void_p(D4.ctypes.data),
                           n, ctypes.c_void_p(hits.ctypes.data), threads or os.cpu_count())
    return hits


def occluded(meshes, instances, O4, D4, threads=None):
    keep, cm, ci = _scene(meshes, instances)
    O4 = np.ascontiguousarray(O4, np.float32).reshape(-1, 4)
    D4 = np.ascontiguousarray(D4, np.float32).reshape(-1, 4)
    n = O4.shape[0]
    occ = np.empty(n, np.uint8)
    lib().orc_occluded(cm, len(keep), ci, len(instances), ctypes.c_void_p(O4.ctypes.data), ctypes.c_void_p(D4.ctypes.data),
                       n, ctypes.c_void_p(occ.ctypes.data), threads or os.cpu_count())
    return occ


# ---- full frame -----------------------------------------------------------------------------------

class OrcTexDesc(ctypes.Structure):
    _fields_ = [("data", ctypes.c_void_p), ("width", ctypes.c_uint), ("height", ctypes.c_uint), ("flags", ctypes.c_uint),
                ("pixelCount", ctypes.c_uint), ("firstPixel", ctypes.c_uint), ("mipLevels", ctypes.c_uint),
                ("storage", ctypes.c_int), ("pad", ctypes.c_uint)]


class OrcMaterialIn(ctypes.Structure):
    _fields_ = [("color", ctypes.c_float * 3), ("absorption", ctypes.c_float * 3), ("params", ctypes.c_float * 12),
                ("flags", ctypes.c_uint), ("tex", ctypes.c_int * 6), ("uvscale", (ctypes.c_float * 2) * 6),
                ("uvoffset", (ctypes.c_float * 2) * 6)]


class OrcFrameIn(ctypes.Structure):
    _fields_ = [("meshes", ctypes.c_void_p), ("coreTris", ctypes.c_void_p), ("meshCount", ctypes.c_int),
                ("instances", ctypes.c_void_p), ("instanceCount", ctypes.c_int),
                ("materials", ctypes.c_void_p), ("materialCount", ctypes.c_int),
                ("textures", ctypes.c_void_p), ("textureCount", ctypes.c_int),
                ("triLights", ctypes.c_void_p), ("triLightCount", ctypes.c_int),
                ("pointLights", ctypes.c_void_p), ("pointLightCount", ctypes.c_int),
                ("spotLights", ctypes.c_void_p), ("spotLightCount", ctypes.c_int),
                ("dirLights", ctypes.c_void_p), ("dirLightCount", ctypes.c_int),
                ("skyPixels", ctypes.c_void_p), ("skyW", ctypes.c_int), ("skyH", ctypes.c_int),
                ("worldToSky", ctypes.c_float * 16), ("blueNoiseBytes", ctypes.c_void_p),
                ("w", ctypes.c_int), ("h", ctypes.c_int), ("spp", ctypes.c_int), ("pass_", ctypes.c_int),
                ("sampleBase", ctypes.c_uint), ("shiftSeed", ctypes.c_uint), ("camRNGseed", ctypes.c_uint),
                ("geometryEpsilon", ctypes.c_float), ("clampValue", ctypes.c_float),
                ("maxPathLength", ctypes.c_int), ("enoughBounces", ctypes.c_uint),
                ("view", ctypes.c_float * 17), ("threads", ctypes.c_int), ("bsdfModel", ctypes.c_int)]


PathRecord = np.dtype([("hit", np.uint32, 4), ("firstShadow", np.float32, 8)])
_PARAM_NAMES = ("metallic", "subsurface", "specular", "roughness", "specularTint", "anisotropic", "sheen", "sheenTint",
                "clearcoat", "clearcoatGloss", "transmission", "eta")
_TEX_NAMES = ("color", "detailColor", "normals", "detailNormals", "specular", "roughness")
_BLUENOISE = os.path.join(os.path.dirname(_HERE), "lighthouse2_b200", "data", "heitz_bluenoise_256spp.bin")


class FrameOracle:
    """Stateful CPU counterpart of one render core: keeps samplesTaken and the two host RNG states
    (rendercore.h:122-123) so consecutive Render calls can be mirrored frame by frame."""

    def __init__(self, scene, width, height, spp=1, epsilon=1e-4, clamp=10.0, max_path_length=3, max_diffuse_bounces=1,
                 threads=None, sample_base=0, total_spp=0, filter=False, bsdf=0):
        self.sd, self.w, self.h, self.spp = scene, width, height, spp
        self.eps, self.clamp, self.maxlen = epsilon, clamp, max_path_length
        self.enough = 0 if max_diffuse_bounces <= 0 else (2 if max_diffuse_bounces < 2 else 8)
        self.threads = threads or os.cpu_count()
        self.sample_base, self.total_spp = sample_base, total_spp
        self.samples_taken = 0
        self.shift_seed, self.cam_seed = 0x11331445, 0x12345678
        self.first_converging = True
        self.filter = filter
        self.bsdf = int(bsdf)          # 0: lambert.h model, 1: principled (disney.h) model
        self.accum = np.zeros((2, height, width, 4) if filter else (height, width, 4), np.float32)
        self.features = np.zeros((height, width, 4), np.uint32) if filter else None
        self.world_pos = np.zeros((height, width, 4), np.float32) if filter else None
        self.delta_depth = np.zeros((height, width, 4), np.float32) if filter else None
        self.ray_counts = (0, 0)

    def render(self, view, converge=1, records=False):
        sd = self.sd
        if converge == 1 or self.first_converging:
            self.samples_taken, self.first_converging, self.cam_seed = 0, True, 0x12345678
        if converge == 0:
            self.first_converging = False
        if self.samples_taken == 0:
            self.accum[:] = 0
        f, keep = self.frame_in(view)
        counts = (ctypes.c_uint64 * 2)()
        seeds = (ctypes.c_uint * 2)()
        rec = np.zeros(self.w * self.h * self.spp, dtype=PathRecord) if records else None
        fp = (lambda a: ctypes.c_void_p(a.ctypes.data)) if self.filter else (lambda a: None)
        lib().orc_render_frame(ctypes.byref(f), ctypes.c_void_p(self.accum.ctypes.data), counts, seeds,
                               ctypes.c_void_p(rec.ctypes.data) if records else None, fp(self.features), fp(self.world_pos), fp(self.delta_depth))
        self.shift_seed, self.cam_seed = seeds[0], seeds[1]
        self.ray_counts = (int(counts[0]), int(counts[1]))
        total = self.total_spp if self.total_spp > 0 else self.spp
        self.samples_taken += total
        local = self.samples_taken * self.spp // total
        self.pixels = self.accum / np.float32(local)
        return (self.pixels, rec) if records else self.pixels

    def frame_in(self, view):
        """OrcFrameIn for the current state plus the list of arrays that must stay alive while it is used."""
        sd = self.sd
        keep = []
        verts = [np.ascontiguousarray(v, np.float32).reshape(-1, 4) for v, _ in sd.meshes]
        tris = [np.ascontiguousarray(t) for _, t in sd.meshes]
        keep += verts + tris
        cm = (OrcMesh * len(verts))()
        ct = (ctypes.c_void_p * len(verts))()
        for i, (v, t) in enumerate(zip(verts, tris)):
            cm[i].verts4, cm[i].triCount = v.ctypes.data, v.shape[0] // 3
            ct[i] = t.ctypes.data
        ci = (OrcInstance * max(len(sd.instances), 1))()
        for i, (mi, xf) in enumerate(sd.instances):
            ci[i].mesh = mi
            x = np.eye(4, dtype=np.float32) if xf is None else np.asarray(xf, np.float32).reshape(4, 4)
            for k in range(12):
                ci[i].xform[k] = float(x.flat[k])
        mats = (OrcMaterialIn * len(sd.materials))()
        for i, m in enumerate(sd.materials):
            mats[i].color[:] = [float(x) for x in m["color"]["value"]]
            mats[i].absorption[:] = [float(x) for x in m["absorption"]["value"]]
            mats[i].params[:] = [float(m[n]["value"]) for n in _PARAM_NAMES]
            mats[i].flags = int(m["flags"])
            for k, n in enumerate(_TEX_NAMES):
                mats[i].tex[k] = int(m[n]["textureID"])
                mats[i].uvscale[k][:] = [float(x) for x in m[n]["uvscale"]]
                mats[i].uvoffset[k][:] = [float(x) for x in m[n]["uvoffset"]]
        texs = (OrcTexDesc * max(len(sd.textures), 1))()
        for i, (tex, storage, w, h, mips, *flags) in enumerate(sd.textures):
            tex = np.ascontiguousarray(tex)
            keep.append(tex)
            texs[i].data, texs[i].width, texs[i].height = tex.ctypes.data, w, h
            texs[i].pixelCount, texs[i].mipLevels, texs[i].storage = tex.size // 4, mips, storage
            texs[i].flags = flags[0] if flags else 0      # HostTexture flags (HDR = 8 matters: rendercore.cpp:539)
        f = OrcFrameIn()
        f.meshes, f.coreTris, f.meshCount = ctypes.addressof(cm), ctypes.addressof(ct), len(verts)
        f.instances, f.instanceCount = ctypes.addressof(ci), len(sd.instances)
        f.materials, f.materialCount = ctypes.addressof(mats), len(sd.materials)
        f.textures, f.textureCount = ctypes.addressof(texs), len(sd.textures)
        lights = [np.ascontiguousarray(a) for a in (sd.tri_lights, sd.point_lights, sd.spot_lights, sd.dir_lights)]
        keep += lights
        f.triLights, f.triLightCount = lights[0].ctypes.data, len(lights[0])
        f.pointLights, f.pointLightCount = lights[1].ctypes.data, len(lights[1])
        f.spotLights, f.spotLightCount = lights[2].ctypes.data, len(lights[2])
        f.dirLights, f.dirLightCount = lights[3].ctypes.data, len(lights[3])
        if sd.sky is not None:
            sky = np.ascontiguousarray(sd.sky[0], np.float32)
            keep.append(sky)
            f.skyPixels, f.skyW, f.skyH = sky.ctypes.data, sd.sky[1], sd.sky[2]
        sky_xf = sd.sky[3] if sd.sky is not None and len(sd.sky) > 3 else np.eye(4, dtype=np.float32)
        f.worldToSky[:] = [float(x) for x in np.asarray(sky_xf, np.float32).flat]
        bn = np.fromfile(_BLUENOISE, dtype=np.uint8)
        keep.append(bn)
        f.blueNoiseBytes = bn.ctypes.data
        f.w, f.h, f.spp, f.pass_ = self.w, self.h, self.spp, self.samples_taken
        f.sampleBase = self.sample_base
        f.shiftSeed, f.camRNGseed = self.shift_seed, self.cam_seed
        f.geometryEpsilon, f.clampValue = self.eps, self.clamp
        f.maxPathLength, f.enoughBounces = self.maxlen, self.enough
        f.view[:] = [float(x) for x in np.frombuffer(np.ascontiguousarray(view).tobytes(), np.float32)]
        f.threads = self.threads
        f.bsdfModel = self.bsdf
        keep += [cm, ct, ci, mats, texs]
        return f, keep

    def shade_paths(self, view, path_length, O4, D4, T4, hits, R0, shift, pass_):
        """Stage-level oracle: shadeKernel on n paths. Returns dict of per-path (uncompacted) outputs + flags."""
        f, keep = self.frame_in(view)
        n = O4.shape[0]
        arrs = {k: np.zeros((n, 4), np.float32) for k in ("extO", "extD", "extT", "shO", "shD", "shE", "deposit")}
        flags = np.zeros(n, np.uint8)
        O4, D4, T4 = (np.ascontiguousarray(a, np.float32) for a in (O4, D4, T4))
        hits = np.ascontiguousarray(hits, np.uint32)
        p = lambda a: ctypes.c_void_p(a.ctypes.data)
        lib().orc_shade_paths(ctypes.byref(f), path_length, n, p(O4), p(D4), p(T4), p(hits), ctypes.c_uint(R0), ctypes.c_uint(shift), pass_,
                              p(arrs["extO"]), p(arrs["extD"]), p(arrs["extT"]), p(arrs["shO"]), p(arrs["shD"]), p(arrs["shE"]),
                              p(arrs["deposit"]), p(flags))
        arrs["flags"] = flags
        return arrs

    def reference_tables(self, view):
        """Converted materials (CUDAMaterial, 128 B each), packed texel arrays, sky table and instance inverses - the
        device-side tables of the reference core, produced by the oracle's restatement of rendercore.cpp:438-565,716-737."""
        f, keep = self.frame_in(view)
        mats = np.zeros((len(self.sd.materials), 32), np.uint32)
        lib().orc_convert_materials(ctypes.byref(f), ctypes.c_void_p(mats.ctypes.data))
        cnt = [ctypes.c_uint(0) for _ in range(4)]
        lib().orc_scene_tables(ctypes.byref(f), None, ctypes.byref(cnt[0]), None, ctypes.byref(cnt[1]), None, ctypes.byref(cnt[2]), None,
                               ctypes.byref(cnt[3]), None)
        a32 = np.zeros((cnt[0].value, 4), np.uint8); a128 = np.zeros((cnt[1].value, 4), np.float32)
        n32 = np.zeros((cnt[2].value, 4), np.uint8); sky = np.zeros((cnt[3].value, 4), np.float32)
        inv = np.zeros((max(len(self.sd.instances), 1), 16), np.float32)
        p = lambda a: ctypes.c_void_p(a.ctypes.data)
        lib().orc_scene_tables(ctypes.byref(f), p(a32), ctypes.byref(cnt[0]), p(a128), ctypes.byref(cnt[1]), p(n32), ctypes.byref(cnt[2]),
                               p(sky), ctypes.byref(cnt[3]), p(inv))
        return dict(materials=mats, argb32=a32, argb128=a128, nrm32=n32, sky=sky, inverses=inv)


# ---- the reference's own shadeKernel, compiled for sm_100a from /root/reference (oracle/_ref) -------------

class RefShadeIn(ctypes.Structure):
    _fields_ = [("meshCount", ctypes.c_int), ("coreTris", ctypes.c_void_p), ("triCounts", ctypes.c_void_p),
                ("instanceCount", ctypes.c_int), ("instMesh", ctypes.c_void_p), ("instInverse16", ctypes.c_void_p),
                ("materials128", ctypes.c_void_p), ("materialCount", ctypes.c_int),
                ("triLights", ctypes.c_void_p), ("triLightCount", ctypes.c_int), ("pointLights", ctypes.c_void_p), ("pointLightCount", ctypes.c_int),
                ("spotLights", ctypes.c_void_p), ("spotLightCount", ctypes.c_int), ("dirLights", ctypes.c_void_p), ("dirLightCount", ctypes.c_int),
                ("argb32", ctypes.c_void_p), ("argb32Count", ctypes.c_int), ("argb128", ctypes.c_void_p), ("argb128Count", ctypes.c_int),
                ("nrm32", ctypes.c_void_p), ("nrm32Count", ctypes.c_int),
                ("skyPixels4", ctypes.c_void_p), ("skyPixelCount", ctypes.c_int), ("skyW", ctypes.c_int), ("skyH", ctypes.c_int),
                ("worldToSky", ctypes.c_float * 16), ("blueNoise", ctypes.c_void_p),
                ("geometryEpsilon", ctypes.c_float), ("clampValue", ctypes.c_float),
                ("pathCount", ctypes.c_int), ("stride", ctypes.c_int),
                ("pathStates", ctypes.c_void_p), ("hits", ctypes.c_void_p), ("connections", ctypes.c_void_p), ("accumulator", ctypes.c_void_p),
                ("R0", ctypes.c_uint), ("shift", ctypes.c_uint), ("pass_", ctypes.c_int), ("probePixelIdx", ctypes.c_int),
                ("pathLength", ctypes.c_int), ("w", ctypes.c_int), ("h", ctypes.c_int), ("spreadAngle", ctypes.c_float), ("useNEE", ctypes.c_int),
                ("countersOut", ctypes.c_uint * 12), ("timingRuns", ctypes.c_int), ("timingMsMin", ctypes.c_float), ("timingMsMean", ctypes.c_float)]


REF_SHADE_GPU = os.path.join(_HERE, "_ref", "libref_shade_gpu.so")                    # reference shadeKernel with lambert.h
REF_SHADE_DISNEY_GPU = os.path.join(_HERE, "_ref", "libref_shade_disney_gpu.so")      # ... with ggxmdf.h + frosted.h + disney.h


def have_ref_shade_gpu(bsdf=0):
    return os.path.exists(REF_SHADE_DISNEY_GPU if bsdf else REF_SHADE_GPU)


def ref_shade_gpu(oracle, view, path_length, O4, D4, T4, hits, R0, shift, pass_, accumulator, bsdf=0, timing_runs=0):
    """Run the REFERENCE shadeKernel (unmodified source, sm_100a build) on n paths. Returns compacted extension rays,
    shadow rays, the accumulator and the counters, like lh2b_shade_paths."""
    sd = oracle.sd
    rl = ctypes.CDLL(REF_SHADE_DISNEY_GPU if bsdf else REF_SHADE_GPU)
    tb = oracle.reference_tables(view)
    n = O4.shape[0]
    stride = 2 * n
    ps = np.zeros((3 * stride, 4), np.float32)
    ps[0:n], ps[stride:stride + n], ps[2 * stride:2 * stride + n] = O4, D4, T4
    hb = np.zeros((stride, 4), np.float32)
    hb[:n] = np.ascontiguousarray(hits).view(np.float32)
    conn = np.zeros((6 * stride, 4), np.float32)
    acc = np.ascontiguousarray(accumulator, np.float32).reshape(-1, 4).copy()
    tris = [np.ascontiguousarray(t) for _, t in sd.meshes]
    tp = (ctypes.c_void_p * len(tris))(*[t.ctypes.data for t in tris])
    tc = np.array([len(t) for t in tris], np.int32)
    im = np.array([m for m, _ in sd.instances], np.int32)
    bn8 = np.fromfile(_BLUENOISE, dtype=np.uint8)
    bn = np.zeros(65536 * 5 + 16, np.uint32)     # + 16 zero words: the sampler reads past the ranking tile for tile pixel (127, 127) at dimensions >= 8
    bn[:65536] = bn8[:65536]; bn[65536:65536 + 131072] = bn8[65536:65536 + 131072]; bn[3 * 65536:3 * 65536 + 131072] = bn8[65536 + 131072:]
    lights = [np.ascontiguousarray(a) for a in (sd.tri_lights, sd.point_lights, sd.spot_lights, sd.dir_lights)]
    r = RefShadeIn()
    p = lambda a: a.ctypes.data if a.size else None
    r.meshCount, r.coreTris, r.triCounts = len(tris), ctypes.addressof(tp), tc.ctypes.data
    r.instanceCount, r.instMesh, r.instInverse16 = len(sd.instances), im.ctypes.data, tb["inverses"].ctypes.data
    r.materials128, r.materialCount = tb["materials"].ctypes.data, len(sd.materials)
    r.triLights, r.triLightCount = p(lights[0]), len(lights[0])
    r.pointLights, r.pointLightCount = p(lights[1]), len(lights[1])
    r.spotLights, r.spotLightCount = p(lights[2]), len(lights[2])
    r.dirLights, r.dirLightCount = p(lights[3]), len(lights[3])
    r.argb32, r.argb32Count = tb["argb32"].ctypes.data, len(tb["argb32"])
    r.argb128, r.argb128Count = tb["argb128"].ctypes.data, len(tb["argb128"])
    r.nrm32, r.nrm32Count = tb["nrm32"].ctypes.data, len(tb["nrm32"])
    r.skyPixels4, r.skyPixelCount = tb["sky"].ctypes.data, len(tb["sky"])
    r.skyW, r.skyH = (sd.sky[1], sd.sky[2]) if sd.sky is not None else (0, 0)
    r.worldToSky[:] = [float(x) for x in np.eye(4, dtype=np.float32).flat]
    r.blueNoise = bn.ctypes.data
    r.geometryEpsilon, r.clampValue = oracle.eps, oracle.clamp
    r.pathCount, r.stride = n, stride
    r.pathStates, r.hits, r.connections, r.accumulator = ps.ctypes.data, hb.ctypes.data, conn.ctypes.data, acc.ctypes.data
    r.R0, r.shift, r.pass_, r.probePixelIdx, r.pathLength = R0, shift, pass_, -1, path_length
    r.w, r.h = oracle.w, oracle.h
    r.spreadAngle = float(np.ascontiguousarray(view)[0]["spreadAngle"])
    r.useNEE = int(sum(len(l) for l in lights) > 0)
    r.timingRuns = timing_runs
    rc = rl.refshade_run(ctypes.byref(r))
    if rc != 0:
        raise RuntimeError("refshade_run failed")
    cnt = list(r.countersOut)
    n_ext, n_sh = cnt[5] - n, cnt[6]          # Counters.extensionRays started at n (see ref_shade_gpu.cu), shadowRays at 0
    ext = dict(O=ps[n:n + n_ext], D=ps[stride + n:stride + n + n_ext], T=ps[2 * stride + n:2 * stride + n + n_ext])
    sh = dict(O=conn[0:n_sh], D=conn[2 * stride:2 * stride + n_sh], E=conn[4 * stride:4 * stride + n_sh])
    if timing_runs:
        return ext, sh, acc.reshape(oracle.h, oracle.w, 4), cnt, {"ms_min": float(r.timingMsMin), "ms_mean": float(r.timingMsMean)}
    return ext, sh, acc.reshape(oracle.h, oracle.w, 4), cnt


# ---- the reference's own SVGF / TAA kernels (oracle/_ref/libref_filter_gpu.so) ---------------------------

class FilterIO(ctypes.Structure):
    """Shared by reffilter_run (reference kernels) and lh2b_filter_chain (product): identical field order."""
    _fields_ = [("w", ctypes.c_int), ("h", ctypes.c_int), ("samplesTaken", ctypes.c_int), ("camIsStationary", ctypes.c_int), ("taa", ctypes.c_int),
                ("directClamp", ctypes.c_float), ("indirectClamp", ctypes.c_float), ("j0", ctypes.c_float), ("j1", ctypes.c_float),
                ("prevj0", ctypes.c_float), ("prevj1", ctypes.c_float), ("prevView", ctypes.c_float * 17),
                ("accumulator", ctypes.c_void_p), ("features", ctypes.c_void_p), ("worldPos", ctypes.c_void_p), ("prevWorldPos", ctypes.c_void_p),
                ("deltaDepth", ctypes.c_void_p), ("prevMoments", ctypes.c_void_p), ("filteredIN", ctypes.c_void_p), ("prevPixels", ctypes.c_void_p),
                ("featuresOut", ctypes.c_void_p), ("shadingAfterPrepare", ctypes.c_void_p), ("motion", ctypes.c_void_p), ("moments", ctypes.c_void_p),
                ("phase1", ctypes.c_void_p), ("phase2", ctypes.c_void_p), ("phase3", ctypes.c_void_p), ("taaPixels", ctypes.c_void_p), ("target", ctypes.c_void_p),
                ("timingRuns", ctypes.c_int), ("stageMs", ctypes.c_float * 8)]


REF_FILTER_GPU = os.path.join(_HERE, "_ref", "libref_filter_gpu.so")
FILTER_OUTPUTS = ("featuresOut", "shadingAfterPrepare", "motion", "moments", "phase1", "phase2", "phase3", "taaPixels", "target")


def have_ref_filter_gpu():
    return os.path.exists(REF_FILTER_GPU)


def make_filter_io(inputs, settings):
    """inputs: dict of numpy arrays (accumulator, features, worldPos, prevWorldPos, deltaDepth, prevMoments, filteredIN, prevPixels);
    settings: dict(w, h, samplesTaken, camIsStationary, taa, directClamp, indirectClamp, j0, j1, prevj0, prevj1, prevView (ViewPyramid array)).
    Returns (FilterIO, outputs dict, keepalive)."""
    io = FilterIO()
    w, h = settings["w"], settings["h"]
    for k in ("w", "h", "samplesTaken", "camIsStationary", "taa", "directClamp", "indirectClamp", "j0", "j1", "prevj0", "prevj1"):
        setattr(io, k, settings[k])
    io.prevView[:] = [float(x) for x in np.frombuffer(np.ascontiguousarray(settings["prevView"]).tobytes(), np.float32)]
    keep = {}
    for k in ("accumulator", "features", "worldPos", "prevWorldPos", "deltaDepth", "prevMoments", "filteredIN", "prevPixels"):
        keep[k] = np.ascontiguousarray(inputs[k])
        setattr(io, k, keep[k].ctypes.data)
    outs = {}
    for k in FILTER_OUTPUTS:
        shape = (h, w, 2) if k == "motion" else (h, w, 4)
        outs[k] = np.zeros(shape, np.uint32 if k == "featuresOut" else np.float32)
        setattr(io, k, outs[k].ctypes.data)
    return io, outs, keep


FILTER_STAGES = ("prepare", "atrous1", "atrous2", "atrous3", "taa", "present", "chain")


def ref_filter_gpu(inputs, settings, timing_runs=0):
    """The reference's own filter kernels on the inputs; with timing_runs also {stage: mean ms} of that many timed runs of the chain."""
    io, outs, keep = make_filter_io(inputs, settings)
    io.timingRuns = timing_runs
    rc = ctypes.CDLL(REF_FILTER_GPU).reffilter_run(ctypes.byref(io))
    if rc != 0:
        raise RuntimeError("reffilter_run failed")
    if timing_runs:
        return outs, dict(zip(FILTER_STAGES, (float(x) for x in io.stageMs)))
    return outs


CWBVH_REPORT = ("nodesVisited", "trisVisited", "maxDepth", "emptySlots", "leafSlots", "innerSlots", "errors", "firstError")


def cwbvh_check(nodes, tris, verts4):
    """Structural check of a CWBVH in the product's documented format (lh2_oracle_cwbvh.h); returns the report as a dict."""
    nodes = np.ascontiguousarray(nodes, np.uint8).reshape(-1, 80)
    tris = np.ascontiguousarray(tris, np.float32).reshape(-1, 12)
    v = np.ascontiguousarray(verts4, np.float32).reshape(-1, 4)
    rep = (ctypes.c_int * 8)()
    lib().orc_cwbvh_check(ctypes.c_void_p(nodes.ctypes.data), nodes.shape[0], ctypes.c_void_p(tris.ctypes.data), tris.shape[0],
                          ctypes.c_void_p(v.ctypes.data), v.shape[0] // 3, rep)
    return dict(zip(CWBVH_REPORT, [int(x) for x in rep]))


def cwbvh_closest_hits(nodes, tris, O4, D4, threads=None):
    """Closest hits through a CWBVH in the product's format, decoded and traversed by the oracle's own reader."""
    nodes = np.ascontiguousarray(nodes, np.uint8).reshape(-1, 80)
    tris = np.ascontiguousarray(tris, np.float32).reshape(-1, 12)
    O4 = np.ascontiguousarray(O4, np.float32).reshape(-1, 4)
    D4 = np.ascontiguousarray(D4, np.float32).reshape(-1, 4)
    hits = np.empty((O4.shape[0], 4), np.uint32)
    lib().orc_cwbvh_closest_hits(ctypes.c_void_p(nodes.ctypes.data), ctypes.c_void_p(tris.ctypes.data), ctypes.c_void_p(O4.ctypes.data),
                                 ctypes.c_void_p(D4.ctypes.data), O4.shape[0], ctypes.c_void_p(hits.ctypes.data), threads or os.cpu_count())
    return hits


def filter_chain_cpu(inputs, settings):
    """The oracle's CPU restatement of the SVGF / TAA chain (lh2_oracle_filter.h) on the same inputs / settings dictionaries as
    ref_filter_gpu; returns the same dictionary of outputs."""
    io, outs, keep = make_filter_io(inputs, settings)
    if lib().orc_filter_chain(ctypes.byref(io)) != 0:
        raise RuntimeError("orc_filter_chain failed")
    return outs


class FilteredFrameOracle:
    """CPU counterpart of the filtering core for whole frame sequences (BASELINE.json configs[4] at oracle sizes): FrameOracle in
    filter mode (feature writes, direct / indirect split: lib/RenderCore_Optix7Filter/kernels/pathtracer.h) followed by the CPU
    restatement of the SVGF / TAA chain (lh2_oracle_filter.h) with the buffer rotation of RenderCore::FinalizeRender
    (lib/RenderCore_Optix7Filter/rendercore.cpp:904-934: swap filteredIN / filteredOUT, shading / prevPixels, worldPos /
    prevWorldPos, moments / prevMoments). render() returns the presented frame (float32 [h, w, 4])."""

    def __init__(self, scene, width, height, epsilon=1e-3, clamp=10.0, max_path_length=3, taa=True, clamp_direct=15.0, clamp_indirect=15.0, threads=None):
        self.frame = FrameOracle(scene, width, height, 1, epsilon, clamp, max_path_length, 1, threads=threads, filter=True)
        self.w, self.h, self.taa = width, height, 1 if taa else 0
        self.clamp_direct, self.clamp_indirect = clamp_direct, clamp_indirect
        z = lambda: np.zeros((height, width, 4), np.float32)
        self.prev_world_pos, self.prev_moments, self.filtered_in, self.prev_pixels = z(), z(), z(), z()
        self.prev_view = None
        self.stages = None

    def render(self, view, converge=1):
        f = self.frame
        f.render(view, converge)
        view_bytes = np.frombuffer(np.ascontiguousarray(view).tobytes(), np.float32).copy()
        st = dict(w=self.w, h=self.h, samplesTaken=f.samples_taken, camIsStationary=0 if f.samples_taken == f.spp else 1, taa=self.taa,
                  directClamp=self.clamp_direct, indirectClamp=self.clamp_indirect, j0=0.0, j1=0.0, prevj0=0.0, prevj1=0.0,
                  prevView=self.prev_view if self.prev_view is not None else view_bytes)
        inputs = dict(accumulator=f.accum, features=f.features, worldPos=f.world_pos, prevWorldPos=self.prev_world_pos, deltaDepth=f.delta_depth,
                      prevMoments=self.prev_moments, filteredIN=self.filtered_in, prevPixels=self.prev_pixels)
        out = filter_chain_cpu(inputs, st)
        f.features[...] = out["featuresOut"]                       # the history counters prepare updated (shade keeps them next frame)
        self.prev_world_pos, self.prev_moments = f.world_pos.copy(), out["moments"]
        self.filtered_in = out["phase1"]                             # this frame's phase-1 output is the next frame's temporal history
        self.prev_pixels = out["taaPixels"] if self.taa else out["phase3"]
        self.prev_view, self.stages = view_bytes, out
        return out["target"]


def _bind_normals(tris):
    """float4 per vertex: vN0..2 (+ the N component riding in w) of every CoreTri, as the skinning code sees them."""
    t = np.ascontiguousarray(tris).view(np.float32).reshape(-1, 52)
    return np.ascontiguousarray(t[:, 8:20].reshape(-1, 4))


def skin_mesh(verts, tris, joints4, weights4, joint_matrices):
    """HostMesh::SetPose( HostSkin ) on the CPU (oracle/lh2_oracle_anim.h). Returns (verts, CoreTri array) of the posed mesh."""
    v0 = np.ascontiguousarray(verts, np.float32).reshape(-1, 4)
    out_v, out_t = v0.copy(), np.ascontiguousarray(tris).copy()
    bn = _bind_normals(tris)
    j, w = np.ascontiguousarray(joints4, np.uint32), np.ascontiguousarray(weights4, np.float32)
    m = np.ascontiguousarray(joint_matrices, np.float32).reshape(-1, 16)
    P = lambda a: ctypes.c_void_p(a.ctypes.data)
    lib().orc_skin_mesh(P(v0), P(bn), P(j), P(w), P(m), len(out_t), P(out_v), P(out_t))
    return out_v, out_t


def morph_mesh(verts, tris, deltas4, normals4, weights):
    """HostMesh::SetPose( weights ) on the CPU."""
    v0 = np.ascontiguousarray(verts, np.float32).reshape(-1, 4)
    out_v, out_t = v0.copy(), np.ascontiguousarray(tris).copy()
    bn = _bind_normals(tris)
    d, n, w = (np.ascontiguousarray(a, np.float32) for a in (deltas4, normals4, weights))
    P = lambda a: ctypes.c_void_p(a.ctypes.data)
    lib().orc_morph_mesh(P(v0), P(bn), P(d), P(n), P(w), len(w), len(out_t), P(out_v), P(out_t))
    return out_v, out_t


# ---- scenes recorded from the reference RenderSystem (oracle/ref_recorder_core.cpp) ------------------------------------------
def load_recording(path):
    """Parses the file oracle/_ref/libRenderCore_Recorder.so writes on Render: everything the reference RenderSystem sent through
    the CoreAPI_Base calls. Returns (SceneDesc, info) with info = {view (ViewPyramid record), width, height, spp, settings{name:
    value}, converge, probe}. Later calls supersede earlier ones the way a core would apply them."""
    import struct
    from lighthouse2_b200 import abi, scenes
    raw = open(path, "rb").read()
    pos, verts, tris, inst, texdescs, texels = 0, {}, {}, {}, None, {}
    sd, info = scenes.SceneDesc(), {"settings": {}, "probe": None}
    sky, skyxf = None, None
    while pos < len(raw):
        tag = raw[pos:pos + 16].split(b"\0")[0].decode()
        a, b, n = struct.unpack_from("<QQQ", raw, pos + 16)
        data = raw[pos + 40:pos + 40 + n]
        pos += 40 + n
        if tag == "verts":
            verts[a] = np.frombuffer(data, np.float32).reshape(-1, 4).copy()
        elif tag == "tris":
            tris[a] = np.frombuffer(data, abi.CoreTri).copy()
        elif tag == "instance":
            model = struct.unpack("<q", struct.pack("<Q", b))[0]
            if model == -1:
                inst = {k: v for k, v in inst.items() if k < a}
            else:
                inst[a] = (int(model), np.frombuffer(data, np.float32).reshape(4, 4).copy())
        elif tag == "materials":
            sd.materials = np.frombuffer(data, abi.CoreMaterial).copy()
        elif tag == "texdescs":
            texdescs, texels = np.frombuffer(data, abi.CoreTexDesc).copy(), {}
        elif tag == "texels":
            texels[a] = np.frombuffer(data, np.float32 if b == 1 else np.uint8).copy()
        elif tag in ("trilights", "pointlights", "spotlights", "dirlights"):
            dt = {"trilights": abi.CoreLightTri, "pointlights": abi.CorePointLight, "spotlights": abi.CoreSpotLight, "dirlights": abi.CoreDirectionalLight}[tag]
            setattr(sd, {"trilights": "tri_lights", "pointlights": "point_lights", "spotlights": "spot_lights", "dirlights": "dir_lights"}[tag],
                    np.frombuffer(data, dt).copy())
        elif tag == "sky":
            sky = (np.frombuffer(data, np.float32).reshape(int(b), int(a), 3).copy(), int(a), int(b))
        elif tag == "skyxform":
            skyxf = np.frombuffer(data, np.float32).reshape(4, 4).copy()
        elif tag == "target":
            info["width"], info["height"], info["spp"] = (int(x) for x in np.frombuffer(data, np.uint32))
        elif tag == "setting":
            info["settings"][data[:60].split(b"\0")[0].decode()] = float(np.frombuffer(data[60:64], np.float32)[0])
        elif tag == "view":
            info["view"], info["converge"] = np.frombuffer(data, abi.ViewPyramid).copy(), int(b)
        elif tag == "probe":
            info["probe"] = tuple(int(x) for x in np.frombuffer(data, np.int32))
    sd.meshes = [(verts[i], tris[i]) for i in sorted(verts)]
    sd.instances = [inst[i] for i in sorted(inst)]
    if sky is not None:
        sd.sky = sky if skyxf is None else sky + (skyxf,)
    if texdescs is not None:
        for i, d in enumerate(texdescs):
            sd.textures.append((texels[i], int(d["storage"]), int(d["width"]), int(d["height"]), int(d["MIPlevels"]), int(d["flags"])))
    return sd, info
