"""numpy views of the reference's core-facing PODs (layouts: include/lh2_core_api.h, which cites
lib/RenderSystem/common_classes.h line by line). Sizes/offsets are checked by tests/test_abi.py."""
import numpy as np

f32, i32, u32 = np.float32, np.int32, np.uint32

CoreTri = np.dtype([
    ("u", f32, 3), ("ltriIdx", i32), ("v", f32, 3), ("material", u32),
    ("vN0", f32, 3), ("Nx", f32), ("vN1", f32, 3), ("Ny", f32), ("vN2", f32, 3), ("Nz", f32),
    ("T", f32, 3), ("area", f32), ("B", f32, 3), ("invArea", f32), ("alpha", f32, 3), ("LOD", f32),
    ("vertex0", f32, 3), ("dummy0", f32), ("vertex1", f32, 3), ("dummy1", f32), ("vertex2", f32, 3), ("dummy2", f32),
    ("u1", f32, 3), ("dummy3", f32), ("v1", f32, 3), ("dummy4", f32)])
assert CoreTri.itemsize == 208

_V3 = [("value", f32, 3), ("textureID", i32), ("scale", f32), ("_p", f32), ("uvscale", f32, 2), ("uvoffset", f32, 2), ("size", u32, 2)]
_SC = [("value", f32), ("textureID", i32), ("component", i32), ("scale", f32), ("uvscale", f32, 2), ("uvoffset", f32, 2), ("size", u32, 2)]
Vec3Value = np.dtype(_V3)
ScalarValue = np.dtype(_SC)
assert Vec3Value.itemsize == 48 and ScalarValue.itemsize == 40


def _material_dtype():
    names, formats, offsets = [], [], []
    off = 0

    def add(name, dt, align=8):
        nonlocal off
        off = (off + align - 1) // align * align
        names.append(name), formats.append(dt), offsets.append(off)
        off += np.dtype(dt).itemsize

    for n in ("color", "detailColor", "normals", "detailNormals"):
        add(n, Vec3Value)
    add("flags", u32, 4)
    add("absorption", Vec3Value)
    for n in ("metallic", "subsurface", "specular", "roughness", "specularTint", "anisotropic", "sheen", "sheenTint",
              "clearcoat", "clearcoatGloss", "transmission", "eta", "reflection", "refraction", "ior"):
        add(n, ScalarValue)
    add("pbrtMaterialType", np.int8, 1)
    add("urough", ScalarValue), add("vrough", ScalarValue)
    add("Ks", Vec3Value), add("eta_rgb", Vec3Value)
    add("sigma", ScalarValue)
    add("thin", np.uint8, 1)
    add("specTrans", ScalarValue), add("diffTrans", ScalarValue)
    add("scatterDistance", Vec3Value)
    add("flatness", ScalarValue)
    add("Kr", Vec3Value), add("opacity", Vec3Value)
    return np.dtype({"names": names, "formats": formats, "offsets": offsets, "itemsize": 1344})


CoreMaterial = _material_dtype()
assert CoreMaterial.itemsize == 1344

CoreLightTri = np.dtype([("centre", f32, 3), ("energy", f32), ("N", f32, 3), ("area", f32), ("radiance", f32, 3), ("dummy2", i32),
                         ("vertex0", f32, 3), ("triIdx", i32), ("vertex1", f32, 3), ("instIdx", i32), ("vertex2", f32, 3), ("dummy1", i32)])
CorePointLight = np.dtype([("position", f32, 3), ("energy", f32), ("radiance", f32, 3), ("dummy", i32)])
CoreSpotLight = np.dtype([("position", f32, 3), ("cosInner", f32), ("radiance", f32, 3), ("cosOuter", f32), ("direction", f32, 3), ("dummy", i32)])
CoreDirectionalLight = np.dtype([("direction", f32, 3), ("energy", f32), ("radiance", f32, 3), ("dummy", i32)])
assert (CoreLightTri.itemsize, CorePointLight.itemsize, CoreSpotLight.itemsize, CoreDirectionalLight.itemsize) == (96, 32, 48, 32)

ViewPyramid = np.dtype([("pos", f32, 3), ("p1", f32, 3), ("p2", f32, 3), ("p3", f32, 3), ("aperture", f32), ("spreadAngle", f32),
                        ("imagePlane", f32), ("focalDistance", f32), ("distortion", f32)])
assert ViewPyramid.itemsize == 68

CoreStats = np.dtype([("deviceName", np.uint64), ("SMcount", u32), ("ccMajor", u32), ("ccMinor", u32), ("VRAM", u32),
                      ("argb32TexelCount", u32), ("argb128TexelCount", u32), ("nrm32TexelCount", u32), ("bvhBuildTime", f32),
                      ("totalRays", u32), ("totalExtensionRays", u32), ("totalShadowRays", u32), ("renderTime", f32),
                      ("frameOverhead", f32), ("primaryRayCount", u32), ("traceTime0", f32), ("bounce1RayCount", u32),
                      ("traceTime1", f32), ("deepRayCount", u32), ("traceTimeX", f32), ("shadowTraceTime", f32),
                      ("shadeTime", f32), ("filterTime", f32), ("probedInstid", i32), ("probedTriid", i32), ("probedDist", f32),
                      ("probedWorldPos", f32, 3)], align=False)
assert CoreStats.itemsize == 120

CoreTexDesc = np.dtype([("data", np.uint64), ("width", u32), ("height", u32), ("flags", u32), ("pixelCount", u32),
                        ("firstPixel", u32), ("MIPlevels", u32), ("storage", i32), ("_pad", u32)])
assert CoreTexDesc.itemsize == 40

FrameStats = np.dtype([("generateExtendMs", f32), ("extendMs", f32), ("shadeMs", f32), ("connectMs", f32), ("finalizeMs", f32),
                       ("buildMs", f32), ("filterMs", f32), ("totalMs", f32), ("primaryRays", u32), ("extensionRays", u32),
                       ("shadowRays", u32), ("kernelLaunches", u32), ("pathLengthReached", u32), ("reserved", u32, 3)])
BvhStats = np.dtype([("nodes", u32), ("triangles", u32), ("bytes", u32), ("buildMs", f32), ("sahCost", f32), ("reserved", u32, 3)])


def default_material(n=1):
    """CoreMaterial array with the reference HostMaterial defaults that matter to the core
    (lib/RenderSystem/host_material.h: color 1, roughness 1... the core only reads the fields converted at
    lib/rendercore_optix7/rendercore.cpp:526-548)."""
    m = np.zeros(n, dtype=CoreMaterial)
    for name in ("color", "detailColor", "normals", "detailNormals", "absorption", "Ks", "eta_rgb", "scatterDistance", "Kr", "opacity"):
        m[name]["textureID"] = -1
    for name in ("metallic", "subsurface", "specular", "roughness", "specularTint", "anisotropic", "sheen", "sheenTint", "clearcoat",
                 "clearcoatGloss", "transmission", "eta", "reflection", "refraction", "ior", "urough", "vrough", "sigma", "specTrans",
                 "diffTrans", "flatness"):
        m[name]["textureID"] = -1
    m["color"]["value"] = 1.0
    m["roughness"]["value"] = 1.0
    m["eta"]["value"] = 1.0
    return m
