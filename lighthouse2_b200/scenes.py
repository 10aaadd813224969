"""Synthetic scenes and cameras for the configurations named in BASELINE.json.

Everything here produces the *inputs* the reference RenderSystem would hand to a core through
CoreAPI_Base: float4 positions, CoreTri records, CoreMaterial records, CoreLightTri records and a
ViewPyramid. The arithmetic of those host-side producers is restated from:
  Camera::GetView ................. lib/RenderSystem/camera.cpp:107-128
  mat4::LookAt .................... lib/RenderSystem/common_types.h:514-538
  HostTriLight::HostTriLight ...... lib/RenderSystem/host_light.cpp:25-42
  CoreTri fields .................. lib/RenderSystem/common_classes.h:57-97
"""
import math
import numpy as np

from . import abi

f32 = np.float32


def _normalize(v):
    v = np.asarray(v, dtype=np.float64)
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def view_pyramid(pos, target, fov_deg=40.0, width=1920, height=1080, focal_distance=5.0, aperture=0.0, distortion=0.0):
    """ViewPyramid as Camera::GetView builds it (p1 top-left, p2 top-right, p3 bottom-left of the focal plane)."""
    pos, target = np.asarray(pos, np.float64), np.asarray(target, np.float64)
    z = _normalize(target - pos)
    x = _normalize(np.cross(z, [0.0, 1.0, 0.0]))
    y = np.cross(x, z)
    aspect = width / height
    screen = math.tan(fov_deg / 2 / (180 / math.pi))
    c = pos + focal_distance * z
    v = np.zeros(1, dtype=abi.ViewPyramid)
    v["pos"] = pos
    v["p1"] = c - screen * x * focal_distance * aspect + screen * focal_distance * y
    v["p2"] = c + screen * x * focal_distance * aspect + screen * focal_distance * y
    v["p3"] = c - screen * x * focal_distance * aspect - screen * focal_distance * y
    v["aperture"], v["focalDistance"], v["distortion"] = aperture, focal_distance, distortion
    v["spreadAngle"] = (fov_deg * math.pi / 180) / height
    up1 = c - screen * x * aspect + screen * y
    up2 = c + screen * x * aspect + screen * y
    up3 = c - screen * x * aspect - screen * y
    v["imagePlane"] = np.linalg.norm(up1 - up2) * np.linalg.norm(up1 - up3)
    return v


def core_tris_from_verts(verts4, material=0, smooth_normals=None):
    """Fill CoreTri records for non-indexed float4 positions (3 per triangle)."""
    v = np.asarray(verts4, f32).reshape(-1, 3, 4)[:, :, :3]
    n = v.shape[0]
    t = np.zeros(n, dtype=abi.CoreTri)
    e1, e2 = v[:, 1] - v[:, 0], v[:, 2] - v[:, 0]
    N = np.cross(e1.astype(np.float64), e2.astype(np.float64))
    ln = np.linalg.norm(N, axis=1, keepdims=True)
    N = np.where(ln > 0, N / np.maximum(ln, 1e-30), [0.0, 1.0, 0.0])
    t["Nx"], t["Ny"], t["Nz"] = N[:, 0], N[:, 1], N[:, 2]
    if smooth_normals is None:
        t["vN0"] = t["vN1"] = t["vN2"] = N
    else:
        sn = np.asarray(smooth_normals, f32).reshape(-1, 3, 3)
        t["vN0"], t["vN1"], t["vN2"] = sn[:, 0], sn[:, 1], sn[:, 2]
    T = _normalize(np.where(np.linalg.norm(e1, axis=1, keepdims=True) > 0, e1, [1.0, 0.0, 0.0]))
    B = np.cross(N, T)
    t["T"], t["B"] = T, B
    a = np.linalg.norm(v[:, 1] - v[:, 0], axis=1).astype(f32)
    b = np.linalg.norm(v[:, 2] - v[:, 1], axis=1).astype(f32)
    c = np.linalg.norm(v[:, 0] - v[:, 2], axis=1).astype(f32)
    s = (a + b + c) * f32(0.5)
    area = np.sqrt(np.maximum(s * (s - a) * (s - b) * (s - c), 0)).astype(f32)
    t["area"] = area
    t["invArea"] = np.where(area > 0, 1.0 / np.maximum(area, 1e-30), 0)
    t["vertex0"], t["vertex1"], t["vertex2"] = v[:, 0], v[:, 1], v[:, 2]
    t["material"] = material
    t["ltriIdx"] = -1
    t["u"] = [0.0, 1.0, 0.0]
    t["v"] = [0.0, 0.0, 1.0]
    return t


def quad(center, normal, width, depth):
    """Two triangles forming a quad like HostScene::AddQuad (facing 'normal')."""
    n = _normalize(normal)
    helper = np.array([1.0, 0.0, 0.0]) if abs(n[0]) < 0.9 else np.array([0.0, 0.0, 1.0])
    t = _normalize(np.cross(n, helper)) * (width * 0.5)
    b = _normalize(np.cross(n, t)) * (depth * 0.5)
    c = np.asarray(center, np.float64)
    p = [c - t - b, c + t - b, c + t + b, c - t + b]
    tri = np.array([p[0], p[1], p[2], p[0], p[2], p[3]])
    # make the winding agree with the requested normal
    if np.dot(np.cross(tri[1] - tri[0], tri[2] - tri[0]), n) < 0:
        tri = tri[[0, 2, 1, 3, 5, 4]]
    v = np.zeros((6, 4), f32)
    v[:, :3] = tri
    return v


def tri_lights(verts4, tris, materials, inst_idx=0, transform=None, first_ltri=0):
    """CoreLightTri records for the emissive triangles of a mesh instance; also sets tris['ltriIdx'].
    A material is emissive when a colour channel exceeds 1 (HostMaterial::IsEmissive)."""
    v = np.asarray(verts4, f32).reshape(-1, 3, 4)[:, :, :3].astype(np.float64)
    lights = []
    for i in range(v.shape[0]):
        col = materials[int(tris[i]["material"])]["color"]["value"]
        if not (col > 1.0).any():
            continue
        p = v[i]
        if transform is not None:
            m = np.asarray(transform, np.float64).reshape(4, 4)
            p = p @ m[:3, :3].T + m[:3, 3]
        l = np.zeros(1, dtype=abi.CoreLightTri)
        p32 = p.astype(f32)
        l["vertex0"], l["vertex1"], l["vertex2"] = p32[0], p32[1], p32[2]
        l["centre"] = f32(0.333333) * (p32[0] + p32[1] + p32[2])
        n = np.cross(p[1] - p[0], p[2] - p[0])
        l["N"] = n / np.linalg.norm(n)
        a = f32(np.linalg.norm(p32[1] - p32[0])); b = f32(np.linalg.norm(p32[2] - p32[1])); c = f32(np.linalg.norm(p32[0] - p32[2]))
        s = (a + b + c) * f32(0.5)
        area = f32(math.sqrt(max(float(s * (s - a) * (s - b) * (s - c)), 0.0)))
        l["area"], l["radiance"] = area, col
        l["energy"] = float((col * area).sum())
        l["triIdx"], l["instIdx"] = i, inst_idx
        tris[i]["ltriIdx"] = first_ltri + len(lights)
        lights.append(l)
    return np.concatenate(lights) if lights else np.zeros(0, dtype=abi.CoreLightTri)


def terrain(nx=1000, nz=500, extent=50.0, seed=0x12345678, floaters=0):
    """Jittered-grid height field over [-extent, extent]^2 with 2 triangles per cell (nx * nz * 2 triangles),
    plus 'floaters' random small triangles above it. Returns float4[3 * T]."""
    rng = np.random.default_rng(seed)
    gx, gz = np.meshgrid(np.arange(nx + 1), np.arange(nz + 1), indexing="ij")
    jx = (rng.random(gx.shape) - 0.5) * 0.6
    jz = (rng.random(gx.shape) - 0.5) * 0.6
    jx[[0, -1], :] = 0; jz[:, [0, -1]] = 0
    x = ((gx + jx) / nx * 2 - 1) * extent
    z = ((gz + jz) / nz * 2 - 1) * extent
    y = (3.0 * np.sin(x * 0.21) * np.cos(z * 0.17) + 1.5 * np.sin(x * 0.63 + 1.3) * np.sin(z * 0.71)
         + 0.5 * np.sin(x * 1.9) * np.cos(z * 2.3) + 0.15 * rng.standard_normal(gx.shape))
    p = np.stack([x, y, z], axis=-1)
    p00, p10, p01, p11 = p[:-1, :-1], p[1:, :-1], p[:-1, 1:], p[1:, 1:]
    t1 = np.stack([p00, p01, p10], axis=2)  # upward facing
    t2 = np.stack([p10, p01, p11], axis=2)
    tris = np.concatenate([t1.reshape(-1, 3, 3), t2.reshape(-1, 3, 3)], axis=0)
    if floaters > 0:
        c = np.stack([(rng.random(floaters) * 2 - 1) * extent, 6 + rng.random(floaters) * 10, (rng.random(floaters) * 2 - 1) * extent], axis=1)
        d = (rng.random((floaters, 3, 3)) - 0.5) * 0.6
        tris = np.concatenate([tris, c[:, None, :] + d], axis=0)
    v = np.zeros((tris.shape[0] * 3, 4), f32)
    v[:, :3] = tris.reshape(-1, 3)
    return v


def random_soup(n, extent=10.0, size=1.0, seed=1):
    """n random triangles in a cube; the parity workhorse (lots of overlaps and near misses)."""
    rng = np.random.default_rng(seed)
    c = (rng.random((n, 1, 3)) * 2 - 1) * extent
    d = (rng.random((n, 3, 3)) - 0.5) * size
    v = np.zeros((n * 3, 4), f32)
    v[:, :3] = (c + d).reshape(-1, 3)
    return v


def camera_rays(view, width, height, sub=(0.5, 0.5)):
    """Pinhole primary rays through pixel centres (distortion 0, aperture 0): float4 O, float4 D."""
    v = view[0]
    pos, p1, p2, p3 = (np.asarray(v[k], np.float64) for k in ("pos", "p1", "p2", "p3"))
    right, up = p2 - p1, p3 - p1
    sx, sy = np.meshgrid(np.arange(width), np.arange(height), indexing="xy")
    u = (sx + sub[0]) / width
    w = (sy + sub[1]) / height
    target = p1 + u[..., None] * right + w[..., None] * up
    d = _normalize(target - pos).reshape(-1, 3)
    O = np.zeros((width * height, 4), f32); D = np.zeros((width * height, 4), f32)
    O[:, :3] = pos; D[:, :3] = d
    return O, D


def random_rays(n, extent=12.0, seed=2):
    rng = np.random.default_rng(seed)
    O = np.zeros((n, 4), f32); D = np.zeros((n, 4), f32)
    O[:, :3] = (rng.random((n, 3)) * 2 - 1) * extent
    t = (rng.random((n, 3)) * 2 - 1) * extent * 0.5
    D[:, :3] = _normalize(t - O[:, :3])
    return O, D


# ---- scene descriptions -------------------------------------------------------------------------

class SceneDesc:
    """Plain container of everything a core (or the oracle) is fed through the CoreAPI_Base calls."""

    def __init__(self):
        self.meshes = []        # list of (verts4 float32[3T,4], CoreTri[T])
        self.instances = []     # list of (meshIdx, 4x4 float32 or None)
        self.materials = None   # CoreMaterial[]
        self.tri_lights = np.zeros(0, dtype=abi.CoreLightTri)
        self.point_lights = np.zeros(0, dtype=abi.CorePointLight)
        self.spot_lights = np.zeros(0, dtype=abi.CoreSpotLight)
        self.dir_lights = np.zeros(0, dtype=abi.CoreDirectionalLight)
        self.sky = None         # (float32[h,w,3], w, h)
        self.textures = []      # list of (texels, storage, w, h, mips)

    def upload(self, core):
        """Issue the calls in the order RenderSystem::SynchronizeSceneData does (rendersystem.cpp:203-211)."""
        if self.sky is not None:
            core.SetSkyData(*self.sky)      # (pixels, w, h[, worldToLight])
        if self.textures:
            core.SetTextures(self.textures)
        core.SetMaterials(self.materials)
        for i, (v, t) in enumerate(self.meshes):
            core.SetGeometry(i, v, t)
        core.SetLights(self.tri_lights, self.point_lights, self.spot_lights, self.dir_lights)
        for i, (m, xf) in enumerate(self.instances):
            core.SetInstance(i, m, xf)
        core.SetInstance(len(self.instances), -1)
        core.FinalizeInstances()


def gradient_sky(w=512, h=256, zenith=(0.35, 0.5, 0.9), horizon=(0.9, 0.9, 0.85), scale=1.0):
    """Equirect sky with a vertical gradient; rows follow theta = acos(D.z) like SampleSkydome (tools_shared.h:196-203)."""
    t = np.linspace(0, 1, h, dtype=np.float32)[:, None, None]
    g = 1.0 - np.abs(2 * t - 1)     # brightest at the middle rows (theta = pi/2)
    sky = (np.asarray(zenith, f32) * (1 - g) + np.asarray(horizon, f32) * g) * scale
    return np.ascontiguousarray(np.broadcast_to(sky, (h, w, 3)).astype(f32)), w, h


def make_materials(specs):
    """specs: list of dicts with optional keys color, roughness, transmission, eta, absorption, smooth and the other
    principled-model scalars (metallic, subsurface, specular, specularTint, anisotropic, sheen, sheenTint, clearcoat,
    clearcoatGloss), all in [0, 1] as HostMaterial stores them."""
    m = abi.default_material(len(specs))
    for i, s in enumerate(specs):
        m[i]["color"]["value"] = s.get("color", (0.8, 0.8, 0.8))
        m[i]["roughness"]["value"] = s.get("roughness", 1.0)
        m[i]["transmission"]["value"] = s.get("transmission", 0.0)
        m[i]["eta"]["value"] = s.get("eta", 1.0)
        m[i]["absorption"]["value"] = s.get("absorption", (0.0, 0.0, 0.0))
        for k in ("metallic", "subsurface", "specular", "specularTint", "anisotropic", "sheen", "sheenTint", "clearcoat", "clearcoatGloss"):
            if k in s:
                m[i][k]["value"] = s[k]
        m[i]["flags"] = 1 if s.get("smooth", False) else 0
    return m


def principled_specs(n, seed=7):
    """n material specs that exercise every lobe of the principled model: rough / glossy dielectrics, metals, sheen cloth,
    clearcoat, anisotropic brushed metal, subsurface, frosted glass."""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        col = tuple(0.25 + 0.7 * rng.random(3))
        kind = i % 8
        s = dict(color=col, roughness=float(0.15 + 0.8 * rng.random()), specular=float(rng.random()), specularTint=float(rng.random()))
        if kind == 1:
            s.update(metallic=1.0, roughness=float(0.1 + 0.5 * rng.random()))
        elif kind == 2:
            s.update(sheen=float(0.5 + 0.5 * rng.random()), sheenTint=float(rng.random()), roughness=1.0)
        elif kind == 3:
            s.update(clearcoat=float(0.5 + 0.5 * rng.random()), clearcoatGloss=float(rng.random()))
        elif kind == 4:
            s.update(metallic=float(0.5 + 0.5 * rng.random()), anisotropic=float(0.3 + 0.7 * rng.random()), roughness=float(0.2 + 0.4 * rng.random()))
        elif kind == 5:
            s.update(subsurface=float(0.3 + 0.7 * rng.random()))
        elif kind == 6:
            s.update(transmission=1.0, eta=float(1.0 / (1.3 + 0.3 * rng.random())), roughness=float(0.05 + 0.4 * rng.random()), absorption=(0.1, 0.3, 0.2))
        elif kind == 7:
            s.update(metallic=float(rng.random()), sheen=float(rng.random()), clearcoat=float(rng.random()), clearcoatGloss=float(rng.random()),
                     anisotropic=float(rng.random()), subsurface=float(rng.random()), transmission=float(0.3 * rng.random()), eta=1.0 / 1.5)
        out.append(s)
    return out


def config2_scene(nx=1000, nz=500, n_materials=1, light_quads=1, seed=0x12345678, floaters=0, material_specs=None):
    """BASELINE.json configs[1]/[2]: height-field terrain + emissive quads.
    n_materials > 1 assigns diffuse/specular materials with roughness in {0, 0.3, 1} to terrain patches."""
    sd = SceneDesc()
    rng = np.random.default_rng(seed ^ 0x9E3779B9)
    specs = []
    if material_specs is not None:
        n_materials = len(material_specs)
    for i in range(max(1, n_materials)):
        col = 0.35 + 0.6 * rng.random(3)
        rough = (1.0, 0.3, 0.0)[i % 3] if n_materials > 1 else 1.0
        specs.append(dict(color=tuple(col), roughness=rough))
    if material_specs is not None:
        specs = [dict(s) for s in material_specs]
    light_mat = len(specs)
    specs.append(dict(color=(100.0, 100.0, 80.0)))
    sd.materials = make_materials(specs)
    tv = terrain(nx, nz, 50.0, seed, floaters)
    tt = core_tris_from_verts(tv)
    if n_materials > 1:
        c = tv.reshape(-1, 3, 4)[:, :, :3].mean(axis=1)
        patch = (np.floor((c[:, 0] + 50) / 12.5).astype(np.int64) * 8 + np.floor((c[:, 2] + 50) / 12.5).astype(np.int64))
        tt["material"] = (patch * 2654435761 % (2 ** 32) >> 8) % n_materials
    sd.meshes.append((tv, tt))
    sd.instances.append((0, None))
    lights = []
    for k in range(light_quads):
        if light_quads == 1:
            center = (0.0, 26.0, 0.0)
        else:
            a = 2 * math.pi * k / light_quads
            center = (30.0 * math.cos(a), 22.0 + 3.0 * (k % 3), 30.0 * math.sin(a))
        qv = quad(center, (0, -1, 0), 6.9, 6.9)
        qt = core_tris_from_verts(qv, material=light_mat)
        mesh_idx = len(sd.meshes)
        inst_idx = len(sd.instances)
        lt = tri_lights(qv, qt, sd.materials, inst_idx=inst_idx, first_ltri=sum(len(l) for l in lights))
        lights.append(lt)
        sd.meshes.append((qv, qt))
        sd.instances.append((mesh_idx, None))
    sd.tri_lights = np.concatenate(lights)
    sd.sky = gradient_sky()
    return sd
