"""Generates tests/golden/shade_reference_vectors.npz ON THE GPU BOX: outputs of the REFERENCE's own shadeKernel
(unmodified source, compiled for sm_100a into oracle/_ref/libref_shade_gpu.so by oracle/Makefile) for fixed, seeded
inputs at path lengths 1..3. Inputs are produced without the product: primary rays from the camera model, hit records
from the brute-force CPU oracle; path length L+1 consumes the reference kernel's own extension rays.
  gpurun -- 'python tools/make_golden_shade.py gpurun_out/shade_reference_vectors.npz'   then copy into tests/golden/.
  gpurun -- 'python tools/make_golden_shade.py gpurun_out/shade_disney_reference_vectors.npz 1'   the same with the reference kernel
            compiled against disney.h / ggxmdf.h / frosted.h (oracle/_ref/libref_shade_disney_gpu.so) and 16 principled materials.
The CPU test tests/test_oracle_golden.py replays the inputs through the oracle's ShadeStep and compares."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lighthouse2_b200 import scenes, abi
from oracle import binding as orc

W, H = 96, 54


def golden_scene(bsdf=0):
    sd = scenes.config2_scene(40, 28, n_materials=6, light_quads=2, floaters=200, seed=0xC0FFEE,
                              material_specs=scenes.principled_specs(16, seed=11) if bsdf else None)
    pl = np.zeros(1, abi.CorePointLight); pl["position"] = (12, 14, -6); pl["radiance"] = (260, 240, 200); pl["energy"] = 700
    sl = np.zeros(1, abi.CoreSpotLight); sl["position"] = (-10, 24, -20); sl["direction"] = (0.2, -0.8, 0.566); sl["radiance"] = (800, 800, 650)
    sl["cosInner"], sl["cosOuter"] = 0.95, 0.8
    dl = np.zeros(1, abi.CoreDirectionalLight); dl["direction"] = (0.3, -0.9, 0.316); dl["radiance"] = (1.2, 1.1, 1.0); dl["energy"] = 3.3
    sd.point_lights, sd.spot_lights, sd.dir_lights = pl, sl, dl
    if not bsdf:
        # one mirror-like and one glass-like material so specular / transmission branches are exercised
        sd.materials[1]["roughness"]["value"] = 0.0
        sd.materials[2]["transmission"]["value"] = 0.9; sd.materials[2]["eta"]["value"] = 1.0 / 1.5
        sd.materials[2]["absorption"]["value"] = (0.2, 0.1, 0.05)
    view = scenes.view_pyramid((4, 16, -62), (0, 1, 0), 45, W, H)
    return sd, view


def primary_state(view):
    O, D = scenes.camera_rays(view, W, H, sub=(0.37, 0.61))
    n = W * H
    O[:, 3] = ((np.arange(n, dtype=np.uint32) << 6) | 1).view(np.float32)
    return O, D, np.ones((n, 4), np.float32)


if __name__ == "__main__":
    out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/shade_reference_vectors.npz"
    bsdf = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    sd, view = golden_scene(bsdf)
    oracle = orc.FrameOracle(sd, W, H, 1, 1e-3, 10.0, 3, 1, bsdf=bsdf)
    meshes = [m for m, _ in sd.meshes]
    O4, D4, T4 = primary_state(view)
    shift, pass_ = 0x2F6B1A55, 0
    store = {"scene_checksum": np.array([float(np.sum(meshes[0][:, :3].astype(np.float64)))]), "shift": np.array([shift], np.uint32)}
    for L in (1, 2, 3):
        hits = orc.closest_hits(meshes, sd.instances, O4, D4)
        R0 = (0x85EBCA6B * L + L * 91771) & 0xFFFFFFFF
        acc0 = np.zeros((H, W, 4), np.float32)
        ext, sh, acc, cnt = orc.ref_shade_gpu(oracle, view, L, O4, D4, T4, hits, R0, shift, pass_, acc0, bsdf)
        for k, v in (("O", O4), ("D", D4), ("T", T4), ("hits", hits), ("extO", ext["O"]), ("extD", ext["D"]), ("extT", ext["T"]),
                     ("shO", sh["O"]), ("shD", sh["D"]), ("shE", sh["E"]), ("acc", acc)):
            store[f"L{L}_{k}"] = np.ascontiguousarray(v).copy()
        store[f"L{L}_R0"] = np.array([R0], np.uint32)
        print(f"L{L}: {len(O4)} paths -> {len(ext['O'])} extension rays, {len(sh['O'])} shadow rays")
        # next level: the reference kernel's own extension rays, sorted by path index for a stable order
        order = np.argsort(ext["O"][:, 3].view(np.uint32) >> 6, kind="stable")
        O4, D4, T4 = ext["O"][order].copy(), ext["D"][order].copy(), ext["T"][order].copy()
        if len(O4) == 0:
            break
    np.savez_compressed(out, **store)
    print("wrote", out, os.path.getsize(out), "bytes")
