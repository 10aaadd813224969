"""Ours vs theirs on the same buffers (VERDICT r1 "missing" #1): the REFERENCE's own kernels, compiled for sm_100a from /root/reference
into oracle/_ref (test infrastructure), timed beside the product's kernels with CUDA events around the kernel launches only.

  shade   lib/rendercore_optix7/kernels/pathtracer.h:54-252 (grid ceil(n/128), block 128) vs csrc/shade_kernels.cu, on the
          path states of a 1920x1080 frame of the C3 scene (64 materials, 8 emissive quads) at path lengths 1, 2, 3:
          ms per launch, paths/s, GB/s at SURVEY 8(d)'s 320 B/path
  filter  lib/CUDA/shared_kernel_code/finalize_shared.h:217-600 with the launch shapes of its host wrappers
          (lib/RenderCore_Optix7Filter/rendercore.cpp:897-948) vs csrc/filter_kernels.cu, on 3840x2160 g-buffers of the same
          scene from two camera positions: ms per stage

Usage (GPU box): python tools/ref_kernel_timing.py [out.json] [shade,filter] [width height]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from lighthouse2_b200 import RenderCore, scenes
from oracle import binding as orc

out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ref_kernels.json"
which = (sys.argv[2] if len(sys.argv) > 2 else "shade,filter").split(",")
res = {}
RUNS = 10


def shade_timing(W, H, nx=1000, nz=500):
    sd = scenes.config2_scene(nx, nz, n_materials=64, light_quads=8)
    core = RenderCore(0)
    core.SetTarget(W, H, 1)
    core.Setting("epsilon", 1e-3)
    sd.upload(core)
    view = scenes.view_pyramid((0, 30, -80), (0, 0, 0), 40, W, H)
    core.Render(view, 1)                                  # fixes the view-dependent constants of the hooks
    oracle = orc.FrameOracle(sd, W, H, 1, 1e-3, 10.0, 3, 1)
    O, D = scenes.camera_rays(view, W, H)
    n = W * H
    O4, D4 = O.copy(), D.copy()
    O4[:, 3] = ((np.arange(n, dtype=np.uint32) << 6) | 1).view(np.float32)
    T4 = np.ones((n, 4), np.float32)
    out = []
    for L in (1, 2, 3):
        n = O4.shape[0]
        hits = core.TraceRays(O4, D4)
        R0, shift = (0x9E3779B9 * L + L * 91771) & 0xFFFFFFFF, 0x5A17C3E1
        ours_min, ours_mean = core.ShadePathsTime(L, O4, D4, T4, hits, R0, shift, 0, RUNS)
        acc0 = np.zeros((H, W, 4), np.float32)
        ext, sh, _ = core.ShadePaths(L, O4, D4, T4, hits, R0, shift, 0, acc0)
        row = {"path_length": L, "paths": n, "extension_rays_out": len(ext["O"]), "shadow_rays_out": len(sh["O"]),
               "ours_ms": ours_min, "ours_ms_mean": ours_mean, "ours_gpaths_per_s": n / ours_min / 1e6, "ours_gb_per_s_at_320_B_per_path": n * 320 / ours_min / 1e6}
        if orc.have_ref_shade_gpu(0):
            rext, rsh, _, _, tm = orc.ref_shade_gpu(oracle, view, L, O4, D4, T4, hits, R0, shift, 0, acc0, 0, timing_runs=RUNS)
            row.update({"reference_ms": tm["ms_min"], "reference_ms_mean": tm["ms_mean"], "reference_gpaths_per_s": n / tm["ms_min"] / 1e6,
                        "reference_gb_per_s_at_320_B_per_path": n * 320 / tm["ms_min"] / 1e6, "speedup": tm["ms_min"] / ours_min,
                        "reference_extension_rays_out": len(rext["O"]), "reference_shadow_rays_out": len(rsh["O"])})
        out.append(row)
        print("shade", json.dumps(row), flush=True)
        O4, D4, T4 = ext["O"], ext["D"], ext["T"]
        if len(O4) == 0:
            break
    core.Shutdown()
    return {"resolution": [W, H], "scene": "C3 scene: %d triangles, 64 materials, 8 emissive quads" % (2 * nx * nz + 16), "runs": RUNS,
            "reference_launch": "grid ceil(n/128) x block 128 (pathtracer.h:244-252), unmodified kernel, -use_fast_math, sm_100a", "levels": out}


def filter_timing(W, H, nx=1000, nz=500):
    import test_filter_gpu as t
    t.W, t.H = W, H
    rng = np.random.default_rng(77)
    sd = scenes.config2_scene(nx, nz, n_materials=64, light_quads=8)
    core = RenderCore(0)
    for i, (v, tr) in enumerate(sd.meshes):
        core.SetGeometry(i, v, tr)
    for i, (m, xf) in enumerate(sd.instances):
        core.SetInstance(i, m, xf)
    core.SetInstance(len(sd.instances), -1)
    core.FinalizeInstances()
    prev_view = scenes.view_pyramid((0.0, 30, -80), (0, 0, 0), 40, W, H)
    view = scenes.view_pyramid((0.2, 30.0, -79.9), (0, 0, 0), 40, W, H)         # the camera step of the C5 run
    feat, wp, dd, albedo = t.gbuffer(core, sd, view, rng, spec_mat=1)
    _, pwp, _, _ = t.gbuffer(core, sd, prev_view, rng, spec_mat=1)
    smooth = (0.4 + 0.3 * np.sin(np.linspace(0, 9, W))[None, :, None] * np.cos(np.linspace(0, 7, H))[:, None, None] + 0 * albedo).astype(np.float32)
    direct = albedo * (smooth + 0.5 * rng.random((H, W, 1)).astype(np.float32))
    indirect = albedo * (0.3 * rng.random((H, W, 3)).astype(np.float32))
    acc = np.zeros((2, H, W, 4), np.float32); acc[0, ..., :3] = direct; acc[1, ..., :3] = indirect
    pm = np.zeros((H, W, 4), np.float32)
    pm[..., 0] = 0.5 + 0.1 * rng.random((H, W)); pm[..., 1] = pm[..., 0] ** 2 + 0.02 * rng.random((H, W))
    pm[..., 2] = 0.2 + 0.1 * rng.random((H, W)); pm[..., 3] = pm[..., 2] ** 2 + 0.01 * rng.random((H, W))
    fin = t.combine(smooth + 0.05 * rng.random((H, W, 3)).astype(np.float32), 0.15 + 0.05 * rng.random((H, W, 3)).astype(np.float32))
    pp = np.zeros((H, W, 4), np.float32); pp[..., :3] = np.sqrt(albedo * 0.6) + 0.02 * rng.random((H, W, 3)).astype(np.float32)
    inputs = dict(accumulator=acc, features=feat, worldPos=wp, prevWorldPos=pwp, deltaDepth=dd, prevMoments=pm, filteredIN=fin, prevPixels=pp)
    out = {"resolution": [W, H], "runs": RUNS, "cases": []}
    for taa, stationary, name in ((1, 0, "moving camera, TAA on (the C5 configuration)"), (0, 1, "stationary camera, TAA off")):
        st = dict(w=W, h=H, samplesTaken=1, camIsStationary=stationary, taa=taa, directClamp=15.0, indirectClamp=15.0, j0=0.0, j1=0.0, prevj0=0.0, prevj1=0.0,
                  prevView=prev_view)
        io, got, keep = orc.make_filter_io(inputs, st)
        io.timingRuns = RUNS
        core.FilterChain(io)
        ours = dict(zip(orc.FILTER_STAGES, (float(x) for x in io.stageMs)))
        row = {"case": name, "ours_ms": ours, "ours_gb_per_s_at_584_B_per_px": W * H * 584 / ours["chain"] / 1e6}
        if orc.have_ref_filter_gpu():
            _, theirs = orc.ref_filter_gpu(inputs, st, timing_runs=RUNS)
            row.update({"reference_ms": theirs, "speedup_chain": theirs["chain"] / ours["chain"],
                        "speedup_per_stage": {k: (theirs[k] / ours[k] if ours[k] > 0 else None) for k in orc.FILTER_STAGES}})
        out["cases"].append(row)
        print("filter", json.dumps(row), flush=True)
    core.Shutdown()
    return out


if __name__ == "__main__":
    if "shade" in which:
        W, H = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (1920, 1080)
        res["shade"] = shade_timing(W, H)
    if "filter" in which:
        W, H = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (3840, 2160)
        res["filter"] = filter_timing(W, H)
    os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
    json.dump(res, open(out_path, "w"), indent=1)
    print("wrote", out_path)
