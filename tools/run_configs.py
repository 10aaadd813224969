"""Measures the BASELINE.json configurations that are not the bench.py workload (C1, C3, C4, C5) on one GPU
and writes a JSON summary. Usage: python tools/run_configs.py [out.json] [c1,c3,c4,c5]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lighthouse2_b200 import RenderCore, scenes

out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/configs.json"
which = (sys.argv[2] if len(sys.argv) > 2 else "c1,c3,c4,c5").split(",")
res = {}


def frame_stats(core, view, frames, converge=1):
    acc = {}
    for i in range(frames):
        core.Render(view, converge)
        fs = core.GetFrameStats()
        if i >= 2:
            for k in fs.dtype.names:
                if k != "reserved":
                    acc[k] = acc.get(k, 0.0) + float(fs[k]) / (frames - 2)
    return acc


def overrides(core):
    """experiments: LH2B_SET_<setting>=<value>"""
    for k, v in os.environ.items():
        if k.startswith("LH2B_SET_"):
            core.Setting(k[9:], float(v))


def counted_frame(core, view, converge=1):
    """one more frame with the traversal work counters on (lh2b_trace_stats): work per ray and SIMT utilisation of the two phases"""
    core.TraceStatsEnable(True)
    core.Render(view, converge)
    st = core.TraceStatsRead()
    core.TraceStatsEnable(False)
    r = max(1, st["rays"])
    return {"rays": st["rays"], "node_steps_per_ray": st["nodeSteps"] / r, "tri_tests_per_ray": st["triTests"] / r, "instance_entries_per_ray": st["instanceEntries"] / r,
            "iterations_x32_per_ray": st["iterations"] * 32 / r, "node_phase_lanes": st["nodeLanes"] / max(1, st["nodePhases"]),
            "tri_phase_lanes": st["triLanes"] / max(1, st["triPhases"])}


if "c1" in which:
    # C1: the literal tinyapp scene (pica glTF + light quad + legocar.obj, camera.xml) as the reference's own RenderSystem hands it
    # to a core (recorded with oracle/_ref/libRenderCore_Recorder.so through oracle/_ref/tinyapp_ref_host), 640x360, 1 spp, path
    # length 3: GPU core vs CPU oracle (BVH-pruned), same seeds. Falls back to a procedural stand-in when oracle/_ref is absent.
    from oracle import binding as orc
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    import test_rendersystem_dropin as rs
    W, H = 640, 360
    literal = os.path.exists(rs.HOST) and os.path.exists(rs.RECORDER)
    if literal:
        rs.run_host(rs.RECORDER, "/tmp/none.bin", frames=1, w=W, h=H, record="/tmp/c1.rec")
        sd, info = orc.load_recording("/tmp/c1.rec")
        view = info["view"]
        t0 = time.perf_counter(); host = rs.run_host(rs.CORE, "/tmp/c1_frame.bin", frames=20, w=W, h=H); host_s = time.perf_counter() - t0
    else:
        sd = scenes.config2_scene(40, 30, n_materials=5, light_quads=1, floaters=1500, seed=42)
        view = scenes.view_pyramid((-19.17, 9.19, 33.1), (-13.5, 7.99, 24.95), 40, W, H, focal_distance=5.0, aperture=1e-4, distortion=0.05)
    core = RenderCore(0); core.SetTarget(W, H, 1); core.Setting("epsilon", 1e-3); core.Setting("clampValue", 10); overrides(core); sd.upload(core)
    core.SetProbePos(W // 2, H // 2)
    a = frame_stats(core, view, 6)
    work = counted_frame(core, view)
    core.Render(view, 1); img = core.ReadPixels(); st = core.GetCoreStats()
    orc.set_accel(1)
    o = orc.FrameOracle(sd, W, H, 1, 1e-3, 10.0, 3, 1)
    o.render(view, 1)                                    # builds and caches the oracle's BVHs
    o = orc.FrameOracle(sd, W, H, 1, 1e-3, 10.0, 3, 1)
    t0 = time.perf_counter(); _, rec = o.render(view, 1, records=True); cpu_s = time.perf_counter() - t0
    # the oracle frame above is frame 1 of its own sequence; rebuild one in lock-step with the core's 8th Restart frame
    o2 = orc.FrameOracle(sd, W, H, 1, 1e-3, 10.0, 3, 1)
    for _ in range(8):
        want = o2.render(view, 1)
    orc.set_accel(0)
    frac, rr, energy = rs.frames_agree(img, want)
    probe = rec[W // 2 + (H // 2) * W]["hit"]
    res["c1"] = {"scene": "tinyapp (pica/scene.gltf + light quad + legocar.obj) via the reference RenderSystem" if literal else "procedural stand-in",
                 "resolution": [W, H], "meshes": len(sd.meshes), "instances": len(sd.instances), "triangles": int(sum(len(t) for _, t in sd.meshes)),
                 "stage_ms": {k: a[k] for k in ("generateExtendMs", "extendMs", "shadeMs", "connectMs", "finalizeMs")}, "traversal_work": work,
                 "gpu_ms_per_frame": a["totalMs"], "gpu_rays_per_frame": a["extensionRays"] + a["shadowRays"],
                 "gpu_mrays_per_s": (a["extensionRays"] + a["shadowRays"]) / a["totalMs"] / 1e3,
                 "flipped_pixel_fraction": float(frac), "rel_rmse_other_pixels": float(rr), "energy_difference": float(energy),
                 "probe_gpu": [int(st["probedInstid"]), int(st["probedTriid"]), float(st["probedDist"])],
                 "probe_oracle_frame1": [int(probe[1]), int(probe[2]), float(probe[3:4].view(np.float32)[0])],
                 "cpu_oracle_seconds": cpu_s, "cpu_oracle_mrays_per_s": sum(o.ray_counts) / cpu_s / 1e6, "cpu_threads": os.cpu_count(),
                 "cpu_oracle_kind": "scalar C++ port, BVH-pruned ray queries"}
    if literal:
        res["c1"]["through_reference_rendersystem"] = {"frames": 20, "wall_s_incl_scene_load": host_s, "core_render_ms_last_frame": host["renderTime"] * 1e3,
                                                       "rays_last_frame": host["totalRays"]}
    core.Shutdown()
    print("c1", res["c1"], flush=True)

if "c3" in which:
    # C3: 1M triangles, 64 diffuse/specular materials, 8 emissive quads, 1080p, 16 spp per Render, path length 8, NEE,
    # two diffuse bounces allowed (the stock Optix7 build stops after one; stated as the variant)
    W, H, SPP = 1920, 1080, 16
    sd = scenes.config2_scene(1000, 500, n_materials=64, light_quads=8)
    core = RenderCore(0); core.SetTarget(W, H, SPP); core.Setting("epsilon", 1e-3); core.Setting("maxPathLength", 8); core.Setting("maxDiffuseBounces", 2); overrides(core)
    sd.upload(core)
    view = scenes.view_pyramid((0, 30, -80), (0, 0, 0), 40, W, H)
    a = frame_stats(core, view, 5)
    work = counted_frame(core, view)
    res["c3"] = {"traversal_work": work, "resolution": [W, H], "spp": SPP, "max_path_length": 8, "max_diffuse_bounces": 2, "ms_per_frame": a["totalMs"],
                 "samples_per_s": W * H * SPP / (a["totalMs"] * 1e-3), "mrays_per_s": (a["extensionRays"] + a["shadowRays"]) / a["totalMs"] / 1e3,
                 "stage_ms": {k: a[k] for k in ("generateExtendMs", "extendMs", "shadeMs", "connectMs", "finalizeMs")},
                 "extension_rays": a["extensionRays"], "shadow_rays": a["shadowRays"], "path_length_reached": a["pathLengthReached"]}
    core.Shutdown()
    print("c3", res["c3"], flush=True)

if "c4" in which:
    # C4: 10 meshes x 1M triangles, 1000 instances, per-frame vertex displacement on every mesh (same triangle count -> refit)
    # and new rigid transforms on every instance (top level rebuilt), 1080p, 1 spp
    W, H = 1920, 1080
    core = RenderCore(0); core.SetTarget(W, H, 1); core.Setting("epsilon", 1e-3); overrides(core)
    base = scenes.terrain(1000, 500, extent=6.0, seed=5)
    base[:, 1] *= 0.3
    mats = scenes.make_materials([dict(color=(0.7, 0.7, 0.7)), dict(color=(80, 80, 64))])
    core.SetSkyData(*scenes.gradient_sky())
    core.SetMaterials(mats)
    tris = scenes.core_tris_from_verts(base)
    for m in range(10):
        core.SetGeometry(m, base, tris)
    lq = scenes.quad((0, 60, 0), (0, -1, 0), 30, 30); lt = scenes.core_tris_from_verts(lq, material=1)
    core.SetGeometry(10, lq, lt)
    core.SetLights(scenes.tri_lights(lq, lt, mats, inst_idx=1000))
    rng = np.random.default_rng(3)
    pos = (rng.random((1000, 3)) * 2 - 1) * np.array([90, 25, 90])

    def set_instances(frame):
        for i in range(1000):
            a = 0.01 * frame * (1 + i % 7)
            m = np.eye(4, dtype=np.float32)
            m[0, 0], m[0, 2], m[2, 0], m[2, 2] = np.cos(a), np.sin(a), -np.sin(a), np.cos(a)
            m[:3, 3] = pos[i]
            core.SetInstance(i, i % 10, m)
        core.SetInstance(1000, 10)
        core.SetInstance(1001, -1)

    set_instances(0); core.FinalizeInstances()
    view = scenes.view_pyramid((0, 60, -200), (0, 0, 0), 45, W, H)
    upload_ms, build_ms, inst_ms, frames = [], [], [], []
    for f in range(1, 6):
        moved = base.copy(); moved[:, 1] += (0.2 * np.sin(base[:, 0] * 2 + f)).astype(np.float32)
        t0 = time.perf_counter()
        for m in range(10):
            core.SetGeometry(m, moved, None)      # positions only: shading records unchanged
        t1 = time.perf_counter()
        set_instances(f)
        t2 = time.perf_counter()
        core.FinalizeInstances()
        t3 = time.perf_counter()
        core.Render(view, 1)
        fs = core.GetFrameStats()
        upload_ms.append(1e3 * (t1 - t0)); inst_ms.append(1e3 * (t2 - t1)); build_ms.append(1e3 * (t3 - t2))
        frames.append({k: float(fs[k]) for k in ("totalMs", "generateExtendMs", "extendMs", "connectMs", "shadeMs", "buildMs")} |
                      {"rays": int(fs["extensionRays"]) + int(fs["shadowRays"])})
    refit_device = sum(float(core.GetBvhStats(m)["buildMs"]) for m in range(10))
    last = frames[-1]
    work = counted_frame(core, view)
    res["c4"] = {"traversal_work": work, "meshes": 10, "triangles_per_mesh": len(base) // 3, "instances": 1001, "vertex_upload_ms_per_frame": float(np.mean(upload_ms[1:])),
                 "set_instance_calls_ms": float(np.mean(inst_ms[1:])), "finalize_instances_wall_ms": float(np.mean(build_ms[1:])),
                 "refit_device_ms_10_meshes": refit_device, "tlas_device_ms": last["buildMs"], "render_ms": last["totalMs"],
                 "trace_mrays_per_s": last["rays"] / (last["totalMs"] - last["shadeMs"]) / 1e3,      # connect( L ) overlaps extend( L + 1 ): wall time spent tracing
                 "frame": last}
    # the same animation done on the device (lh2b_set_skin / lh2b_set_pose): 4 joints per mesh, new joint
    # matrices every frame -> skinning kernel + refit, no triangle data crosses PCIe (64 B per joint do)
    nverts = base.reshape(-1, 4).shape[0]
    jrng = np.random.default_rng(11)
    joints = jrng.integers(0, 4, (nverts, 4)).astype(np.uint32)
    wts = jrng.random((nverts, 4)).astype(np.float32); wts /= wts.sum(axis=1, keepdims=True)
    for m in range(10):
        core.SetGeometry(m, base, tris)
    core.FinalizeInstances()
    for m in range(10):
        core.SetSkin(m, joints, wts)
    pose_ms, fin_ms, rend = [], [], []
    for f in range(1, 6):
        mats4 = np.tile(np.eye(4, dtype=np.float32), (4, 1, 1))
        for q in range(4):
            mats4[q, 1, 3] = 0.2 * np.sin(f + q)
        t0 = time.perf_counter()
        for m in range(10):
            core.SetPose(m, mats4)
        t1 = time.perf_counter()
        set_instances(f)
        t2 = time.perf_counter()
        core.FinalizeInstances()
        t3 = time.perf_counter()
        core.Render(view, 1)
        fs = core.GetFrameStats()
        pose_ms.append(1e3 * (t1 - t0)); fin_ms.append(1e3 * (t3 - t2)); rend.append(float(fs["totalMs"]))
    res["c4"]["device_animation"] = {"set_pose_wall_ms_10_meshes": float(np.mean(pose_ms[1:])), "finalize_instances_wall_ms": float(np.mean(fin_ms[1:])),
                                     "refit_device_ms_10_meshes": sum(float(core.GetBvhStats(m)["buildMs"]) for m in range(10)),
                                     "render_ms": float(np.mean(rend[1:])), "pcie_bytes_per_frame": 10 * 4 * 64}
    core.Shutdown()
    print("c4", res["c4"], flush=True)

if "c5" in which:
    # C5: the C3 scene at 3840x2160, 1 spp, SVGF filter + TAA, moving camera: ms per stage
    W, H = 3840, 2160
    sd = scenes.config2_scene(1000, 500, n_materials=64, light_quads=8)
    core = RenderCore(0); core.SetTarget(W, H, 1); core.Setting("epsilon", 1e-3); core.Setting("filter", 1); core.Setting("TAA", 1); overrides(core)
    sd.upload(core)
    acc = {}
    n = 8
    for f in range(n):
        view = scenes.view_pyramid((0.2 * f, 30, -80 + 0.1 * f), (0, 0, 0), 40, W, H)
        core.Render(view, 1)
        fs = core.GetFrameStats()
        if f >= 2:
            for k in fs.dtype.names:
                if k != "reserved":
                    acc[k] = acc.get(k, 0.0) + float(fs[k]) / (n - 2)
    px = W * H
    res["c5"] = {"resolution": [W, H], "ms_per_frame": acc["totalMs"], "filter_ms": acc["filterMs"], "filter_gb_per_s_at_584_B_per_px": px * 584 / (acc["filterMs"] * 1e-3) / 1e9,
                 "stage_ms": {k: acc[k] for k in ("generateExtendMs", "extendMs", "shadeMs", "connectMs", "filterMs")},
                 "rays_per_frame": acc["extensionRays"] + acc["shadowRays"]}
    core.Shutdown()
    print("c5", res["c5"], flush=True)

os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
json.dump(res, open(out_path, "w"), indent=1)
print("wrote", out_path)
