#!/usr/bin/env python
"""bench.py - BASELINE.json headline metric on its configs[1] workload: a synthetic 1M-triangle static scene,
1920x1080, primary + shadow rays only (path length 1: generate+extend, shade with NEE, connect, finalize),
reported as Mrays/s (extend + shadow).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

ours:       a "step" is one Render() of the workload on every rank. `value` = rays of all ranks / max-over-ranks
            device time of K steps (CUDA events on the core's launch stream, inputs resident in HBM).
            `e2e` = the same metric through the public API with host buffers: the ViewPyramid goes in from host
            memory and the finished RGBA32F frame is read back into pinned host memory every step.
            N > 1 (torchrun, one process per GPU): every rank renders the full frame for its own sample index
            (sample sharding, scene replicated, weak scaling) and the accumulators are summed on rank 0 with an
            NCCL reduce - the only exchange the path has.
reference:  the reference has no CPU (or any runnable) implementation of this path here (OptiX is closed and absent),
            so this arm times the CPU oracle port of the path (oracle/: generate, shade, connect restated in C++; ray
            queries = the exhaustive search pruned by the oracle's own binary BVH, lh2_oracle_bvh.h) on all host cores;
            each step is one full 1920x1080 frame of the same workload (BVH build excluded, as on the GPU).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, SPP = 1920, 1080, 1
NX, NZ = 1000, 500                     # 1000 x 500 cells x 2 = 1,000,000 terrain triangles (+ 2 light triangles)
CAM_POS, CAM_TARGET, FOV = (0.0, 30.0, -80.0), (0.0, 0.0, 0.0), 40.0
WORKLOAD = "synthetic 1M-triangle static scene, 1920x1080, primary+shadow rays only (configs[1])"
METRIC, UNIT = "Mrays/s (extend+shadow)", "Mrays/s"
EXTEND_BYTES_PER_RAY = 48              # 32 B ray (O4+D4) + 16 B hit record per extension ray


def build_scene():
    from lighthouse2_b200 import scenes
    sd = scenes.config2_scene(NX, NZ, n_materials=1, light_quads=1, seed=0x12345678)
    view = scenes.view_pyramid(CAM_POS, CAM_TARGET, FOV, W, H)
    return sd, view


class ClockSampler:
    """nvidia-smi sampled every 20 ms (recipe: /opt/skills/guides/B200_PROFILING.md). Started well before the timed region (the tool
    takes a few hundred ms to produce its first line); stop() keeps the samples that fall inside the window it is given."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            import atexit
            atexit.register(lambda: self.proc is not None and self.proc.poll() is None and self.proc.terminate())
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def window(self, t0, t1):
        """clock median / throttle reasons of the samples inside [t0, t1]"""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        rows = [r for t, r in self.rows if t0 - 0.02 <= t <= t1 + 0.05]
        if not rows:      # (nvidia-smi slower than the window: the samples closest to it)
            rows = [r for _, r in sorted(self.rows, key=lambda tr: min(abs(tr[0] - t0), abs(tr[0] - t1)))[:3]]
        sm, mx, pw, reasons = [], [], [], set()
        for r in rows:
            try:
                sm.append(float(r[0])), mx.append(float(r[1])), pw.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(pw) if pw else None}

    def close(self):
        if self.proc is not None and self.proc.poll() is None:
            self.proc.terminate()


def oracle_crop_run(sd, view, crop_w, crop_h, threads):
    """One pass of the CPU oracle over a centred crop_w x crop_h pixel crop of the 1080p frame (same rays as the GPU
    frame for those pixels). Returns (rays, seconds)."""
    import numpy as np
    from oracle import binding as orc
    # a crop keeps the pixel footprint: render the sub-window by shifting the view corners
    v = view.copy()
    p1, p2, p3 = (np.asarray(view[0][k], np.float64) for k in ("p1", "p2", "p3"))
    right, up = (p2 - p1) / W, (p3 - p1) / H
    x0, y0 = (W - crop_w) // 2, (H - crop_h) // 2
    n1 = p1 + right * x0 + up * y0
    v["p1"], v["p2"], v["p3"] = n1, n1 + right * crop_w, n1 + up * crop_h
    o = orc.FrameOracle(sd, crop_w, crop_h, 1, 1e-3, 10.0, 1, 1, threads=threads)
    t0 = time.perf_counter()
    o.render(v, 1)
    dt = time.perf_counter() - t0
    return sum(o.ray_counts), dt


CPU_SAMPLE = ("full 1920x1080 frames of the same workload ({frames} x {rays} rays, {secs:.1f} s): scalar C++ port of the path, ray queries "
              "through the oracle's binary SAH BVH (identical results to its exhaustive search), BVH build excluded")


def cpu_frames(sd, view, threads, frames, warmup=1):
    """Times `frames` full-frame passes of the CPU oracle (BVH-pruned) after `warmup` passes (the first builds and caches the
    BVH). Returns (rays, seconds)."""
    from oracle import binding as orc
    with orc.accel(1):
        for _ in range(warmup):
            oracle_crop_run(sd, view, W, H, threads)
        rays, secs = 0, 0.0
        for _ in range(frames):
            r, dt = oracle_crop_run(sd, view, W, H, threads)
            rays, secs = rays + r, secs + dt
    return rays, secs


def run_reference(args, rank, world):
    """--impl reference: CPU oracle port on all host cores, bounded crop per step. Rank 0 only."""
    if rank != 0:
        return
    from oracle import binding as orc
    orc.build()
    sd, view = build_scene()
    threads = os.cpu_count() or 1
    steps = min(args.steps, 200)                # one step = one full frame (about 1 s on 16 cores): bounded to a few minutes
    rays, secs = cpu_frames(sd, view, threads, steps, warmup=max(1, min(args.warmup, 5)))
    value = rays / secs / 1e6
    sample = CPU_SAMPLE.format(frames=steps, rays=rays // max(1, steps), secs=secs)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": secs / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "the reference ships no runnable implementation of this path (OptiX closed, no CPU tracer): "
                   "CPU oracle port timed instead", "sample": sample, "timed_steps": steps},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# Algorithmic instruction model of the traversal kernels (DESIGN.md section 3 derives it from the SASS of traverse_wide.cuh): thread
# instructions a perfectly packed warp would need - what the work counters of a launch are multiplied with for the issue roofline.
I_NODE = 183      # one node step: 5 LDG.128, 15 set-up, 8 x (6 PRMT + 6 FFMA + 2 FMNMX3 + FFMA + FADD + LOP3 + SHF), 12 SEL, 15 pop / push / masks
I_TRI = 45        # one Moeller-Trumbore test: 3 LDG.128, 3 cross products, 4 dot products, reciprocal, 6 compares / selects
I_PRIMARY = 70    # generating one primary ray (blue-noise fetches, pixel target, normalisation) and storing O4 / D4 / the hit record


def work_counters(core, view, lights_off_sd=None):
    """Traversal work of one frame from the counting instantiations (lh2b_trace_stats). lights_off_sd: render that frame without
    lights (no connect launch), so that the counters describe the generate+extend kernel alone; the lights are restored."""
    if lights_off_sd is not None:
        core.SetLights()
    core.Render(view, 1)
    core.TraceStatsEnable(True)
    core.Render(view, 1)
    st = core.TraceStatsRead()
    core.TraceStatsEnable(False)
    if lights_off_sd is not None:
        core.SetLights(lights_off_sd.tri_lights, lights_off_sd.point_lights, lights_off_sd.spot_lights, lights_off_sd.dir_lights)
        core.Render(view, 1)
    r = max(1, st["rays"])
    return {"rays": st["rays"], "node_steps_per_ray": st["nodeSteps"] / r, "tri_tests_per_ray": st["triTests"] / r,
            "warp_iterations_x32_per_ray": st["iterations"] * 32 / r, "lanes_per_node_phase": st["nodeLanes"] / max(1, st["nodePhases"]),
            "lanes_per_tri_phase": st["triLanes"] / max(1, st["triPhases"])}


def overrides(core):
    for k, v in os.environ.items():             # experiments: LH2B_SET_<setting>=<value> (recorded in config.overrides)
        if k.startswith("LH2B_SET_"):
            core.Setting(k[9:], float(v))


def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist
    from lighthouse2_b200 import RenderCore, scenes
    from lighthouse2_b200.distributed import PeerGatherRenderer, PipelinedShardedRenderer, TileShardedRenderer
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the core has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    device = f"cuda:{local_rank}"
    sampler = ClockSampler(local_rank) if rank == 0 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max_sum(mx_vals, sum_vals):
        if world == 1:
            return mx_vals, sum_vals
        t = torch.tensor(list(mx_vals) + list(sum_vals), dtype=torch.float64, device=device)
        mx, sm = t.clone(), t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        n = len(mx_vals)
        return [mx[i].item() for i in range(n)], [sm[n + i].item() for i in range(len(sum_vals))]

    # ================================ headline: C2 (configs[1]) =================================================================
    sd, view = build_scene()
    core = RenderCore(local_rank)
    core.SetTarget(W, H, SPP)
    core.Setting("epsilon", 1e-3)
    core.Setting("clampValue", 10.0)
    core.Setting("maxPathLength", 1)            # primary + shadow rays only
    overrides(core)
    sd.upload(core)
    bvh = core.GetBvhStats(0)
    work = work_counters(core, view, lights_off_sd=sd)      # generate+extend alone (outside every timed region)
    host_imgs = [torch.empty((H, W, 4), dtype=torch.float32, pin_memory=True) for _ in range(2)]
    host_nps = [t.numpy() for t in host_imgs]
    stream = torch.cuda.ExternalStream(core.Stream(), device=device)
    # Frames are pipelined (Setting "pipeline"): Render(async) enqueues frame k+1 behind frame k and then harvests frame k, so the
    # device never waits for the host between frames. N > 1: every rank renders its sample shard; accumulator snapshots are pushed
    # to rank 0 over NVLink peer memory and summed + finalized there by one kernel (csrc/gather.cu) while the next frame renders.
    # --collective nccl selects the torch.distributed reduce instead (the baseline this replaces).
    psr = None
    if world > 1:
        psr = PeerGatherRenderer(core, SPP, rank, world) if args.collective == "peer" else PipelinedShardedRenderer(core, SPP, rank, world, device)
    if psr is None:
        core.Setting("pipeline", 1)

    def enqueue(k, to_host):
        """One Restart frame; with to_host the finished frame of rank 0 goes to pinned host memory (asynchronously)."""
        if psr is not None:
            psr.frame(view, 1, host_imgs[k & 1] if (to_host and rank == 0) else None)
        else:
            core.Render(view, 1, True)               # ViewPyramid from host memory
            if to_host:
                core.ReadPixelsAsync(host_nps[k & 1])    # 33 MB device -> pinned host, overlapping the next frame

    def drain():
        if psr is not None:
            psr.finish()
            if args.collective == "peer":
                psr.join(core.Stream())              # so that an event recorded on the launch stream also covers the gather
            else:
                stream.wait_event(psr.last_event())
        else:
            core.WaitForRender()
            core.WaitReadPixels()

    stage = {"generateExtendMs": 0.0, "shadeMs": 0.0, "connectMs": 0.0, "finalizeMs": 0.0}

    def frame_rays():
        fs = core.GetFrameStats()
        for k in stage:
            stage[k] += float(fs[k])
        return int(fs["primaryRays"]) + int(fs["shadowRays"])

    def run(steps, to_host):
        rays = 0
        for k in range(steps):
            enqueue(k, to_host)
            if k > 0:
                rays += frame_rays()                 # statistics of the frame harvested by this call (k-1)
        drain()
        return rays + frame_rays()

    # ---- N > 1: in-run parity of the sharded frame (outside the timed region): frame 3 of this run, gathered on rank 0, against
    # frame 3 of ONE core rendering all world x SPP samples with the same seeds (float sums in another order: 1e-5 relative)
    parity_ok = None
    run(3, True)
    if world > 1:
        if rank == 0:
            got = host_nps[0].copy()                 # frame index 2 (k & 1 == 0)
            single = RenderCore(local_rank)
            single.SetTarget(W, H, SPP * world)
            single.Setting("epsilon", 1e-3), single.Setting("clampValue", 10.0), single.Setting("maxPathLength", 1)
            overrides(single)
            sd.upload(single)
            single.Render(view, 1), single.Render(view, 1)       # the two frames the work counters took on the sharded core
            single.Render(view, 1)                               # and the one that restored the lights
            for _ in range(3):
                single.Render(view, 1)
            want = single.ReadPixels()
            single.Shutdown()
            err = float(np.abs(got - want).max() / max(1e-6, np.abs(want).max()))
            parity_ok = {"ok": bool(err < 1e-5), "max_rel_err": err, "what": "gathered %d-sample frame vs one GPU rendering all samples, same seeds" % (SPP * world)}
        barrier()
    run(max(args.warmup, 3), True)
    # ---- device-timed run: inputs resident, CUDA events on the launch stream ---------------------------------
    for k in stage:
        stage[k] = 0.0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.time()
    ev0.record(stream)
    rays = run(args.steps, False)
    ev1.record(stream)
    barrier()
    t1 = time.time()
    ms = ev0.elapsed_time(ev1)
    timed_stage = dict(stage)
    # ---- end-to-end run through the public API: view from host memory, every frame read back to pinned host memory ----
    barrier()
    e0 = time.perf_counter()
    e_rays = run(args.steps, True)
    barrier()
    e_secs = time.perf_counter() - e0
    t2 = time.time()
    # clocks under load: the device-timed region and the end-to-end region that follows it run the same workload back to back
    clocks = sampler.window(t0, t2) if sampler else None
    # ---- sustained: the same step back to back for at least --sustain seconds (the headline region is a burst of tens of ms) ----
    sustained = None
    if args.sustain > 0:
        barrier()
        s0, sus_ms, sus_rays, sus_steps = time.time(), 0.0, 0, 0
        while time.time() - s0 < args.sustain and sus_steps < 200000:
            ev0.record(stream)
            sus_rays += run(200, False)
            ev1.record(stream)
            torch.cuda.synchronize()
            sus_ms += ev0.elapsed_time(ev1)
            sus_steps += 200
        s1 = time.time()
        (sus_ms,), (sus_rays,) = reduce_max_sum([sus_ms], [float(sus_rays)])
        sustained = {"value": sus_rays / (sus_ms * 1e-3) / 1e6, "unit": UNIT, "steps": sus_steps, "seconds": s1 - s0, "ms_per_step": sus_ms / sus_steps,
                     "clocks": sampler.window(s0, s1) if sampler else None}
    stage = timed_stage
    (ms, e_secs), (rays, e_rays) = reduce_max_sum([ms, e_secs], [float(rays), float(e_rays)])
    if psr is not None and args.collective == "peer":
        psr.close()
    core.Shutdown()
    del core
    extra = {}
    if not args.no_extra_configs:
        extra["c3"] = bench_c3(args, rank, world, local_rank, barrier, reduce_max_sum)
        extra["c5"] = bench_c5(args, rank, world, local_rank, barrier, reduce_max_sum)
        extra["c4"] = bench_c4(args, rank, world, local_rank)
        barrier()
    if sampler:
        sampler.close()
    if rank != 0:
        return
    value = rays / (ms * 1e-3) / 1e6
    e2e = e_rays / e_secs / 1e6
    # ---- roofline of the dominant kernel (generate+extend): live per-launch time from CUDA events on the launch stream. The kernel
    # is bound by instruction issue (the BVH is L2-resident: SURVEY.md 8d), so the roofline is thread instructions per second:
    # achieved = ALGORITHMIC instructions of one launch (work counters x the per-step instruction model above) / launch time,
    # peak = 148 SMs x 4 schedulers x 32 lanes x SM clock. Wasted instructions and idle lanes both lower the fraction. ------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    if os.path.exists(peaks_path):
        hbm_peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    ge_ms = stage["generateExtendMs"] / args.steps
    n_primary = W * H * SPP
    inst_per_ray = work["node_steps_per_ray"] * I_NODE + work["tri_tests_per_ray"] * I_TRI + I_PRIMARY
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    issue_peak = 148 * 4 * 32 * sm_mhz * 1e6 / 1e9                  # G thread-instructions / s
    issue_achieved = n_primary * inst_per_ray / (ge_ms * 1e-3) / 1e9
    hbm_achieved = n_primary * EXTEND_BYTES_PER_RAY / (ge_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("wideGenerateExtendKernel_dram_bytes_per_launch")
    roofline = {"kernel": "wideGenerateExtendKernel", "bound": "issue", "achieved": issue_achieved, "peak": issue_peak, "unit": "G thread-instructions/s",
                "frac": issue_achieved / issue_peak, "traffic": traffic, "ms_per_launch": ge_ms, "mrays_per_s": n_primary / ge_ms / 1e3,
                "algorithmic_thread_instructions_per_ray": inst_per_ray, "model": {"node_step": I_NODE, "triangle_test": I_TRI, "primary_ray": I_PRIMARY},
                "work": work, "peak_source": "148 SMs x 4 schedulers x 32 lanes x %.0f MHz (SM clock sampled under load)" % sm_mhz,
                "hbm": {"achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak, "peak_source": peak_src,
                        "note": "48 B/ray algorithmic (SURVEY.md 8d); not the bound: the 64 MB BVH is L2-resident, DRAM traffic per launch ~ the algorithmic bytes"},
                "note": "traversal is bound by instruction issue and the L1 data path, not by HBM or tensor cores (SURVEY.md 8d; ncu: profiles/r2_*): "
                        "the fraction is algorithmic instructions (work counters of this run x the per-step model) over the issue peak"}
    # ---- CPU baseline: oracle port on this box's host cores (rank 0, N = 1 only) ------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import binding as orc
        orc.build()
        threads = os.cpu_count() or 1
        r, dt = cpu_frames(sd, view, threads, 8)
        cpu = {"value": r / dt / 1e6, "unit": UNIT, "cores": threads, "kind": "port", "sample": CPU_SAMPLE.format(frames=8, rays=r // 8, secs=dt)}
    ref_kernels = None
    rk = os.path.join(ROOT, "profiles", "r2_reference_kernels.json")
    if os.path.exists(rk):
        ref_kernels = dict(json.load(open(rk)), source="profiles/r2_reference_kernels.json (tools/ref_kernel_timing.py on a B200; not re-measured by this run)")
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "triangles": int(bvh["triangles"]) + 2, "resolution": [W, H], "spp_per_gpu": SPP,
                   "path_length": 1, "overrides": {k[9:]: v for k, v in os.environ.items() if k.startswith("LH2B_SET_")}, "bvh": "8-wide compressed BVH, 80-byte nodes", "bvh_nodes": int(bvh["nodes"]),
                   "parallelism": "1 GPU" if world == 1 else f"sample-sharded x{world}, scene replicated, accumulators gathered on rank 0 ({args.collective})",
                   "frames": "pipelined: the next frame is enqueued while the previous one runs",
                   "l2": "no explicit flush: per-step working set (path state 0.2 GB + scene 0.3 GB) exceeds the 126 MB L2; the 64 MB BVH "
                         "staying L2-resident across frames is the steady state of the renderer"},
        "stage_ms_per_step": {k: v / args.steps for k, v in stage.items()},
        "rays_per_step": rays / args.steps,
        "roofline": roofline, "cpu_baseline": cpu,
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": 68, "d2h_bytes_per_step": W * H * 16, "ms_per_step": e_secs / args.steps * 1e3},
        "gpu_launches": (4 if world == 1 else 5) * args.steps, "clocks": clocks, "sustained": sustained, "parity_ok": parity_ok,
        "extra_configs": extra, "reference_kernels": ref_kernels}
    print(json.dumps(out))


def _timed_frames(core, stream, frame_fn, finish_fn, frames, warm, barrier):
    """`frames` pipelined frames after `warm` untimed ones, device-timed with CUDA events on the core's launch stream."""
    import torch
    for f in range(warm):
        frame_fn(f)
    finish_fn()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for f in range(frames):
        frame_fn(warm + f)
    finish_fn()
    ev1.record(stream)
    barrier()
    return ev0.elapsed_time(ev1)


def bench_c3(args, rank, world, local_rank, barrier, reduce_max_sum):
    """BASELINE.json configs[2]: 1M triangles, 64 materials, 8 emissive quads, 1080p, 16 spp per frame, path length 8, NEE,
    two diffuse bounces (the stock Optix7 build stops after one: stated variant). N GPUs: the 16 samples are sharded (16 / N per
    rank, same seeds as one GPU), accumulators gathered on rank 0 - strong scaling of one frame, reported as samples / s."""
    import torch
    from lighthouse2_b200 import RenderCore, scenes
    from lighthouse2_b200.distributed import PeerGatherRenderer
    TOTAL = 16
    if TOTAL % world:
        return {"skipped": "16 spp do not divide over %d ranks" % world}
    spp = TOTAL // world
    sd = scenes.config2_scene(NX, NZ, n_materials=64, light_quads=8)
    view = scenes.view_pyramid(CAM_POS, CAM_TARGET, FOV, W, H)
    core = RenderCore(local_rank)
    core.SetTarget(W, H, spp)
    core.Setting("epsilon", 1e-3), core.Setting("maxPathLength", 8), core.Setting("maxDiffuseBounces", 2)
    overrides(core)
    sd.upload(core)
    stream = torch.cuda.ExternalStream(core.Stream(), device=f"cuda:{local_rank}")
    psr = PeerGatherRenderer(core, spp, rank, world) if world > 1 else None
    if psr is None:
        core.Setting("pipeline", 1)
    rays = [0]

    def frame(f):
        if psr is not None:
            psr.frame(view, 1, None)
        else:
            core.Render(view, 1, True)

    def finish():
        if psr is not None:
            psr.finish()
            psr.join(core.Stream())
        else:
            core.WaitForRender()

    frames = 6
    ms = _timed_frames(core, stream, frame, finish, frames, 2, barrier)
    fs = core.GetFrameStats()
    rays_per_frame = float(fs["extensionRays"]) + float(fs["shadowRays"])
    stages = {k: float(fs[k]) for k in ("generateExtendMs", "extendMs", "shadeMs", "connectMs")}
    (ms,), (rays_per_frame,) = reduce_max_sum([ms], [rays_per_frame])
    if psr is not None:
        psr.close()
    core.Shutdown()
    return {"workload": "synthetic 1M-triangle scene, 1080p 16 spp, path length 8 with NEE, sharedBSDF (configs[2])", "n_gpus": world, "spp_per_gpu": spp,
            "frames": frames, "ms_per_frame": ms / frames, "samples_per_s": W * H * TOTAL / (ms / frames * 1e-3), "mrays_per_s": rays_per_frame / (ms / frames) / 1e3,
            "rays_per_frame": rays_per_frame, "stage_ms_rank0_last_frame": stages, "max_diffuse_bounces": 2,
            "parallelism": "1 GPU" if world == 1 else f"sample-sharded x{world} (strong scaling of one 16-spp frame), accumulators gathered on rank 0 over NVLink peer memory"}


def bench_c4(args, rank, world, local_rank):
    """BASELINE.json configs[3]: 10 meshes x 1M triangles, 1000 instances (+ a light quad), every frame new vertex positions on every mesh
    (same triangle count: BLAS refit) and new rigid transforms on every instance (top level rebuilt), 1080p 1 spp. The animation runs on the
    device (lh2b_set_pose: skinning kernel feeding the refit), so a frame moves joint matrices, not 2.5 GB of vertices, over PCIe. One GPU
    (rank 0); device times from the core's events."""
    if rank != 0:
        return None
    import numpy as np
    from lighthouse2_b200 import RenderCore, scenes
    core = RenderCore(local_rank)
    core.SetTarget(W, H, 1)
    core.Setting("epsilon", 1e-3)
    overrides(core)
    base = scenes.terrain(1000, 500, extent=6.0, seed=5)
    base[:, 1] *= 0.3
    mats = scenes.make_materials([dict(color=(0.7, 0.7, 0.7)), dict(color=(80, 80, 64))])
    core.SetSkyData(*scenes.gradient_sky())
    core.SetMaterials(mats)
    tris = scenes.core_tris_from_verts(base)
    for m in range(10):
        core.SetGeometry(m, base, tris)
    lq = scenes.quad((0, 60, 0), (0, -1, 0), 30, 30)
    lt = scenes.core_tris_from_verts(lq, material=1)
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

    set_instances(0)
    core.FinalizeInstances()
    nverts = base.reshape(-1, 4).shape[0]
    jrng = np.random.default_rng(11)
    joints = jrng.integers(0, 4, (nverts, 4)).astype(np.uint32)
    wts = jrng.random((nverts, 4)).astype(np.float32)
    wts /= wts.sum(axis=1, keepdims=True)
    for m in range(10):
        core.SetSkin(m, joints, wts)
    view = scenes.view_pyramid((0, 60, -200), (0, 0, 0), 45, W, H)
    rows = []
    for f in range(1, 6):
        mats4 = np.tile(np.eye(4, dtype=np.float32), (4, 1, 1))
        for q in range(4):
            mats4[q, 1, 3] = 0.2 * np.sin(f + q)
        t0 = time.perf_counter()
        for m in range(10):
            core.SetPose(m, mats4)
        set_instances(f)
        core.FinalizeInstances()
        host_ms = (time.perf_counter() - t0) * 1e3
        core.Render(view, 1)
        fs = core.GetFrameStats()
        refit = sum(float(core.GetBvhStats(m)["buildMs"]) for m in range(10))
        rays = int(fs["extensionRays"]) + int(fs["shadowRays"])
        trace = float(fs["totalMs"]) - float(fs["shadeMs"]) - float(fs["finalizeMs"])     # connect( L ) overlaps extend( L + 1 ): wall time spent tracing
        rows.append((float(fs["totalMs"]), refit, float(fs["buildMs"]), host_ms, rays, trace))
    core.Shutdown()
    r = np.array(rows[1:])
    return {"workload": "synthetic 10M-triangle, 1k-instance animated scene with per-frame BLAS refit + TLAS rebuild, 1080p 1 spp (configs[3])", "n_gpus": 1,
            "frames": len(r), "render_ms_per_frame": float(r[:, 0].mean()), "refit_ms_10_meshes": float(r[:, 1].mean()), "tlas_build_ms": float(r[:, 2].mean()),
            "host_ms_set_pose_set_instance_finalize": float(r[:, 3].mean()), "rays_per_frame": float(r[:, 4].mean()),
            "trace_mrays_per_s": float((r[:, 4] / r[:, 5]).mean() / 1e3), "traversal": "two-level (1000 transformed instances of ten 1M-triangle meshes)",
            "pcie_bytes_per_frame": 10 * 4 * 64 + 1001 * 64, "note": "trace_mrays_per_s = rays / (frame - shade - finalize) device time: connect( L ) overlaps extend( L + 1 )"}


def bench_c5(args, rank, world, local_rank, barrier, reduce_max_sum):
    """BASELINE.json configs[4]: 4K, 1 spp, SVGF filter + TAA, moving camera, real-time frame-time mode. N GPUs: the frame is
    tile-sharded - rendering AND the filter chain (csrc/tile_gather.cu, Setting tileFilterShard): every rank path-traces interleaved
    tile rows and filters a band, rank 0 collects the presented bands. The frame ends in rank 0's device pixel buffer (what a display
    path consumes); reported as ms per frame."""
    import torch
    from lighthouse2_b200 import RenderCore, scenes
    from lighthouse2_b200.distributed import TileShardedRenderer
    W5, H5 = 3840, 2160
    sd = scenes.config2_scene(NX, NZ, n_materials=64, light_quads=8)
    core = RenderCore(local_rank)
    core.SetTarget(W5, H5, 1)
    core.Setting("epsilon", 1e-3), core.Setting("filter", 1), core.Setting("TAA", 1)
    overrides(core)
    sd.upload(core)
    stream = torch.cuda.ExternalStream(core.Stream(), device=f"cuda:{local_rank}")
    view_of = lambda f: scenes.view_pyramid((0.2 * f, 30, -80 + 0.1 * f), (0, 0, 0), 40, W5, H5)
    r = TileShardedRenderer(core, rank, world, filter_shard=1, interleave=1) if world > 1 else None
    if r is None:
        core.Setting("pipeline", 1)

    def frame(f):
        if r is not None:
            r.frame(view_of(f), 1, None)
        else:
            core.Render(view_of(f), 1, True)

    def finish():
        if r is not None:
            r.finish()
        else:
            core.WaitForRender()

    frames = 12
    t0 = time.perf_counter()
    ms = _timed_frames(core, stream, frame, finish, frames, 4, barrier)
    fs = core.GetFrameStats()
    stages = {k: float(fs[k]) for k in ("generateExtendMs", "extendMs", "shadeMs", "connectMs", "filterMs")}
    if world > 1:
        stages["filterMs"] = None          # the tail of a tile-sharded frame is enqueued by the gatherer, outside the core's stage events
    (ms,), _ = reduce_max_sum([ms], [])
    if r is not None:
        r.close()
    core.Shutdown()
    px = W5 * H5
    return {"workload": "4K 1 spp + SVGF temporal / a-trous filter + TAA, moving camera, real-time frame-time mode (configs[4])", "n_gpus": world,
            "frames": frames, "ms_per_frame": ms / frames, "fps": frames / (ms * 1e-3), "stage_ms_rank0_last_frame": stages,
            "filter_gb_per_s_at_584_B_per_px": (px * 584 / (stages["filterMs"] * 1e-3) / 1e9) if (world == 1 and stages["filterMs"] > 0) else None,
            "parallelism": "1 GPU" if world == 1 else f"tile-sharded x{world}: every rank path-traces interleaved 4-row tile rows and filters a band of the frame (band + 16 halo rows; "
                           "history rows read from their owner over NVLink inside the filter kernels; the chain of frame k runs next to the path tracing of frame k + 1); rank 0 collects the presented bands"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip the short C3 / C5 runs (extra_configs)")
    ap.add_argument("--sustain", type=float, default=2.0, help="seconds of the sustained C2 loop (0: skip)")
    ap.add_argument("--collective", default="peer", choices=["peer", "nccl"], help="N > 1: own NVLink peer-memory gather (default) or NCCL reduce")
    args = ap.parse_args()
    rank, world, local_rank = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
