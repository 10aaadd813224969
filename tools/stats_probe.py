"""Traversal work counters of the bench frame (C2) and of a C3-shaped frame: node steps / triangle tests per ray, SIMT utilisation of
the two phases. Usage: python tools/stats_probe.py [out.json]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lighthouse2_b200 import RenderCore, scenes

out = {}
W, H = 1920, 1080
for name, spp, plen, nmat, lights in (("c2", 1, 1, 1, 1), ("c3", 4, 8, 64, 8)):
    sd = scenes.config2_scene(1000, 500, n_materials=nmat, light_quads=lights, seed=0x12345678)
    view = scenes.view_pyramid((0, 30, -80), (0, 0, 0), 40, W, H)
    core = RenderCore(0)
    core.SetTarget(W, H, spp)
    core.Setting("epsilon", 1e-3); core.Setting("maxPathLength", plen); core.Setting("maxDiffuseBounces", 2)
    for k, v in os.environ.items():
        if k.startswith("LH2B_SET_"):
            core.Setting(k[9:], float(v))
    sd.upload(core)
    for _ in range(3):
        core.Render(view, 1)
    fs0 = core.GetFrameStats()
    core.TraceStatsEnable(True)
    core.Render(view, 1)
    st = core.TraceStatsRead()
    core.TraceStatsEnable(False)
    fs = core.GetFrameStats()
    r = max(1, st["rays"])
    out[name] = dict(st, node_steps_per_ray=st["nodeSteps"] / r, tri_tests_per_ray=st["triTests"] / r,
                     iterations_per_ray_x32=st["iterations"] * 32 / r,
                     node_phase_lanes=st["nodeLanes"] / max(1, st["nodePhases"]), tri_phase_lanes=st["triLanes"] / max(1, st["triPhases"]),
                     frame_ms=float(fs0["totalMs"]), generate_extend_ms=float(fs0["generateExtendMs"]), extend_ms=float(fs0["extendMs"]),
                     connect_ms=float(fs0["connectMs"]), shade_ms=float(fs0["shadeMs"]), counted_frame_ms=float(fs["totalMs"]),
                     bvh_nodes=int(core.GetBvhStats(0)["nodes"]))
    print(name, json.dumps(out[name]), flush=True)
    core.Shutdown()
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
