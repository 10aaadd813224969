"""Build / refit / top-level timings and trace quality of the GPU LBVH builder vs the host SAH builder."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from lighthouse2_b200 import RenderCore, scenes

W, H = 1920, 1080
mesh = scenes.terrain(1000, 500, extent=50, seed=0x12345678)
view = scenes.view_pyramid((0, 30, -80), (0, 0, 0), 40, W, H)
O, D = scenes.camera_rays(view, W, H)
dO, dD = torch.from_numpy(O).cuda(), torch.from_numpy(D).cuda()
hits = torch.empty((W * H, 4), dtype=torch.float32, device="cuda")
for name, b in (("gpu-ploc", 0), ("gpu-lbvh", 2), ("host-sah", 1)):
    core = RenderCore(0)
    core.Setting("bvhBuilder", b)
    t0 = time.time(); core.SetGeometry(0, mesh); t1 = time.time()
    core.SetInstance(0, 0); core.SetInstance(1, -1)
    core.FinalizeInstances(); t2 = time.time()
    st = core.GetBvhStats(0)
    ms = core.TraceRaysDevice(dO.data_ptr(), dD.data_ptr(), W * H, hits.data_ptr(), repeat=10)
    print(f"{name}: upload {1e3*(t1-t0):.1f} ms, finalize(wall) {1e3*(t2-t1):.1f} ms, build(device) {float(st['buildMs']):.2f} ms, nodes {int(st['nodes'])}, primary {W*H*10/ms/1e3:.0f} Mrays/s")
    if b != 1:
        # refit: same topology, displaced vertices
        m2 = mesh.copy(); m2[:, 1] += np.sin(m2[:, 0] * 0.3).astype(np.float32)
        core.SetGeometry(0, m2); t3 = time.time(); core.FinalizeInstances(); t4 = time.time()
        st = core.GetBvhStats(0)
        ms = core.TraceRaysDevice(dO.data_ptr(), dD.data_ptr(), W * H, hits.data_ptr(), repeat=10)
        print(f"  refit: finalize(wall) {1e3*(t4-t3):.1f} ms, device {float(st['buildMs']):.2f} ms, primary {W*H*10/ms/1e3:.0f} Mrays/s")
        core.Setting("bvhRefit", 0)
        core.SetGeometry(0, m2); core.FinalizeInstances(); st = core.GetBvhStats(0)
        ms = core.TraceRaysDevice(dO.data_ptr(), dD.data_ptr(), W * H, hits.data_ptr(), repeat=10)
        print(f"  full rebuild of displaced mesh: device {float(st['buildMs']):.2f} ms, primary {W*H*10/ms/1e3:.0f} Mrays/s")
    core.Shutdown()
# top level: 1000 instances of a small mesh, rebuilt per frame
core = RenderCore(0)
small = scenes.random_soup(2000, extent=1.0, size=0.4, seed=5)
core.SetGeometry(0, small)
rng = np.random.default_rng(1)
xf = []
for i in range(1000):
    m = np.eye(4, dtype=np.float32); m[:3, 3] = (rng.random(3) * 2 - 1) * 40; xf.append(m)
for i, m in enumerate(xf):
    core.SetInstance(i, 0, m)
core.SetInstance(1000, -1)
core.FinalizeInstances()
for rep in range(3):
    t0 = time.time(); core.FinalizeInstances(); t1 = time.time()
    fs_ms = float(core.GetFrameStats()["buildMs"])
    print(f"TLAS 1000 instances: FinalizeInstances wall {1e3*(t1-t0):.2f} ms, device {fs_ms:.3f} ms")
