import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lighthouse2_b200 import RenderCore, scenes
W, H = 1920, 1080
mesh = scenes.terrain(1000, 500, extent=50, seed=0x12345678)
view = scenes.view_pyramid((0, 30, -80), (0, 0, 0), 40, W, H)
O, D = scenes.camera_rays(view, W, H)
dO, dD = torch.from_numpy(O).cuda(), torch.from_numpy(D).cuda()
hits = torch.empty((W * H, 4), dtype=torch.float32, device="cuda")
# incoherent set
rng = np.random.default_rng(1)
def run(settings):
    core = RenderCore(0)
    for k, v in settings.items(): core.Setting(k, v)
    core.SetGeometry(0, mesh); core.SetInstance(0, 0); core.SetInstance(1, -1); core.FinalizeInstances()
    core.SetGeometry(0, mesh); core.Setting("bvhRefit", 0); core.FinalizeInstances()   # warm second build for timing
    st = core.GetBvhStats(0)
    ms = min(core.TraceRaysDevice(dO.data_ptr(), dD.data_ptr(), W * H, hits.data_ptr(), repeat=10) for _ in range(5))
    h = hits.cpu().numpy(); t = h[:, 3]; hit = h.view(np.uint32)[:, 2] != 0xFFFFFFFF
    P = O[:, :3] + D[:, :3] * t[:, None]
    bO = np.zeros_like(O[hit]); bD = np.zeros_like(bO); n = int(hit.sum())
    r2 = np.random.default_rng(1)
    bO[:, :3] = P[hit] + np.array([0, 1e-2, 0], np.float32)
    d = r2.standard_normal((n, 3)); d[:, 1] = np.abs(d[:, 1]); d /= np.linalg.norm(d, axis=1, keepdims=True); bD[:, :3] = d
    dO3, dD3 = torch.from_numpy(bO).cuda(), torch.from_numpy(bD).cuda(); h3 = torch.empty((n, 4), dtype=torch.float32, device="cuda")
    ms3 = min(core.TraceRaysDevice(dO3.data_ptr(), dD3.data_ptr(), n, h3.data_ptr(), repeat=10) for _ in range(5))
    print(settings, f"build {float(st['buildMs']):.2f} ms nodes {int(st['nodes'])} primary {W*H*10/ms/1e3:.0f} diffuse {n*10/ms3/1e3:.0f} Mrays/s", flush=True)
    core.Shutdown()
for s in ({"bvhBuilder": 0, "plocRadius": 8}, {"bvhBuilder": 0, "plocRadius": 16}, {"bvhBuilder": 0, "plocRadius": 32}, {"bvhBuilder": 2},
          {"bvhBuilder": 2, "bvhMaxLeaf": 3}, {"bvhBuilder": 2, "bvhMaxLeaf": 1}, {"bvhBuilder": 0, "bvhMaxLeaf": 3}, {"bvhBuilder": 0, "bvhMaxLeaf": 1}, {"bvhBuilder": 1}):
    run(s)
