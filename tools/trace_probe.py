"""Quick traversal throughput probe on the config-2 scene (1M triangles, 1080p primary rays)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from lighthouse2_b200 import RenderCore, scenes

nx, nz = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1000, 500)
t0 = time.time()
mesh = scenes.terrain(nx, nz, extent=50, seed=0x12345678)
print("scene: %d tris, gen %.1fs" % (mesh.shape[0] // 3, time.time() - t0))
core = RenderCore(0)
for k, v in os.environ.items():
    if k.startswith("LH2B_SET_"):
        core.Setting(k[9:], float(v)); print("setting", k[9:], v)
t0 = time.time()
core.SetGeometry(0, mesh)
core.SetInstance(0, 0); core.SetInstance(1, -1)
core.FinalizeInstances()
print("build %.2fs" % (time.time() - t0), core.GetBvhStats(0))
W, H = 1920, 1080
view = scenes.view_pyramid((0, 30, -80), (0, 0, 0), 40, W, H)
O, D = scenes.camera_rays(view, W, H)
dO, dD = torch.from_numpy(O).cuda(), torch.from_numpy(D).cuda()
hits = torch.empty((W * H, 4), dtype=torch.float32, device="cuda")
torch.cuda.synchronize()
for rep in (1, 3, 10):
    ms = core.TraceRaysDevice(dO.data_ptr(), dD.data_ptr(), W * H, hits.data_ptr(), repeat=rep)
    print("extend primary x%d: %.3f ms/launch, %.1f Mrays/s" % (rep, ms / rep, W * H * rep / ms / 1e3))
h = hits.cpu().numpy().view(np.uint32)
hit = h[:, 2] != 0xFFFFFFFF
print("hit fraction", hit.mean())
# shadow rays: from hit points toward a light at (0, 26, 0)
t = hits.cpu().numpy()[:, 3]
P = O[:, :3] + D[:, :3] * t[:, None]
L = np.array([0, 60.0, 0], np.float32) - P
dist = np.linalg.norm(L, axis=1)
sO = np.zeros_like(O); sD = np.zeros_like(D)
sO[:, :3] = P + L / dist[:, None] * 1e-3 + np.array([0, 1e-3, 0], np.float32)
sD[:, :3] = L / dist[:, None]; sD[:, 3] = dist - 2e-3
sO, sD = sO[hit], sD[hit]
n = sO.shape[0]
dO2, dD2 = torch.from_numpy(sO).cuda(), torch.from_numpy(sD).cuda()
occ = torch.empty(n, dtype=torch.uint8, device="cuda")
for rep in (1, 3, 10):
    ms = core.TraceShadowRaysDevice(dO2.data_ptr(), dD2.data_ptr(), n, occ.data_ptr(), repeat=rep)
    print("shadow x%d: %.3f ms/launch, %.1f Mrays/s" % (rep, ms / rep, n * rep / ms / 1e3))
print("occluded fraction", occ.float().mean().item())
# incoherent: shuffled secondary-like rays
rng = np.random.default_rng(1)
bO = np.zeros_like(O[hit]); bD = np.zeros_like(bO)
bO[:, :3] = P[hit] + np.array([0, 1e-2, 0], np.float32)
d = rng.standard_normal((n, 3)); d[:, 1] = np.abs(d[:, 1]); d /= np.linalg.norm(d, axis=1, keepdims=True)
bD[:, :3] = d
dO3, dD3 = torch.from_numpy(bO).cuda(), torch.from_numpy(bD).cuda()
hits3 = torch.empty((n, 4), dtype=torch.float32, device="cuda")
for rep in (1, 3, 10):
    ms = core.TraceRaysDevice(dO3.data_ptr(), dD3.data_ptr(), n, hits3.data_ptr(), repeat=rep)
    print("extend diffuse-bounce x%d: %.3f ms/launch, %.1f Mrays/s" % (rep, ms / rep, n * rep / ms / 1e3))
