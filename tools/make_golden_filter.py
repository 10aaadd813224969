"""Generates tests/golden/filter_reference_vectors.npz ON THE GPU BOX: outputs of the REFERENCE's own SVGF / TAA kernels
(lib/CUDA/shared_kernel_code/finalize_shared.h, unmodified, compiled for sm_100a into oracle/_ref/libref_filter_gpu.so by
oracle/Makefile) for fixed, seeded inputs. The inputs are produced without the product: the g-buffer comes from the CPU oracle's
ray queries, shading / history buffers from a seeded generator (the same recipe as tests/test_filter_gpu.py).
  gpurun -- 'python tools/make_golden_filter.py gpurun_out/filter_reference_vectors.npz'   then copy into tests/golden/.
Two runs are stored: A = moving camera with TAA (prepare incl. the diamond search, a-trous x3, TAA, unsharp), B = stationary
camera without TAA (finalizeNoTAA). The reference's TAA pass reads and writes one buffer in place - its output differs from run
to run - so for run A `taaPixels` and `target` are stored as the per-pixel minimum and maximum over 12 executions.
The CPU test tests/test_oracle_golden.py replays the inputs through the oracle's filter restatement and compares."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from lighthouse2_b200 import scenes
from oracle import binding as orc
import test_filter_gpu as t

W, H = 96, 64
t.W, t.H = W, H


class OracleTracer:
    """Stands in for the core in test_filter_gpu.gbuffer: hit records from the CPU oracle."""
    def __init__(self, sd):
        self.meshes, self.instances = [m[0] for m in sd.meshes], list(sd.instances)

    def TraceRays(self, O, D):
        with orc.accel(1):
            return orc.closest_hits(self.meshes, self.instances, O, D)


def golden_inputs():
    rng = np.random.default_rng(2024)
    sd = scenes.config2_scene(40, 28, n_materials=6, light_quads=2, floaters=200, seed=0xF117E2)
    tracer = OracleTracer(sd)
    prev_view = scenes.view_pyramid((0.0, 30, -80), (0, 0, 0), 40, W, H)
    view = scenes.view_pyramid((0.6, 30.2, -79.5), (0.1, 0, 0), 40, W, H)
    feat, wp, dd, albedo = t.gbuffer(tracer, sd, view, rng, spec_mat=1)
    _, pwp, _, _ = t.gbuffer(tracer, sd, prev_view, rng, spec_mat=1)
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
    return inputs, prev_view


def settings(prev_view, taa, stationary):
    return dict(w=W, h=H, samplesTaken=1, camIsStationary=stationary, taa=taa, directClamp=15.0, indirectClamp=15.0, j0=0.0, j1=0.0, prevj0=0.0, prevj1=0.0,
                prevView=prev_view)


if __name__ == "__main__":
    out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/filter_reference_vectors.npz"
    inputs, prev_view = golden_inputs()
    save = {"in_" + k: v for k, v in inputs.items()}
    save["prevView"] = np.frombuffer(np.ascontiguousarray(prev_view).tobytes(), np.float32).copy()
    runs = [orc.ref_filter_gpu(inputs, settings(prev_view, 1, 0)) for _ in range(12)]
    for k in orc.FILTER_OUTPUTS:
        stack = np.stack([np.asarray(r[k]) for r in runs])
        if k in ("taaPixels", "target"):
            save["A_" + k + "_min"], save["A_" + k + "_max"] = stack.min(axis=0), stack.max(axis=0)
        else:
            assert all(np.array_equal(stack[0], s) for s in stack[1:]), k + " of the reference is not deterministic"
            save["A_" + k] = stack[0]
    b = orc.ref_filter_gpu(inputs, settings(prev_view, 0, 1))
    for k in ("shadingAfterPrepare", "motion", "moments", "featuresOut", "phase3", "target"):
        save["B_" + k] = np.asarray(b[k])
    os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
    np.savez_compressed(out, **save)
    print("wrote", out, os.path.getsize(out), "bytes;", "racy pixels in A.target:",
          float((save["A_target_max"] - save["A_target_min"] > 1e-3).any(axis=-1).mean()))
