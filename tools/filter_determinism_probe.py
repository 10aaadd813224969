"""Runs our filter chain and the reference's (oracle/_ref/libref_filter_gpu.so) repeatedly on the inputs of tests/test_filter_gpu.py
and reports, per output buffer, how many runs differ from the first: ours must never differ; the reference's TAA stage reads and
writes one buffer in place and may."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_filter_gpu as t
from oracle import binding as orc
gen = t.case.__wrapped__() if hasattr(t.case, "__wrapped__") else t.case.__pytest_wrapped__.obj()
core, inputs, prev_view = next(gen)
st = dict(w=t.W, h=t.H, samplesTaken=1, camIsStationary=0, taa=1, directClamp=15.0, indirectClamp=15.0, j0=0.0, j1=0.0, prevj0=0.0, prevj1=0.0, prevView=prev_view)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 12
first_o, first_r, diff_o, diff_r, worst = None, None, {}, {}, {}
for i in range(N):
    io, got, keep = orc.make_filter_io(inputs, st)
    core.FilterChain(io)
    got = {k: np.array(v, copy=True) for k, v in got.items()}
    ref = {k: np.array(v, copy=True) for k, v in orc.ref_filter_gpu(inputs, st).items()}
    if first_o is None:
        first_o, first_r = got, ref
        continue
    for k in got:
        if not np.array_equal(got[k].view(np.uint32) if got[k].dtype == np.float32 else got[k], first_o[k].view(np.uint32) if got[k].dtype == np.float32 else first_o[k]):
            diff_o[k] = diff_o.get(k, 0) + 1
    for k in ref:
        a, b = np.asarray(ref[k]), np.asarray(first_r[k])
        if not np.array_equal(a, b):
            diff_r[k] = diff_r.get(k, 0) + 1
            if a.dtype == np.float32:
                worst[k] = max(worst.get(k, 0.0), float((np.abs(a - b) > 1e-2).any(axis=-1).mean()))
print("runs", N, "ours: buffers that differed from run 0:", diff_o or "none")
print("reference: buffers that differed from run 0:", diff_r or "none", "worst fraction of pixels off by > 1e-2:", worst)
