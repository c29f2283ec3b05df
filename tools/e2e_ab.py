#!/usr/bin/env python3
"""End-to-end (pinned host in, mesh out) time with and without the early mesh download; checks that both return the device mesh (dev tool).
usage: e2e_ab.py [config] [steps]"""
import hashlib, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from poissonrecon_gpu_b200 import PoissonRecon, synth
cfg = sys.argv[1] if len(sys.argv) > 1 else "scan5m_d10"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
p, n, D = synth.make(cfg, None)
N = p.shape[0]
hp, hn = torch.from_numpy(p.copy()).pin_memory(), torch.from_numpy(n.copy()).pin_memory()
dp, dn = hp.cuda(), hn.cuda()
pr = PoissonRecon(D)
stream = torch.cuda.ExternalStream(pr.stream())


def timed(fn, k):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(k):
        fn()
    e1.record(stream)
    e1.synchronize()
    return e0.elapsed_time(e1) / k


def resident():
    pr.set_points(dp.data_ptr(), dn.data_ptr(), N)
    pr.run()


def e2e():
    pr.set_points(hp.data_ptr(), hn.data_ptr(), N)
    pr.run()
    return pr.mesh_host_view()


for _ in range(3):
    resident()
print("resident ms", round(timed(resident, steps), 3), flush=True)
ref = None
for early in (0, 1, 0, 1):
    pr.set_option("early_mesh_copy", early)
    for _ in range(2):
        e2e()
    ms = timed(e2e, steps)
    v, t = e2e()
    dv, dt = pr.get("mesh_v", "<f4").reshape(-1, 3), pr.get("mesh_t", "<i4").reshape(-1, 3)
    ok = np.array_equal(v, dv) and np.array_equal(t, dt)
    h = hashlib.sha256(np.ascontiguousarray(v).tobytes() + np.ascontiguousarray(t).tobytes()).hexdigest()[:16]
    ref = ref or h
    print(f"early_mesh_copy={early}: e2e {ms:.3f} ms = {N / ms / 1e3:.1f} Mpoints/s, host mesh == device mesh: {ok}, digest {h} {'same' if h == ref else 'DIFFERENT'}", flush=True)
    assert ok and h == ref
