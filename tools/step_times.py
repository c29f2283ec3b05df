#!/usr/bin/env python3
"""Per-step stage timings, resident then host-input steps (dev tool)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from poissonrecon_gpu_b200 import PoissonRecon, synth
cfg = sys.argv[1] if len(sys.argv) > 1 else "scan5m_d10"
p, n, D = synth.make(cfg)
N = p.shape[0]
hp, hn = torch.from_numpy(p).pin_memory(), torch.from_numpy(n).pin_memory()
dp, dn = hp.cuda(), hn.cuda()
pr = PoissonRecon(D)
for kv in sys.argv[2:]:
    k, v = kv.split('=')
    pr.set_option(k, float(v))
def show(tag, k, wall):
    st = pr.stats()
    print(tag, k, 'wall', round(wall * 1e3, 2), {a[3:]: round(b, 2) for a, b in st.items() if a.startswith('ms_')}, flush=True)
for k in range(5):
    t0 = time.time(); pr.set_points(dp.data_ptr(), dn.data_ptr(), N); pr.run(); w = time.time() - t0
    show('resident', k, w)
for k in range(3):
    t0 = time.time(); pr.set_points(hp.data_ptr(), hn.data_ptr(), N); pr.run(); pr.mesh_host_view(); w = time.time() - t0
    show('host', k, w)
for k in range(0):
    t0 = time.time(); pr.set_points(dp.data_ptr(), dn.data_ptr(), N); pr.run(); w = time.time() - t0
    show('resident2', k, w)
