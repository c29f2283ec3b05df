#!/usr/bin/env python3
"""Quick per-stage timing of the CUDA pipeline on a synthetic config (dev tool)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from poissonrecon_gpu_b200 import PoissonRecon, synth
cfg = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2] != '-' else None
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
t0 = time.time(); p, nr, D = synth.make(cfg, n); print('gen', round(time.time() - t0, 2), 's', p.shape, 'depth', D, flush=True)
pr = PoissonRecon(D)
for k in range(reps):
    t0 = time.time(); pr.set_points(p, nr); pr.run(); st = pr.stats(); wall = time.time() - t0
    print(k, 'wall_ms', round(wall * 1e3, 2), {a: round(b, 3) for a, b in st.items() if a.startswith('ms_')}, 'M', st['n_nodes'], 'nv', st['n_vertices'], 'nt', st['n_triangles'],
          'launch', st['kernel_launches'], 'iters', st['cg_iters'][:D + 1], 'sub', st['n_subdivide'], 'passes', st['n_passes'], flush=True)
print('Mpts/s (device total)', round(p.shape[0] / st['ms_total'] / 1e3, 3))
print('CG GB/s (57.5 B/row/iter)', round(57.5 * st['cg_row_iters'] / st['ms_solve'] / 1e6, 1))
print('nodes/depth', st['nodes_per_depth'][:D + 1])
