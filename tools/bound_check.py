#!/usr/bin/env python3
"""Verify the certified brick signs of the refinement passes against a full evaluation on a named config (dev tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from poissonrecon_gpu_b200 import PoissonRecon, synth
cfg = sys.argv[1] if len(sys.argv) > 1 else "multi20m_d11"
p, n, D = synth.make(cfg)
pr = PoissonRecon(D)
pr.set_points(p, n); pr.run()
v, t = pr.mesh()
v, t = v.copy(), t.copy()
pr.set_option("refine_bound_check", 1)
pr.set_points(p, n); pr.run()          # raises if a certificate is wrong
v2, t2 = pr.mesh()
print(cfg, "certificates verified; mesh identical:", bool(np.array_equal(v, v2) and np.array_equal(t, t2)), v.shape, t.shape, flush=True)
