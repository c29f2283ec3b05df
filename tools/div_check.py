#!/usr/bin/env python3
"""Dev tool: block-table / profile divergence (div_mode 1) against the first-version kernels (div_mode 0) on one context."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from poissonrecon_gpu_b200 import PoissonRecon, synth
from tests.cases import make_case
for cfg in sys.argv[1:]:
    p, n, D = synth.make(cfg) if cfg in synth.CONFIGS else make_case(cfg)
    pr = PoissonRecon(D)
    pr.set_points(p, n); pr.build_octree(); pr.splat()
    base = pr.get("base", "<i4")
    d1 = pr.get("divergence", "<f4")
    pr.set_option("div_mode", 0); pr.run_stage("divergence")
    d0 = pr.get("divergence", "<f4")
    for d in range(D + 1):
        a, b = d1[base[d]:base[d + 1]].astype(np.float64), d0[base[d]:base[d + 1]].astype(np.float64)
        nb = np.linalg.norm(b)
        print(cfg, "depth", d, "rows", a.size, "n_diff", int((a != b).sum()), "rel_l2", float(np.linalg.norm(a - b) / nb) if nb else float(np.linalg.norm(a - b)), "nan", int(np.isnan(a).sum()), flush=True)
    for mode in (1, 0):
        pr.set_option("div_mode", mode)
        ts = []
        for _ in range(5):
            pr.set_points(p, n); pr.build_octree(); pr.splat()
            ts.append(pr.stats()["ms_divergence"])
        print(cfg, "div_mode", mode, "ms_divergence", [round(t, 3) for t in ts], flush=True)
    pr.close()
