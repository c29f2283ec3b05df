#!/usr/bin/env python3
"""Dev tool: the CG kernel with TMA bulk-copy streaming (cg_bulk 1) against the per-thread cp.async ring (cg_bulk 0): x bit-identical, timings."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from poissonrecon_gpu_b200 import PoissonRecon, synth
for cfg in sys.argv[1:]:
    p, n, D = synth.make(cfg)
    pr = PoissonRecon(D)
    res = {}
    for mode in (1, 0, 1):
        pr.set_option("cg_bulk", mode)
        ts = []
        for _ in range(3):
            pr.set_points(p, n); pr.build_octree(); pr.splat(); pr.solve()
            st = pr.stats()
            ts.append(st["ms_solve"])
        x = pr.get("x", "<f4")
        print(cfg, "cg_bulk", mode, "ms_solve", [round(t, 3) for t in ts], "iters", st["cg_iters"][:D + 1], "GB/s", round(57.5 * st["cg_row_iters"] / min(ts) / 1e6, 1), "nan", int(np.isnan(x).sum()), flush=True)
        if mode in res:
            print(cfg, "repeat identical", np.array_equal(res[mode], x))
        res[mode] = x
    print(cfg, "bulk == ring bit-identical:", np.array_equal(res[0], res[1]), flush=True)
    pr.close()
