#!/usr/bin/env python3
"""Dev tool: where the CG kernel spends its time (CTA 0 of every rank: phase C / barrier / SpMV / barrier / phase B / barrier), plus
the stage timings, at any number of ranks (python tools/mg_phases.py cfg, or under torchrun)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from poissonrecon_gpu_b200 import PoissonRecon, synth
cfg = sys.argv[1] if len(sys.argv) > 1 else "scan5m_d10"
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
p, n, D = synth.make(cfg)
N = p.shape[0]
pr = PoissonRecon(D, device=local)
if world > 1:
    pr.mg_setup(int(float(os.environ.get("PRB_ARENA_GB", "8")) * (1 << 30)))
s0, s1 = (N * rank) // world, (N * (rank + 1)) // world
dp, dn = torch.from_numpy(p[s0:s1].copy()).cuda(), torch.from_numpy(n[s0:s1].copy()).cuda()
pr.set_option("cg_timing", 1)
pr.set_option("detail", 1)
for k in range(4):
    if world > 1:
        dist.barrier()
        pr.set_points_sharded(dp.data_ptr(), dn.data_ptr(), N)
    else:
        pr.set_points(dp.data_ptr(), dn.data_ptr(), N)
    pr.run()
    st = pr.stats()
    ph = pr.get("cg_phase_ns", "<i8") / 1e6
    if k >= 2:
        print(f"[rank {rank}/{world}] {cfg} stages", {a[3:]: round(b, 2) for a, b in st.items() if a.startswith("ms_")},
              "| CG ms: phaseC %.2f sync %.2f spmv %.2f sync %.2f phaseB %.2f sync %.2f" % tuple(ph[:6]), "iters", max(st["cg_iters"]), flush=True)
        if k == 3 and rank in (0, world - 1):
            names = pr.get("detail_names", "u1").tobytes().decode().split("\0")[:-1]
            ms = pr.get("detail_ms", "<f4")
            print(f"[rank {rank}/{world}] timeline: " + "; ".join(f"{n.split(' -> ')[1]} {m:.2f}" for n, m in zip(names, ms)), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
