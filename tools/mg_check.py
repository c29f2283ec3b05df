#!/usr/bin/env python3
"""Multi-GPU check + timing (torchrun, one rank per GPU): the ranks reconstruct one cloud together (every rank uploads its slice of
the samples; splat / divergence / CG / iso value / marching cubes sharded by Morton range, refinement passes dealt out) and the
gathered result is compared with a single-GPU run on the same GPU: x, pass list and the whole mesh BIT-IDENTICAL."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from poissonrecon_gpu_b200 import PoissonRecon, synth  # noqa: E402


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "torus1m_d9"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    p, n, D = synth.make(cfg)
    # single-GPU result on this rank
    ref = PoissonRecon(D, device=local)
    ref.set_points(p, n)
    ref.run()
    rx, (rv, rt), rst = ref.get("x", "<f4"), ref.mesh(), ref.stats()
    rpasses = ref.get("passes", "<i4").tolist()
    for _ in range(2):
        ref.set_points(p, n); ref.run()
    t1 = ref.stats()
    ref.close()
    pr = PoissonRecon(D, device=local)
    pr.mg_setup(arena_bytes=int(os.environ.get("PRB_ARENA_BYTES", 6 << 30)))
    ok = True
    for k in range(reps):
        dist.barrier()
        t0 = time.time()
        s0, s1 = (p.shape[0] * rank) // world, (p.shape[0] * (rank + 1)) // world
        pr.set_points_sharded(p[s0:s1], n[s0:s1], p.shape[0])
        pr.run()
        st = pr.stats()
        wall = time.time() - t0
        x = pr.get("x", "<f4")
        v, t = pr.mesh_global()
        rel = float(np.linalg.norm(x.astype(np.float64) - rx) / np.linalg.norm(rx.astype(np.float64)))
        same_mesh = v.shape == rv.shape and t.shape == rt.shape and np.array_equal(t, rt) and (v.size == 0 or float(np.abs(v - rv).max()) <= 1e-6)
        same_mesh = same_mesh and np.array_equal(v, rv) and np.array_equal(x, rx) and pr.get("passes", "<i4").tolist() == rpasses
        ok &= rel <= 1e-5 and same_mesh and st["cg_iters"] == rst["cg_iters"]
        print(f"[rank {rank}] run {k}: wall {wall * 1e3:.1f} ms  stages " + str({a: round(b, 2) for a, b in st.items() if a.startswith('ms_')}) +
              f"  x rel-L2 vs 1-GPU {rel:.2e}  mesh identical {same_mesh}  iters equal {st['cg_iters'] == rst['cg_iters']}", flush=True)
    if rank == 0:
        print("1-GPU stages", {a: round(b, 2) for a, b in t1.items() if a.startswith("ms_")}, flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MG_CHECK", "PASS" if int(flag.item()) == 1 else "FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
