"""Multi-GPU host logic on CPU (gloo, world_size 2) and, on a box with >= 2 GPUs, the sharded
pipeline against the single-GPU one."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, ctypes
sys.path.insert(0, os.environ["PRB_ROOT"])
import torch, torch.distributed as dist
from poissonrecon_gpu_b200 import shard_plan, api
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
# every rank derives the same plan from the same replicated counts
counts = [1, 8, 64, 448, 1704, 6608, 30568, 149176, 767080, 3694696, 14106192, 0, 5]
plans = [shard_plan(c, world) for c in counts]
allp = [None] * world
dist.all_gather_object(allp, plans)
assert all(p == allp[0] for p in allp), "ranks disagree on the shard plan"
for c, pl in zip(counts, plans):
    assert pl[0] == 0 and pl[-1] == c and all(b >= a for a, b in zip(pl, pl[1:])), (c, pl)
    sizes = [b - a for a, b in zip(pl, pl[1:])]
    assert max(sizes) - min(sizes) <= 1, (c, pl)
# my share + the other ranks' shares tile the range exactly
mine = [(pl[rank], pl[rank + 1]) for pl in plans]
shares = [None] * world
dist.all_gather_object(shares, mine)
for k, c in enumerate(counts):
    covered = sorted(s[k] for s in shares)
    assert covered[0][0] == 0 and covered[-1][1] == c and all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
# handle exchange transport used by PoissonRecon.mg_setup (64-byte blobs through all_gather_object)
blob = bytes([rank]) * 64
got = [None] * world
dist.all_gather_object(got, blob)
assert [g[0] for g in got] == list(range(world)) and all(len(g) == 64 for g in got)
# refinement passes are dealt out whole: every rank derives the same deal, heaviest passes are spread, nothing is lost
import numpy as np
from poissonrecon_gpu_b200 import deal_passes, assemble_mesh
D = 10
depths = [1, 2, 2, 3, 4, 5, 6, 7, 8, 9]
roots = [1, 1, 1, 4, 37, 210, 1400, 9000, 52000, 301000]
own = deal_passes(D, depths, roots, world)
owns = [None] * world
dist.all_gather_object(owns, own)
assert all(o == owns[0] for o in owns), "ranks disagree on the deal of the passes"
assert len(own) == len(depths) and set(own) <= set(range(world)) and len(set(own)) == world
assert deal_passes(D, [], [], world) == []
# the distributed mesh: every rank holds pieces (pass, first vertex, vertices, first triangle, triangles); writing all pieces at
# their offsets reproduces the whole mesh.  Pieces here: the main pass split by rank + one refinement pass per rank
g = np.random.default_rng(3)
NV, NT = 1000, 1800
Vfull = g.random((NV, 3)).astype(np.float32)
Tfull = g.integers(0, NV, (NT, 3)).astype(np.int32)
cutsV = [0, 300, 700, 850, 1000]
cutsT = [0, 500, 1100, 1500, 1800]
pieces = {0: [(0, 0, 1)], 1: [(0, 1, 2)]}                       # (pass, piece index range) per rank
pieces[0].append((2, 3, 4)); pieces[1].append((1, 2, 3))
lay, vv, tt = [], [], []
for ps, a, b in pieces[rank % 2] if world == 2 else []:
    lay.append([ps, cutsV[a], cutsV[b] - cutsV[a], cutsT[a], cutsT[b] - cutsT[a]])
    vv.append(Vfull[cutsV[a]:cutsV[b]]); tt.append(Tfull[cutsT[a]:cutsT[b]])
parts = [None] * world
dist.all_gather_object(parts, (np.array(lay, np.int64), np.concatenate(vv), np.concatenate(tt)))
V, T = assemble_mesh(parts, NV, NT)
assert np.array_equal(V, Vfull) and np.array_equal(T, Tfull)
# without a GPU the multi-GPU entry points fail loudly too
lib = api.load_library()
assert lib.prb_mg_init(None, 0, 2, 1 << 20, None) == -1
out = (ctypes.c_int64 * 3)()
assert lib.prb_mg_plan(-1, 2, out) == -1 and lib.prb_mg_plan(10, 9, out) == -1
assert lib.prb_set_points_sharded(None, None, None, 10) == -1
dist.barrier()
dist.destroy_process_group()
print("WORKER_OK", rank)
'''


def test_shard_plan_and_handle_exchange_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, PRB_ROOT=ROOT, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:]
    assert r.stdout.count("WORKER_OK") == 2


@pytest.mark.gpu
@pytest.mark.parametrize("config", ["sphere100k_d8", "torus1m_d9"])
def test_sharded_pipeline_matches_single_gpu(config):
    """2 ranks (torchrun, NCCL for the handle exchange): solution within 1e-5 rel-L2 of the 1-GPU
    run, identical iteration counts and identical mesh (tools/mg_check.py)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", PRB_ARENA_BYTES=str(3 << 30))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29534", os.path.join(ROOT, "tools", "mg_check.py"), config, "2"],
                       env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and "MG_CHECK PASS" in r.stdout, r.stdout[-3000:]
