#!/bin/bash
# round 2, call A: sanity of the GPU suite, reference CUDA builds at depth 9 / 10 (ref+widen), both bench arms
mkdir -p gpurun_out/r02a
nproc > gpurun_out/r02a/host.txt; free -g >> gpurun_out/r02a/host.txt; nvidia-smi -L >> gpurun_out/r02a/host.txt
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r02a/pytest.log 2>&1
( time timeout 1500 python tools/ref_compare.py --config torus1m_d9 --runs 2 --out gpurun_out/r02a/ref_compare_torus1m_d9 ) > gpurun_out/r02a/ref_compare_d9.log 2>&1
python - > gpurun_out/r02a/ref_d10.log 2>&1 <<'P'
import os, sys, time, subprocess
sys.path.insert(0, os.getcwd())
from poissonrecon_gpu_b200 import synth, plyio
p, n, D = synth.make("scan5m_d10")
plyio.write_points_ply("/tmp/scan5m.ply", p, n)
t0 = time.time()
try:
    r = subprocess.run(["oracle/_ref/ref_poisson_d10_widen", "/tmp/scan5m.ply", "/tmp/scan5m_out.ply"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    print("rc", r.returncode, "wall", time.time() - t0)
    print(r.stdout[-6000:])
    os.system("head -12 /tmp/scan5m_out.ply | strings")
except subprocess.TimeoutExpired as e:
    print("TIMEOUT after", time.time() - t0, (e.stdout or b"")[-4000:])
P
( time python bench.py --steps 10 --warmup 3 ) > gpurun_out/r02a/bench.log 2>&1
( time python bench.py --impl reference --steps 10 --warmup 3 ) > gpurun_out/r02a/bench_ref.log 2>&1
tail -3 gpurun_out/r02a/pytest.log; tail -2 gpurun_out/r02a/bench.log | cut -c1-600; tail -1 gpurun_out/r02a/bench_ref.log | cut -c1-900
