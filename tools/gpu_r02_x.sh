#!/bin/bash
# round 2, call X: compute-sanitizer (memcheck / racecheck / synccheck / initcheck) on a small reconstruction through the command-line tool,
# early_mesh_copy test, bench re-measure
mkdir -p gpurun_out/r02x
python tools/make_input.py sphere100k_d8 /tmp/san.bnpts 20000 > gpurun_out/r02x/input.log 2>&1
B=poissonrecon_gpu_b200/poisson_recon
for tool in memcheck racecheck synccheck initcheck; do
  ( time timeout 170 compute-sanitizer --tool $tool --print-limit 30 $B --in /tmp/san.bnpts --out /tmp/san_$tool.ply --depth 7 --json ) > gpurun_out/r02x/san_$tool.log 2>&1
  echo "exit $?" >> gpurun_out/r02x/san_$tool.log
done
timeout 120 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "early_mesh_copy or config1" > gpurun_out/r02x/tests.log 2>&1
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-cuda > gpurun_out/r02x/bench.log 2>&1
for tool in memcheck racecheck synccheck initcheck; do echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|exit|real" gpurun_out/r02x/san_$tool.log | tail -4; done
tail -3 gpurun_out/r02x/tests.log; tail -1 gpurun_out/r02x/bench.log | cut -c1-300
