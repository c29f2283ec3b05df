#!/bin/bash
# dev loop: parity tests + per-step timings + one ncu capture of the CG kernel (run under gpurun)
tag=$1
python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_$tag.log
python tools/step_times.py scan5m_d10 > gpurun_out/step_times_$tag.log 2>&1
python tools/step_times.py scan5m_d10 cg_zigzag=0 > gpurun_out/step_times_${tag}_nozz.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_cg_all_depths -s 3 -c 1 -o gpurun_out/prof_cg_$tag -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-reference-cuda > gpurun_out/ncu_cg_$tag.log 2>&1
