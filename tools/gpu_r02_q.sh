#!/bin/bash
# final 1-GPU pass: the whole GPU suite, both bench arms, ncu evidence, the depth-12 config
mkdir -p gpurun_out/r02q
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r02q/pytest_gpu.log 2>&1
( time timeout 400 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r02q/bench.log 2>&1
( time timeout 500 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/r02q/bench_ref.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02q/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-cuda > /dev/null 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_cg_all_depths -s 3 -c 1 -o gpurun_out/r02q/prof_cg -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-reference-cuda > gpurun_out/r02q/ncu_cg.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"k_div_fine|k_rv_brick_values|k_point_values|k_splat$" -s 4 -c 4 -o gpurun_out/r02q/prof_misc -f python tools/step_times.py scan5m_d10 > gpurun_out/r02q/ncu_misc.log 2>&1
( time timeout 600 python tools/quick_bench.py dense100m_d12 - 2 ) > gpurun_out/r02q/dense100m_d12_1gpu.log 2>&1
tail -4 gpurun_out/r02q/pytest_gpu.log; tail -2 gpurun_out/r02q/bench.log | cut -c1-400; tail -2 gpurun_out/r02q/bench_ref.log | cut -c1-600; tail -6 gpurun_out/r02q/dense100m_d12_1gpu.log | cut -c1-600
