#!/bin/bash
# round 2, call AD: bench line of BASELINE config 4 (20 M points, depth 11) on one GPU, final build
mkdir -p gpurun_out/r02ad
( time timeout 300 python bench.py --workload multi20m_d11 --steps 5 --warmup 3 --no-cpu-baseline --no-reference-cuda ) > gpurun_out/r02ad/bench_d11.log 2>&1
grep '^{"metric"' gpurun_out/r02ad/bench_d11.log | cut -c1-400; tail -4 gpurun_out/r02ad/bench_d11.log | cut -c1-200
