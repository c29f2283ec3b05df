#!/bin/bash
# 8 GPUs: depth 10 (check, phases, bench) and depth 11 (phases, bench)
mkdir -p gpurun_out/r02p
export PRB_ARENA_BYTES=$((6<<30)) PRB_ARENA_GB=16
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 120 $TR --master-port 29591 tools/mg_check.py scan5m_d10 1 > gpurun_out/r02p/mg_check_scan5m_d10.log 2>&1
timeout 100 $TR --master-port 29592 tools/mg_phases.py scan5m_d10 > gpurun_out/r02p/phases_d10_8gpu.log 2>&1
timeout 120 $TR --master-port 29593 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02p/bench_d10_8gpu.log 2>&1
timeout 150 $TR --master-port 29594 tools/mg_phases.py multi20m_d11 > gpurun_out/r02p/phases_d11_8gpu.log 2>&1
timeout 200 $TR --master-port 29595 bench.py --gpus 8 --steps 5 --warmup 3 --workload multi20m_d11 --arena-gb 16 > gpurun_out/r02p/bench_d11_8gpu.log 2>&1
grep -E "MG_CHECK" gpurun_out/r02p/mg_check_scan5m_d10.log | tail -1
grep "rank 0/8\] timeline" gpurun_out/r02p/phases_d10_8gpu.log | tail -1 | cut -c1-1500
tail -1 gpurun_out/r02p/bench_d10_8gpu.log | cut -c1-200
grep "rank 0/8\] timeline\|rank 7/8\] timeline" gpurun_out/r02p/phases_d11_8gpu.log | tail -2 | cut -c1-1500
tail -1 gpurun_out/r02p/bench_d11_8gpu.log | cut -c1-200
