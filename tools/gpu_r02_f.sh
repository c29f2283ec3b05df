#!/bin/bash
# 2 GPUs: sharded pipeline against the single-GPU one, CLI --gpus 2, bench at N=2
mkdir -p gpurun_out/r02f
export PRB_ARENA_BYTES=$((6<<30))
for cfg in sphere100k_d8 torus1m_d9 scan5m_d10; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/mg_check.py $cfg 3 > gpurun_out/r02f/mg_check_$cfg.log 2>&1
  tail -4 gpurun_out/r02f/mg_check_$cfg.log | cut -c1-400
done
( time timeout 300 python -m pytest tests/test_cli_gpu.py -m gpu -q -x ) > gpurun_out/r02f/pytest_cli.log 2>&1
tail -3 gpurun_out/r02f/pytest_cli.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02f/bench_2gpu.log 2>&1
tail -1 gpurun_out/r02f/bench_2gpu.log | cut -c1-1500
