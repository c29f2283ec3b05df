#!/bin/bash
# round 2, call Y (2 GPUs): 2-rank tests + sharded bench line after the upload / scan-total / option changes
mkdir -p gpurun_out/r02y
( time timeout 300 python -m pytest tests/test_multi_gpu.py tests/test_cli_gpu.py -m gpu -q ) > gpurun_out/r02y/pytest_mg.log 2>&1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02y/bench_2gpu.log 2>&1
tail -3 gpurun_out/r02y/pytest_mg.log; grep '^{"metric"' gpurun_out/r02y/bench_2gpu.log | cut -c1-330; grep -o '"digests".*' gpurun_out/r02y/bench_2gpu.log | cut -c1-400
