#!/bin/bash
# bench line at N GPUs (N = $1), plus the 2-rank tests at N = 2
N=$1
mkdir -p gpurun_out/r02s
if [ "$N" = "2" ]; then ( time timeout 400 python -m pytest tests/test_multi_gpu.py tests/test_cli_gpu.py -m gpu -q ) > gpurun_out/r02s/pytest_mg.log 2>&1; tail -2 gpurun_out/r02s/pytest_mg.log; fi
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2960$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02s/bench_${N}gpu.log 2>&1
tail -1 gpurun_out/r02s/bench_${N}gpu.log | cut -c1-260
