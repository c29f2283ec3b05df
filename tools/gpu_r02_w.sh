#!/bin/bash
# round 2, call W: zero-copy scan totals, normals upload under the sort, early mesh download: correctness + e2e A/B
mkdir -p gpurun_out/r02w
timeout 120 python -m pytest tests/test_blocks_gpu.py -x -q -m gpu > gpurun_out/r02w/blocks.log 2>&1
timeout 240 python -m pytest tests/test_parity_gpu.py tests/test_cli_gpu.py -x -q -m gpu -k "small_configs or edge_cases or config1 or cli" > gpurun_out/r02w/parity.log 2>&1
timeout 200 python tools/e2e_ab.py scan5m_d10 10 > gpurun_out/r02w/e2e_ab.log 2>&1
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-cuda > gpurun_out/r02w/bench.log 2>&1
tail -2 gpurun_out/r02w/blocks.log; tail -4 gpurun_out/r02w/parity.log; tail -6 gpurun_out/r02w/e2e_ab.log; tail -1 gpurun_out/r02w/bench.log | cut -c1-700
