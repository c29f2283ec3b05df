#!/bin/bash
# round 2, call AB: k_splat with only the innermost neighbour loop unrolled: parity subset + timing
mkdir -p gpurun_out/r02ab
timeout 240 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "small_configs or edge_cases or config1" > gpurun_out/r02ab/parity.log 2>&1
timeout 120 python tools/step_times.py scan5m_d10 > gpurun_out/r02ab/steps.log 2>&1
tail -3 gpurun_out/r02ab/parity.log; grep resident gpurun_out/r02ab/steps.log | tail -3
