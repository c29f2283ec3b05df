#!/bin/bash
mkdir -p gpurun_out/r02d
( time timeout 600 python -m pytest tests/test_blocks_gpu.py -m gpu -q ) > gpurun_out/r02d/pytest_blocks.log 2>&1
( time timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "small_configs" ) > gpurun_out/r02d/pytest_small.log 2>&1
( time timeout 1500 python -m pytest tests/test_parity_gpu.py -m gpu -x -q ) > gpurun_out/r02d/pytest_parity.log 2>&1
timeout 300 python tools/step_times.py scan5m_d10 > gpurun_out/r02d/step_times.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02d/launches.csv python tools/step_times.py scan5m_d10 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r02d/launches.csv 50 > gpurun_out/r02d/launch_summary.txt 2>&1
grep -E "passed|failed" gpurun_out/r02d/pytest_blocks.log | tail -2; tail -3 gpurun_out/r02d/pytest_small.log; tail -5 gpurun_out/r02d/pytest_parity.log; tail -3 gpurun_out/r02d/step_times.log
