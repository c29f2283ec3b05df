#!/bin/bash
mkdir -p gpurun_out/r02e
( time timeout 180 python -m pytest tests/test_blocks_gpu.py -m gpu -q -x ) > gpurun_out/r02e/pytest_blocks.log 2>&1
( time timeout 120 python tools/cg_check.py sphere100k_d8 ) > gpurun_out/r02e/cg_check_small.log 2>&1
( time timeout 240 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "small_configs" ) > gpurun_out/r02e/pytest_small.log 2>&1
if grep -q "passed" gpurun_out/r02e/pytest_small.log && ! grep -q "failed" gpurun_out/r02e/pytest_small.log; then
  ( time timeout 200 python tools/cg_check.py scan5m_d10 ) > gpurun_out/r02e/cg_check.log 2>&1
  ( time timeout 1200 python -m pytest tests/test_parity_gpu.py -m gpu -x -q ) > gpurun_out/r02e/pytest_parity.log 2>&1
  ( time timeout 300 python -m pytest tests -m gpu -q --deselect tests/test_parity_gpu.py --deselect tests/test_blocks_gpu.py ) > gpurun_out/r02e/pytest_rest.log 2>&1
  timeout 200 python tools/step_times.py scan5m_d10 > gpurun_out/r02e/step_times.log 2>&1
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02e/launches.csv python tools/step_times.py scan5m_d10 > /dev/null 2>&1
  python tools/launch_summary.py gpurun_out/r02e/launches.csv 50 > gpurun_out/r02e/launch_summary.txt 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_vertex_values_stream -s 2 -c 1 -o gpurun_out/r02e/prof_vv -f python tools/step_times.py scan5m_d10 > gpurun_out/r02e/ncu_vv.log 2>&1
fi
grep -E "passed|failed" gpurun_out/r02e/pytest_blocks.log | tail -2; tail -4 gpurun_out/r02e/cg_check_small.log; tail -3 gpurun_out/r02e/pytest_small.log; tail -3 gpurun_out/r02e/cg_check.log; tail -5 gpurun_out/r02e/pytest_parity.log; tail -3 gpurun_out/r02e/pytest_rest.log; tail -3 gpurun_out/r02e/step_times.log
