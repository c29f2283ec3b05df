#!/bin/bash
# dev loop: parity tests + per-step timings + ncu launch list of one step (run under gpurun)
tag=$1
python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_$tag.log
python tools/step_times.py scan5m_d10 > gpurun_out/step_times_$tag.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$tag.csv python tools/step_times.py scan5m_d10 > /dev/null 2>&1
