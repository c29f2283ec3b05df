#!/bin/bash
# round 2, call AF: chunked brick classification + prechecked lower-face values: full GPU suite, then timings at depth 10 and 11
mkdir -p gpurun_out/r02af
( time timeout 600 python -m pytest tests -m gpu -q -x ) > gpurun_out/r02af/pytest_gpu.log 2>&1
timeout 100 python tools/step_times.py scan5m_d10 > gpurun_out/r02af/steps_d10.log 2>&1
timeout 100 python tools/quick_bench.py multi20m_d11 - 3 > gpurun_out/r02af/quick_d11.log 2>&1
tail -4 gpurun_out/r02af/pytest_gpu.log; grep -E "^FAILED|Error" gpurun_out/r02af/pytest_gpu.log | head -5; grep resident gpurun_out/r02af/steps_d10.log | tail -2; grep wall_ms gpurun_out/r02af/quick_d11.log | tail -2 | cut -c1-300
