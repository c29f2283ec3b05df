#!/bin/bash
mkdir -p gpurun_out/r02u
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/r02u/pytest_gpu.log 2>&1
( time timeout 200 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r02u/smoke.log 2>&1
tail -4 gpurun_out/r02u/pytest_gpu.log; grep -E "^FAILED" gpurun_out/r02u/pytest_gpu.log | head; tail -3 gpurun_out/r02u/smoke.log
