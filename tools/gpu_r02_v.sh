#!/bin/bash
# round 2, call V: the two opt-in modes + a quick default-path regression + smoke
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "cascadic or density_weighted or small_configs or edge_cases or config1" > gpurun_out/v_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/v_tests.log
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/v_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/v_smoke.log
tail -15 gpurun_out/v_tests.log; tail -3 gpurun_out/v_smoke.log
