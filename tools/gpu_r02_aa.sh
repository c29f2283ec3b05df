#!/bin/bash
# round 2, call AA: k_point_values with vector-loaded, predicated base functions: parity subset + A/B timing
mkdir -p gpurun_out/r02aa
timeout 240 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "small_configs or edge_cases or config1 or density" > gpurun_out/r02aa/parity.log 2>&1
timeout 120 python tools/step_times.py scan5m_d10 iso_vec_fn=0 > gpurun_out/r02aa/steps_vec0.log 2>&1
timeout 120 python tools/step_times.py scan5m_d10 iso_vec_fn=1 > gpurun_out/r02aa/steps_vec1.log 2>&1
tail -3 gpurun_out/r02aa/parity.log; grep resident gpurun_out/r02aa/steps_vec0.log | tail -2; grep resident gpurun_out/r02aa/steps_vec1.log | tail -2
