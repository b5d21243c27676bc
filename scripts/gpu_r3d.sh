#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_nn_match_gpu.py -m gpu -q -x --timeout 120 > gpurun_out/r3d_pytest.log 2>&1; echo "tests rc=$?" > gpurun_out/r3d_rc.txt
timeout 300 python scripts/nn_match_time.py > gpurun_out/r3d_nn_time.txt 2>&1; echo "time rc=$?" >> gpurun_out/r3d_rc.txt
cat gpurun_out/r3d_rc.txt; tail -n 15 gpurun_out/r3d_pytest.log; cat gpurun_out/r3d_nn_time.txt
