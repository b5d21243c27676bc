#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mlp32_gpu.py -m gpu -q -x --timeout 300 -k "wgrad or training or autograd or linear" > gpurun_out/r3a_pytest.log 2>&1; echo "tests rc=$?" > gpurun_out/r3a_rc.txt
timeout 300 python scripts/wgrad_time.py > gpurun_out/r3a_wgrad_time.txt 2>&1; echo "time rc=$?" >> gpurun_out/r3a_rc.txt
cat gpurun_out/r3a_rc.txt; tail -n 5 gpurun_out/r3a_pytest.log; cat gpurun_out/r3a_wgrad_time.txt
