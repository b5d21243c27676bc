#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_fit_gpu.py tests/test_fit_bwd_gpu.py tests/test_model_gpu.py -m gpu -q -x --timeout 120 > gpurun_out/r3i_pytest.log 2>&1; echo "tests rc=$?" > gpurun_out/r3i_rc.txt
timeout 200 python scripts/split_trace.py 32768 > gpurun_out/r3i_trace.txt 2>&1; echo "trace rc=$?" >> gpurun_out/r3i_rc.txt
timeout 200 python scripts/split_time.py > gpurun_out/r3i_split_time.txt 2>&1; echo "time rc=$?" >> gpurun_out/r3i_rc.txt
cat gpurun_out/r3i_rc.txt; tail -n 3 gpurun_out/r3i_pytest.log; head -n 6 gpurun_out/r3i_trace.txt; grep "split" gpurun_out/r3i_split_time.txt | head -12
