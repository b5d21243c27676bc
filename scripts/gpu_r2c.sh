#!/bin/bash
# round 2, GPU call C: fixed tests + GoodCorresNet, MLP32 per-layer timing, ncu of the mlp32 GEMM and of the backward fit
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mlp32_gpu.py tests/test_goodcorresnet.py tests/test_all_layers.py -m gpu -q -s --timeout 180 > gpurun_out/r2c_new.log 2>&1; echo "new tests rc=$?" > gpurun_out/r2c_rc.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r2c_pytest_gpu.log 2>&1; echo "gpu suite rc=$?" >> gpurun_out/r2c_rc.txt
timeout 300 python scripts/mlp32_time.py 512 1000 7 > gpurun_out/r2c_mlp32_time.log 2>&1
timeout 300 python scripts/mlp32_time.py 64 1000 7 >> gpurun_out/r2c_mlp32_time.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2c_mlp32_launches.csv python scripts/ncu_mlp32_target.py 512 > gpurun_out/r2c_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fepe_mlp32_gemm_kernel -s 4 -c 4 -o gpurun_out/r2c_mlp32_gemm python scripts/ncu_mlp32_target.py 512 > gpurun_out/r2c_ncu2.log 2>&1
echo "ncu rc=$?" >> gpurun_out/r2c_rc.txt
cat gpurun_out/r2c_rc.txt; tail -3 gpurun_out/r2c_new.log; cat gpurun_out/r2c_mlp32_time.log
