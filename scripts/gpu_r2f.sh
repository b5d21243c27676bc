#!/bin/bash
# round 2, GPU call F: full GPU suite, sanitizer over the new kernels, ncu of the backward kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r2f_pytest_gpu.log 2>&1; echo "gpu suite rc=$?" > gpurun_out/r2f_rc.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_mlp32_gpu.py -m gpu -q -x --timeout 600 -k "test_first_layer or test_last_layer or test_normbwd or test_prepare or test_statistics or (test_wgrad and (1-128 or 2-300)) or (test_gemm_matches and (1-128-64-64 or 2-100-64-128 or 5-700))" > gpurun_out/r2f_sanitizer_memcheck_mlp32.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2f_rc.txt
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_mlp32_gpu.py -m gpu -q -x --timeout 600 -k "(test_wgrad and 1-128) or (test_gemm_matches and (1-128-64-64 or 2-100-64-128))" > gpurun_out/r2f_sanitizer_racecheck_mlp32.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r2f_rc.txt
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"fepe_fit_bwd|fepe_pose_bwd|normbwd|wgrad|last_bwd|first_bwd|fepe_mlp32_gemm" --csv --log-file gpurun_out/r2f_bwd_launches.csv python scripts/ncu_bwd_target.py > gpurun_out/r2f_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fepe_fit_bwd_kernel" -c 4 -o gpurun_out/r2f_fit_bwd python scripts/ncu_bwd_target.py > gpurun_out/r2f_ncu2.log 2>&1
echo "ncu rc=$?" >> gpurun_out/r2f_rc.txt
cat gpurun_out/r2f_rc.txt; tail -4 gpurun_out/r2f_pytest_gpu.log; tail -3 gpurun_out/r2f_sanitizer_memcheck_mlp32.log; tail -3 gpurun_out/r2f_sanitizer_racecheck_mlp32.log
