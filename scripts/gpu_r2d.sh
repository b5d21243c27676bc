#!/bin/bash
# round 2, GPU call D: refined split pipeline (fit tests + saturating bench), linear_tc32 diagnostics
mkdir -p gpurun_out
timeout 300 python scripts/diag_linear_tc32.py > gpurun_out/r2d_diag.log 2>&1
timeout 900 python -m pytest tests/test_fit_gpu.py tests/test_fit_bwd_gpu.py tests/test_model_gpu.py -m gpu -q -s --timeout 180 > gpurun_out/r2d_fit.log 2>&1; echo "fit tests rc=$?" > gpurun_out/r2d_rc.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?" >> gpurun_out/r2d_rc.txt
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"fepe_gram|fepe_solve|fepe_resid" -c 12 --csv --log-file gpurun_out/r2d_split_launches.csv python scripts/ncu_split.py > gpurun_out/r2d_ncu.log 2>&1
cat gpurun_out/r2d_rc.txt; cat gpurun_out/r2d_diag.log; grep -E "passed|failed|refined" gpurun_out/r2d_fit.log | tail -12
