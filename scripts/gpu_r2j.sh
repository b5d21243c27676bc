#!/bin/bash
# round 2, GPU call J (1 GPU): L2 prefetch of the streamed operand + vector reductions in wgrad: tests, timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mlp32_gpu.py tests/test_goodcorresnet.py tests/test_all_layers.py -m gpu -q --timeout 180 > gpurun_out/r2j_tests.log 2>&1; echo "mlp tests rc=$?" > gpurun_out/r2j_rc.txt
timeout 300 python scripts/mlp32_time.py 512 1000 7 > gpurun_out/r2j_mlp32_time.log 2>&1
timeout 300 python scripts/mlp32_time.py 64 1000 7 >> gpurun_out/r2j_mlp32_time.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"wgrad|fepe_mlp32_gemm|normbwd|last_bwd|first_bwd" --csv --log-file gpurun_out/r2j_bwd_launches.csv python scripts/ncu_bwd_target.py > gpurun_out/r2j_ncu1.log 2>&1
timeout 600 python bench.py --workload C4 --steps 20 --warmup 5 > gpurun_out/r2j_bench_c4.json 2> gpurun_out/r2j_bench_c4.err; echo "c4 rc=$?" >> gpurun_out/r2j_rc.txt
timeout 600 python bench.py --workload C5 --steps 20 --warmup 5 > gpurun_out/r2j_bench_c5.json 2> gpurun_out/r2j_bench_c5.err; echo "c5 rc=$?" >> gpurun_out/r2j_rc.txt
cat gpurun_out/r2j_rc.txt; tail -3 gpurun_out/r2j_tests.log; cat gpurun_out/r2j_mlp32_time.log
