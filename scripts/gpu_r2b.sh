#!/bin/bash
# round 2, GPU call B: mlp32 forward + backward tests, all-layer / reference-callers tests, whole suite, C4 + C5 lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mlp32_gpu.py -m gpu -q -s --timeout 180 > gpurun_out/r2b_mlp32.log 2>&1; echo "mlp32 rc=$?" > gpurun_out/r2b_rc.txt
timeout 600 python -m pytest tests/test_all_layers.py tests/test_reference_callers.py -m gpu -q -s --timeout 180 > gpurun_out/r2b_layers.log 2>&1; echo "layers rc=$?" >> gpurun_out/r2b_rc.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r2b_pytest_gpu.log 2>&1; echo "gpu suite rc=$?" >> gpurun_out/r2b_rc.txt
timeout 600 python bench.py --workload C5 --steps 10 --warmup 3 > gpurun_out/r2b_bench_c5.json 2> gpurun_out/r2b_bench_c5.err; echo "bench c5 rc=$?" >> gpurun_out/r2b_rc.txt
cat gpurun_out/r2b_rc.txt
grep -E "passed|failed" gpurun_out/r2b_mlp32.log | tail -2
