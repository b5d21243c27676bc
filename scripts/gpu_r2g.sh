#!/bin/bash
# round 2, GPU call G (1 GPU): all-layer + mlp32 tests after the cache fix, full bench line with K1 phase cycles
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_all_layers.py tests/test_mlp32_gpu.py tests/test_model_gpu.py -m gpu -q --timeout 180 > gpurun_out/r2g_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r2g_rc.txt
timeout 900 python bench.py --phases > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc=$?" >> gpurun_out/r2g_rc.txt
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2g_bench_ref.json 2> gpurun_out/r2g_bench_ref.err; echo "bench ref rc=$?" >> gpurun_out/r2g_rc.txt
timeout 600 python bench.py --workload C4 --steps 20 --warmup 5 > gpurun_out/r2g_bench_c4.json 2> gpurun_out/r2g_bench_c4.err; echo "c4 rc=$?" >> gpurun_out/r2g_rc.txt
timeout 600 python bench.py --workload C5 --steps 20 --warmup 5 > gpurun_out/r2g_bench_c5.json 2> gpurun_out/r2g_bench_c5.err; echo "c5 rc=$?" >> gpurun_out/r2g_rc.txt
cat gpurun_out/r2g_rc.txt; tail -3 gpurun_out/r2g_tests.log
