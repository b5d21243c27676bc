#!/bin/bash
# round 2, GPU call I (1 GPU): backward latency kernel + warp-per-item pose backward: tests, bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fit_bwd_gpu.py tests/test_pose_gpu.py tests/test_model_gpu.py tests/test_reference_callers.py -m gpu -q -s --timeout 180 > gpurun_out/r2i_tests.log 2>&1; echo "bwd tests rc=$?" > gpurun_out/r2i_rc.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r2i_pytest_gpu.log 2>&1; echo "gpu suite rc=$?" >> gpurun_out/r2i_rc.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "bench rc=$?" >> gpurun_out/r2i_rc.txt
timeout 600 python bench.py --workload C5 --steps 20 --warmup 5 > gpurun_out/r2i_bench_c5.json 2> gpurun_out/r2i_bench_c5.err; echo "c5 rc=$?" >> gpurun_out/r2i_rc.txt
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_fit_bwd_gpu.py -m gpu -q -x --timeout 600 -k "test_against_oracle_autograd and all and (333 or 37)" > gpurun_out/r2i_sanitizer_memcheck_bwd.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2i_rc.txt
cat gpurun_out/r2i_rc.txt; tail -3 gpurun_out/r2i_tests.log; tail -3 gpurun_out/r2i_pytest_gpu.log
