#!/bin/bash
# round 2, GPU call 3b (1 GPU): C5 after the wgrad change; compute-sanitizer over the kernels changed late in the round
mkdir -p gpurun_out
timeout 300 python bench.py --workload C5 --steps 20 --warmup 5 --no-extras > gpurun_out/r3b_c5.json 2> gpurun_out/r3b_c5.err; echo "c5 rc=$?" > gpurun_out/r3b_rc.txt
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_mlp32_gpu.py -m gpu -q -x --timeout 800 -k "last_layer or normbwd or (wgrad and 1-128) or (wgrad and 256-768) or (wgrad and 4-1000) or gradients_are" > gpurun_out/r3b_memcheck.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/r3b_rc.txt
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_mlp32_gpu.py -m gpu -q -x --timeout 800 -k "last_layer or (wgrad and 256-768)" > gpurun_out/r3b_racecheck.txt 2>&1; echo "racecheck rc=$?" >> gpurun_out/r3b_rc.txt
cat gpurun_out/r3b_rc.txt; tail -n 4 gpurun_out/r3b_memcheck.txt; tail -n 4 gpurun_out/r3b_racecheck.txt
