#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --workload C5 --steps 20 --warmup 5 > gpurun_out/r3e_c5.json 2> gpurun_out/r3e_c5.err; echo "c5 rc=$?" > gpurun_out/r3e_rc.txt
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_nn_match_gpu.py -m gpu -q -x --timeout 500 -k "tc" > gpurun_out/r3e_memcheck.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/r3e_rc.txt
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_nn_match_gpu.py -m gpu -q -x --timeout 500 -k "tc and 129" > gpurun_out/r3e_racecheck.txt 2>&1; echo "racecheck rc=$?" >> gpurun_out/r3e_rc.txt
cat gpurun_out/r3e_rc.txt; tail -n 3 gpurun_out/r3e_memcheck.txt; tail -n 3 gpurun_out/r3e_racecheck.txt
