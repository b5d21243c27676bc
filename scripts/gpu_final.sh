#!/bin/bash
# end-of-round verification (1 GPU): whole GPU suite, smoke, contract line, reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/final_pytest_gpu.log 2>&1; echo "gpu suite rc=$?" > gpurun_out/final_rc.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/final_rc.txt
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?" >> gpurun_out/final_rc.txt
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; echo "ref rc=$?" >> gpurun_out/final_rc.txt
cat gpurun_out/final_rc.txt; tail -n 3 gpurun_out/final_pytest_gpu.log; tail -n 1 gpurun_out/final_smoke.log; cut -c1-300 gpurun_out/final_bench.json
