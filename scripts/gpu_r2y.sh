#!/bin/bash
# round 2, GPU call Y (1 GPU): smoke, contract line, reference arm, C4, C3
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2y_smoke.log 2>&1; echo "smoke rc=$?" > gpurun_out/r2y_rc.txt
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; echo "bench rc=$?" >> gpurun_out/r2y_rc.txt
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 > gpurun_out/r2y_bench_ref.json 2> gpurun_out/r2y_bench_ref.err; echo "ref rc=$?" >> gpurun_out/r2y_rc.txt
timeout 300 python bench.py --workload C4 --steps 10 --warmup 3 > gpurun_out/r2y_c4.json 2> gpurun_out/r2y_c4.err; echo "c4 rc=$?" >> gpurun_out/r2y_rc.txt
timeout 300 python bench.py --workload C3 --steps 20 --warmup 5 > gpurun_out/r2y_c3.json 2> gpurun_out/r2y_c3.err; echo "c3 rc=$?" >> gpurun_out/r2y_rc.txt
cat gpurun_out/r2y_rc.txt; tail -n 3 gpurun_out/r2y_smoke.log
