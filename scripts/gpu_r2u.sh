#!/bin/bash
# round 2, GPU call U (1 GPU): C5 with fused Adam; launch list of one eager C5 step after the small-batch changes
mkdir -p gpurun_out
timeout 300 python bench.py --workload C5 --steps 20 --warmup 5 > gpurun_out/r2u_c5_graph.json 2> gpurun_out/r2u_c5_graph.err; echo "c5 graph rc=$?" > gpurun_out/r2u_rc.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2u_c5_launches.csv python bench.py --workload C5 --steps 1 --warmup 3 --no-graph --no-extras > gpurun_out/r2u_c5_ncu.log 2>&1; echo "ncu rc=$?" >> gpurun_out/r2u_rc.txt
cat gpurun_out/r2u_rc.txt; wc -l gpurun_out/r2u_c5_launches.csv
