#!/bin/bash
# round 2, GPU call S (1 GPU): launch list of one eager C5 training step; graphed-step tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_graph_step_gpu.py -m gpu -q -s --timeout 300 > gpurun_out/r2s_pytest.log 2>&1; echo "tests rc=$?" > gpurun_out/r2s_rc.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2s_c5_launches.csv python bench.py --workload C5 --steps 3 --warmup 3 --no-graph --mlp tc32 > gpurun_out/r2s_c5_ncu.log 2>&1; echo "ncu rc=$?" >> gpurun_out/r2s_rc.txt
cat gpurun_out/r2s_rc.txt; grep -n "bf16 replay\|passed\|failed" gpurun_out/r2s_pytest.log; wc -l gpurun_out/r2s_c5_launches.csv
