#!/bin/bash
# round 2, GPU call L (1 GPU): graphed training step test + C5 with / without the graph
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_graph_step_gpu.py tests/test_model_gpu.py tests/test_losses_gpu.py -m gpu -q --timeout 300 > gpurun_out/r2l_pytest.log 2>&1; echo "tests rc=$?" > gpurun_out/r2l_rc.txt
timeout 300 python bench.py --workload C5 --steps 20 --warmup 5 > gpurun_out/r2l_c5_graph.json 2> gpurun_out/r2l_c5_graph.err; echo "c5 graph rc=$?" >> gpurun_out/r2l_rc.txt
timeout 300 python bench.py --workload C5 --steps 20 --warmup 5 --no-graph > gpurun_out/r2l_c5_eager.json 2> gpurun_out/r2l_c5_eager.err; echo "c5 eager rc=$?" >> gpurun_out/r2l_rc.txt
cat gpurun_out/r2l_rc.txt; tail -15 gpurun_out/r2l_pytest.log; cat gpurun_out/r2l_c5_graph.json | cut -c1-1500; tail -3 gpurun_out/r2l_c5_graph.err
