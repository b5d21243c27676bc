#!/bin/bash
# round 2, GPU call T (1 GPU): MLP / model / pose / graph tests after the small-batch kernel changes; C5 graph + eager
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mlp32_gpu.py tests/test_graph_step_gpu.py tests/test_pose_gpu.py tests/test_model_gpu.py tests/test_all_layers.py tests/test_goodcorresnet.py -m gpu -q -x --timeout 300 > gpurun_out/r2t_pytest.log 2>&1; echo "tests rc=$?" > gpurun_out/r2t_rc.txt
timeout 300 python bench.py --workload C5 --steps 20 --warmup 5 > gpurun_out/r2t_c5_graph.json 2> gpurun_out/r2t_c5_graph.err; echo "c5 graph rc=$?" >> gpurun_out/r2t_rc.txt
timeout 300 python bench.py --workload C5 --steps 20 --warmup 5 --no-graph > gpurun_out/r2t_c5_eager.json 2> gpurun_out/r2t_c5_eager.err; echo "c5 eager rc=$?" >> gpurun_out/r2t_rc.txt
cat gpurun_out/r2t_rc.txt; tail -n 12 gpurun_out/r2t_pytest.log; tail -n 3 gpurun_out/r2t_c5_graph.err
