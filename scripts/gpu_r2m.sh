#!/bin/bash
# round 2, GPU call M (1 GPU): split pipeline launch modes; graphed training step test
mkdir -p gpurun_out
timeout 600 python scripts/split_pipe_time.py > gpurun_out/r2m_split_pipe.txt 2>&1; echo "pipe rc=$?" > gpurun_out/r2m_rc.txt
timeout 600 python -m pytest tests/test_graph_step_gpu.py tests/test_model_gpu.py -m gpu -q --timeout 300 > gpurun_out/r2m_pytest.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2m_rc.txt
cat gpurun_out/r2m_rc.txt; cat gpurun_out/r2m_split_pipe.txt; tail -15 gpurun_out/r2m_pytest.log
