#!/bin/bash
# round 2, GPU call K (2 GPUs): whole GPU suite incl. the DataParallel test, 2-GPU contract line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r2r_pytest_gpu.log 2>&1; echo "gpu suite rc=$?" > gpurun_out/r2r_rc.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2r_bench_2gpu.json 2> gpurun_out/r2r_bench_2gpu.err; echo "bench 2gpu rc=$?" >> gpurun_out/r2r_rc.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --workload C5 --steps 20 --warmup 5 > gpurun_out/r2r_bench_c5_2gpu.json 2> gpurun_out/r2r_bench_c5_2gpu.err; echo "c5 2gpu rc=$?" >> gpurun_out/r2r_rc.txt
cat gpurun_out/r2r_rc.txt; tail -4 gpurun_out/r2r_pytest_gpu.log; wc -l gpurun_out/r2r_bench_2gpu.json
