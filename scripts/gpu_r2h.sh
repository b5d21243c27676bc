#!/bin/bash
# round 2, GPU call H (8 GPUs): the contract line at N = 8 (with the training_step block) and the C5 line
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > gpurun_out/r2h_smi.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2h_bench_8gpu.json 2> gpurun_out/r2h_bench_8gpu.err; echo "bench 8gpu rc=$?" > gpurun_out/r2h_rc.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --workload C5 --steps 20 --warmup 5 > gpurun_out/r2h_bench_c5_8gpu.json 2> gpurun_out/r2h_bench_c5_8gpu.err; echo "c5 8gpu rc=$?" >> gpurun_out/r2h_rc.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --workload C5 --steps 20 --warmup 5 > gpurun_out/r2h_bench_c5_4gpu.json 2> gpurun_out/r2h_bench_c5_4gpu.err; echo "c5 4gpu rc=$?" >> gpurun_out/r2h_rc.txt
cat gpurun_out/r2h_rc.txt; head -c 400 gpurun_out/r2h_bench_8gpu.json; tail -3 gpurun_out/r2h_bench_8gpu.err
