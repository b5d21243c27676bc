#!/bin/bash
# round 2, GPU call V (N GPUs, N = $1): contract line and the C5 training step at N GPUs
N=$1
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2v_bench_${N}gpu.json 2> gpurun_out/r2v_bench_${N}gpu.err; echo "bench ${N}gpu rc=$?" > gpurun_out/r2v_rc_$N.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --workload C5 --steps 20 --warmup 5 > gpurun_out/r2v_bench_c5_${N}gpu.json 2> gpurun_out/r2v_bench_c5_${N}gpu.err; echo "c5 ${N}gpu rc=$?" >> gpurun_out/r2v_rc_$N.txt
cat gpurun_out/r2v_rc_$N.txt; wc -c gpurun_out/r2v_bench_${N}gpu.json gpurun_out/r2v_bench_c5_${N}gpu.json
