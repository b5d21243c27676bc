#!/bin/bash
# round 2, GPU call W (N GPUs, N = $1): C5 training step only
N=$1
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --workload C5 --steps 20 --warmup 5 > gpurun_out/r3f_bench_c5_${N}gpu.json 2> gpurun_out/r3f_bench_c5_${N}gpu.err; echo "c5 ${N}gpu rc=$?" > gpurun_out/r3f_rc_$N.txt
cat gpurun_out/r3f_rc_$N.txt
