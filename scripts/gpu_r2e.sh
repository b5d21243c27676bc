#!/bin/bash
# round 2, GPU call E (2 GPUs): multi-rank bench line with the training_step block, C5 on 2 GPUs, fit / GoodCorresNet tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fit_gpu.py tests/test_goodcorresnet.py tests/test_mlp32_gpu.py -m gpu -q --timeout 180 > gpurun_out/r2e_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r2e_rc.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2e_bench_2gpu.json 2> gpurun_out/r2e_bench_2gpu.err; echo "bench 2gpu rc=$?" >> gpurun_out/r2e_rc.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2e_bench_2gpu_ref.json 2> gpurun_out/r2e_bench_2gpu_ref.err; echo "bench ref 2gpu rc=$?" >> gpurun_out/r2e_rc.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload C5 --steps 10 --warmup 3 > gpurun_out/r2e_bench_c5_2gpu.json 2> gpurun_out/r2e_bench_c5_2gpu.err; echo "c5 2gpu rc=$?" >> gpurun_out/r2e_rc.txt
cat gpurun_out/r2e_rc.txt; tail -3 gpurun_out/r2e_tests.log; tail -c 1500 gpurun_out/r2e_bench_2gpu.json; tail -5 gpurun_out/r2e_bench_2gpu.err
