#!/bin/bash
# round 2, GPU call A: new tests first (each under a timeout), then the whole GPU suite, then bench lines
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 600 python -m pytest tests/test_mlp32_gpu.py -m gpu -x -q -s --timeout 120 > gpurun_out/r2a_mlp32.log 2>&1; echo "mlp32 rc=$?" > gpurun_out/r2a_rc.txt
timeout 600 python -m pytest tests/test_all_layers.py tests/test_reference_callers.py -m gpu -q -s --timeout 180 > gpurun_out/r2a_layers.log 2>&1; echo "layers rc=$?" >> gpurun_out/r2a_rc.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r2a_pytest_gpu.log 2>&1; echo "gpu suite rc=$?" >> gpurun_out/r2a_rc.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?" >> gpurun_out/r2a_rc.txt
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err; echo "bench ref rc=$?" >> gpurun_out/r2a_rc.txt
timeout 600 python bench.py --workload C3 --steps 20 --warmup 5 --no-extras > gpurun_out/r2a_bench_c3.json 2> gpurun_out/r2a_bench_c3.err; echo "bench c3 rc=$?" >> gpurun_out/r2a_rc.txt
timeout 600 python bench.py --workload C4 --steps 10 --warmup 3 > gpurun_out/r2a_bench_c4.json 2> gpurun_out/r2a_bench_c4.err; echo "bench c4 rc=$?" >> gpurun_out/r2a_rc.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2a_rc.txt
cat gpurun_out/r2a_rc.txt
tail -5 gpurun_out/r2a_mlp32.log
