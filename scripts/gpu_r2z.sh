#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"last_kernel|normbwd|last_bwd|affine" --launch-skip 16 --launch-count 8 -o gpurun_out/r2z_small_batch python scripts/ncu_small_batch_target.py > gpurun_out/r2z_ncu.log 2>&1; echo "ncu rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2z_launches.csv python scripts/ncu_small_batch_target.py > /dev/null 2>&1; echo "list rc=$?"
ls -la gpurun_out/r2z*
