#!/bin/bash
# development aid: contract number (no extras) for library build variants x parallel branches
for v in default "$@"; do
  for s in ${STREAMS:-2 4}; do
    if [ "$v" = default ]; then unset FEPE_B200_LIB; else export FEPE_B200_LIB=$PWD/pytorch-deepfepe_b200/lib/variants/$v.so; fi
    echo -n "$v streams=$s: "
    python bench.py --streams $s --no-extras --steps 2000 --warmup 20 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value %.2fM pairs/s, %.2f us/step, fit launch alone %.2f us' % (d['value']/1e6, d['ms_per_step']*1e3, d['roofline']['launch_us']))"
  done
done
