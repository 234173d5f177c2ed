#!/bin/bash
# usage (on the GPU box): scripts/ab_bench.sh NAME...  - quick device-resident bench of each build/variants/NAME.so
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for n in "$@"; do
  timeout 120 env PB200_LIB_PATH=$PWD/build/variants/$n.so python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-e2e $ABARGS 2>gpurun_out/ab_$n.err \
    | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$n', d['value'], d['unit'], 'frac', d['roofline']['frac'], 'ms', d['ms_per_step'])" | tee -a gpurun_out/ab.log
done
