#!/bin/bash
# usage (on the GPU box): scripts/ab_lib.sh NAME...   device-resident bench of build/variants/NAME.so (burst + 1 s sustained)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for n in "$@"; do
  timeout 300 env PB200_LIB_PATH=$PWD/build/variants/$n.so python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-e2e --no-extras --sustained-s 1.0 $ABARGS 2>gpurun_out/ab_$n.err \
    | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$n', round(d['value']/1e3,1), 'Gpx/s frac', round(d['roofline']['frac'],4), 'sustained', round(d['roofline']['sustained']['frac'],4), d['roofline']['sustained']['clocks']['sm_mhz'])" | tee -a gpurun_out/ab.log
done
