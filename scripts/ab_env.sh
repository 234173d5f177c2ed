#!/bin/bash
# usage (on the GPU box): scripts/ab_env.sh "NAME=ENV..." ...   quick device-resident bench per environment setting
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for spec in "$@"; do
  name=${spec%%:*}; envs=${spec#*:}
  timeout 300 env $envs python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-e2e --no-extras --sustained-s 1.0 $ABARGS 2>gpurun_out/ab_$name.err \
    | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$name', round(d['value']/1e3,1), 'Gpx/s frac', round(d['roofline']['frac'],4), 'sustained', round(d['roofline']['sustained']['frac'],4), d['roofline']['sustained']['clocks']['sm_mhz'], d['clocks'])" | tee -a gpurun_out/ab.log
done
