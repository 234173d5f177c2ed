#!/bin/bash
# Roofline curve over the batch size: tiles per launch 1..64, time-series workload
# (acquisitions of one tile share DEM / LAND / ocean).  usage (GPU box): scripts/batch_sweep.sh > gpurun_out/batch_sweep.json
cd "$(dirname "$0")/.."
echo "["
first=1
for n in 1 2 4 8 16 32 64; do
  line=$(python bench.py --workload timeseries --tiles $n --steps 40 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 |
         python -c "import json,sys; d=json.loads(sys.stdin.read()); print(json.dumps({'tiles_per_launch': $n, 'Gpixel_per_s': round(d['value']/1e3,1), 'ms_per_launch': round(d['ms_per_step'],4), 'frac_of_hbm_peak': round(d['roofline']['frac'],3)}))")
  [ $first -eq 0 ] && echo ","
  first=0
  echo "  $line"
done
echo "]"
