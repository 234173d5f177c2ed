#!/bin/bash
# Evidence run on the GPU box (one gpurun call): parity tests, smoke, both bench arms, launch list, one full ncu
# capture of the dominant kernel.  Everything lands in gpurun_out/.   usage: scripts/gpu_round.sh TAG [quick]
tag=${1:-r2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv | tee gpurun_out/gpu_$tag.txt
lscpu | grep -E "^CPU\(s\)|Model name|NUMA node" | tee -a gpurun_out/gpu_$tag.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_$tag.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/smoke_$tag.log
# bench logic on tiny sizes first: a crash must not cost the full run's time
timeout 600 python bench.py --size 512 --tiles 2 --mosaic-size 2048 --series 9 --steps 5 --sustained-s 0.2 \
    > gpurun_out/bench_${tag}_tiny.json 2> gpurun_out/bench_${tag}_tiny.err || tail -20 gpurun_out/bench_${tag}_tiny.err
tail -c 600 gpurun_out/bench_${tag}_tiny.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_${tag}_ref.json 2> gpurun_out/bench_${tag}_ref.err
timeout 900 python bench.py > gpurun_out/bench_${tag}_n1.json 2> gpurun_out/bench_${tag}_n1.err || tail -20 gpurun_out/bench_${tag}_n1.err
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_${tag}_n1.json').read().strip().splitlines()[-1])
    print('value', d['value'], 'frac', d['roofline']['frac'], 'sustained', (d['roofline'].get('sustained') or {}).get('frac'),
          'e2e', d['e2e'] and (d['e2e']['value'], d['e2e']['ms_per_tile'], d['e2e'].get('ceiling_ms')), 'parity', d.get('parity'))
    for k in ("mosaic", "batch64_strong", "timeseries365", "adversarial_worst_case", "swath_edge_tiles", "config0_l30"):
        v = d.get(k) or {}
        print(k, v.get('value'), v.get('error'), v.get('clocks'))
except Exception as e:
    print('bench line unreadable', e)
PY
if [ "$2" != "quick" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --sustained-s 0 > gpurun_out/launches_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dswx_fused_stream -s 3 -c 1 -f -o gpurun_out/prof_$tag \
    python bench.py --steps 3 --warmup 3 --tiles 4 --no-cpu-baseline --no-e2e --no-extras --sustained-s 0 > gpurun_out/ncu_$tag.log 2>&1
fi
ls -la gpurun_out | tail -8
