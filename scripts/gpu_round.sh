#!/bin/bash
# Round-end evidence run on the GPU box (one gpurun call): parity tests, smoke, both bench arms, launch list, one
# full ncu capture of the dominant kernel.  Everything lands in gpurun_out/.   usage: scripts/gpu_round.sh TAG
tag=${1:-r1}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_$tag.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/smoke_$tag.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${tag}_ref.json 2> gpurun_out/bench_${tag}_ref.err
python bench.py > gpurun_out/bench_${tag}_n1.json 2> gpurun_out/bench_${tag}_n1.err
tail -c 2500 gpurun_out/bench_${tag}_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dswx_fused_fast -s 3 -c 1 -f -o gpurun_out/prof_$tag \
    python bench.py --steps 3 --warmup 3 --tiles 4 --no-cpu-baseline --no-e2e > gpurun_out/ncu_$tag.log 2>&1
ls -la gpurun_out | tail -12
