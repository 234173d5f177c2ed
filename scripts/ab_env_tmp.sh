cd /root/repo
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e --sustained-s 0 2>gpurun_out/tmp_bench.err | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); a=d['adversarial_worst_case']; print('value', round(d['value']/1e3,1), 'adversarial', round(a['value']/1e3,1), 'config0', round(d['config0_l30']['value']/1e3,1), d['config0_l30'].get('roofline_frac_this_rank'), 'mosaic', round(d['mosaic']['value']/1e3,1))"
tail -3 gpurun_out/tmp_bench.err
