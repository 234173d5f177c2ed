cd /root/repo
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stream or mixed or fixture or seeded or guard_band or full_size" 2>&1 | tail -3
for i in 1 2; do
bash scripts/ab_env.sh "dyn:PB200_STREAM_DYNAMIC=1" "static:PB200_STREAM_DYNAMIC=0"
done
