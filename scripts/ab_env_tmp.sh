cd /root/repo
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
bash scripts/ab_lib.sh gen glob gen glob gen glob
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 99 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "many_small or stream or mixed" > gpurun_out/sanitizer_racecheck2.log 2>&1; echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer_racecheck2.log | tail -2
