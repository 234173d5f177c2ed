#!/bin/bash
# compute-sanitizer over the kernels' stress tests (one gpurun call); logs -> gpurun_out/sanitizer_*.log
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SEL='stream or mixed or cover_mode or seeded or fixture or one_shot or reuses or many_small'
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 99 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool exit $?" | tee -a gpurun_out/sanitizer_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer_$tool.log | tail -3
done
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 99 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/sanitizer_smoke.log
