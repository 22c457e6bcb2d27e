#!/bin/bash
# quick GPU session: GPU tests + full bench (+ optional A/B with RPQ_NO_PIPELINE=1)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
echo "== bench full"; timeout 1500 python bench.py > gpurun_out/bench_full.log 2>&1; echo "bench full rc=$?"; tail -c 600 gpurun_out/bench_full.log
if [ "$1" == "ab" ]; then
echo "== bench full (no pipeline)"; RPQ_NO_PIPELINE=1 timeout 1500 python bench.py --no-cpu > gpurun_out/bench_full_nopipe.log 2>&1; echo "rc=$?"; tail -c 300 gpurun_out/bench_full_nopipe.log
fi
