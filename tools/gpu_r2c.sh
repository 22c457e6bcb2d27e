#!/bin/bash
# round 2, run C: quick check: smoke, GPU tests (-x), device bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
WANT=$(cat $(ls repaq_b200/csrc/*.cu repaq_b200/csrc/*.cuh repaq_b200/csrc/*.h repaq_b200/csrc/*.inc repaq_b200/csrc/*.cpp include/repaq_b200.h | sort) | sha1sum | cut -c1-40)
if [ "$WANT" != "$(cat repaq_b200/.build_stamp 2>/dev/null)" ]; then echo "STALE BUILD"; exit 9; fi
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
echo "== bench (device only)"; timeout 600 python bench.py --no-e2e --no-cpu --no-extra > gpurun_out/bench_dev.log 2>&1; echo "rc=$?"; tail -c 1500 gpurun_out/bench_dev.log
if [ -z "$NO_TESTS" ]; then echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log; fi
