#!/bin/bash
# validation of a build on a B200 box: smoke, all GPU tests, the full default bench line (logs -> gpurun_out/)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
WANT=$(cat $(ls repaq_b200/csrc/*.cu repaq_b200/csrc/*.cuh repaq_b200/csrc/*.h repaq_b200/csrc/*.inc repaq_b200/csrc/*.cpp include/repaq_b200.h | sort) | sha1sum | cut -c1-40)
if [ "$WANT" != "$(cat repaq_b200/.build_stamp 2>/dev/null)" ]; then echo "STALE BUILD"; exit 9; fi
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
echo "== bench full"; T0=$(date +%s); timeout 1500 python bench.py "$@" > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err; echo "rc=$? seconds=$(( $(date +%s) - T0 ))"; tail -5 gpurun_out/bench_full.err | cut -c1-300; tail -c 2500 gpurun_out/bench_full.log
