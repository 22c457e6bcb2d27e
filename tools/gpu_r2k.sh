#!/bin/bash
# round 2, k_streams7: selected GPU parity tests, both shapes with and without the new coder, the device-resident bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
WANT=$(cat $(ls repaq_b200/csrc/*.cu repaq_b200/csrc/*.cuh repaq_b200/csrc/*.h repaq_b200/csrc/*.inc repaq_b200/csrc/*.cpp include/repaq_b200.h | sort) | sha1sum | cut -c1-40)
if [ "$WANT" != "$(cat repaq_b200/.build_stamp 2>/dev/null)" ]; then echo "STALE BUILD"; exit 9; fi
echo "== pytest gpu (subset)"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "${TESTS:-oracle or adversarial or n_positions or dense or golden}" > gpurun_out/pytest_gpu_subset.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_subset.log
echo "== shapes (k_streams7)"; timeout 600 python tools/shape_probe.py 8000000 > gpurun_out/shapes_s7.log 2>&1; echo "rc=$?"; cut -c1-900 gpurun_out/shapes_s7.log
echo "== shapes (k_streams4/6)"; RPQ_DEBUG_STREAMS7=0 timeout 600 python tools/shape_probe.py 8000000 > gpurun_out/shapes_s46.log 2>&1; echo "rc=$?"; cut -c1-900 gpurun_out/shapes_s46.log
echo "== bench (device only)"; timeout 900 python bench.py --no-e2e --no-cpu --no-extra > gpurun_out/bench_dev.log 2>&1; echo "rc=$?"; tail -c 1800 gpurun_out/bench_dev.log
