#!/bin/bash
# quick look at a build: selected GPU parity tests, both shapes, the device-resident bench line (and optional ncu captures: NCU_KERNELS="k_a k_b")
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
WANT=$(cat $(ls repaq_b200/csrc/*.cu repaq_b200/csrc/*.cuh repaq_b200/csrc/*.h repaq_b200/csrc/*.inc repaq_b200/csrc/*.cpp include/repaq_b200.h | sort) | sha1sum | cut -c1-40)
if [ "$WANT" != "$(cat repaq_b200/.build_stamp 2>/dev/null)" ]; then echo "STALE BUILD"; exit 9; fi
echo "== pytest gpu (subset)"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "${TESTS:-oracle or adversarial or n_positions or dense or golden or crlf or edge}" > gpurun_out/pytest_gpu_subset.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_subset.log
echo "== shapes"; timeout 600 python tools/shape_probe.py 8000000 > gpurun_out/shapes.log 2>&1; echo "rc=$?"; cut -c1-900 gpurun_out/shapes.log
echo "== bench (device only)"; timeout 900 python bench.py --no-e2e --no-cpu --no-extra > gpurun_out/bench_dev.log 2>&1; echo "rc=$?"; tail -c 1500 gpurun_out/bench_dev.log
for K in $NCU_KERNELS; do
  echo "== ncu full $K"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:^$K\$ -s ${NCU_SKIP:-2} -c 1 -f -o /tmp/prof_$K python bench.py --pairs 600000 --steps 1 --warmup 3 --no-e2e --no-cpu --no-roofline --no-extra > gpurun_out/ncu_$K.log 2>&1; echo "rc=$?"
  ncu -i /tmp/prof_$K.ncu-rep --page raw --csv > gpurun_out/raw_$K.csv 2>/dev/null
  python tools/ncu_summary.py gpurun_out/raw_$K.csv > gpurun_out/sum_$K.txt 2>&1
  python tools/ncu_hot_lines.py /tmp/prof_$K.ncu-rep $K 45 > gpurun_out/hot_$K.txt 2>&1
done
