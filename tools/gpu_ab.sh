#!/bin/bash
# A/B runs of the device-resident bench under environment knobs.  usage: gpu_ab.sh "VAR=val" "VAR=val" ...   ("-" = no knob)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
WANT=$(cat $(ls repaq_b200/csrc/*.cu repaq_b200/csrc/*.cuh repaq_b200/csrc/*.h repaq_b200/csrc/*.inc repaq_b200/csrc/*.cpp include/repaq_b200.h | sort) | sha1sum | cut -c1-40)
if [ "$WANT" != "$(cat repaq_b200/.build_stamp 2>/dev/null)" ]; then echo "STALE BUILD"; exit 9; fi
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
i=0
for KV in "$@"; do
  i=$((i+1))
  echo "== bench [$KV]"
  if [ "$KV" = "-" ]; then timeout 900 python bench.py --no-e2e --no-cpu --no-extra > gpurun_out/ab_$i.log 2>&1
  else env $KV timeout 900 python bench.py --no-e2e --no-cpu --no-extra > gpurun_out/ab_$i.log 2>&1; fi
  echo "rc=$? $KV" >> gpurun_out/ab_$i.log
  tail -c 1800 gpurun_out/ab_$i.log
done
