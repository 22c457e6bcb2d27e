#!/bin/bash
# round 2, run D: the full default bench line at N=1 (timed by the shell as the driver would)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
WANT=$(cat $(ls repaq_b200/csrc/*.cu repaq_b200/csrc/*.cuh repaq_b200/csrc/*.h repaq_b200/csrc/*.inc repaq_b200/csrc/*.cpp include/repaq_b200.h | sort) | sha1sum | cut -c1-40)
if [ "$WANT" != "$(cat repaq_b200/.build_stamp 2>/dev/null)" ]; then echo "STALE BUILD"; exit 9; fi
nproc; free -g | head -2
echo "== bench full"; T0=$(date +%s); timeout 1500 python bench.py "$@" > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err; echo "rc=$? seconds=$(( $(date +%s) - T0 ))"; tail -5 gpurun_out/bench_full.err | cut -c1-300; tail -c 6000 gpurun_out/bench_full.log
