#!/bin/bash
# One ncu session over EVERY kernel of one bench step:
#   1. full-size (3.4 GB) launch list with DRAM byte counters   -> gpurun_out/traffic_full.csv   (roofline.traffic, kernel shares)
#   2. `--set full` capture of one step at 0.43 GB               -> gpurun_out/prof_all.ncu-rep  (per-kernel analysis, source pages)
# plus an unprofiled bench line first (numbers under ncu are never bench values).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
WANT=$(cat $(ls repaq_b200/csrc/*.cu repaq_b200/csrc/*.cuh repaq_b200/csrc/*.h repaq_b200/csrc/*.inc repaq_b200/csrc/*.cpp include/repaq_b200.h | sort) | sha1sum | cut -c1-40)
if [ "$WANT" != "$(cat repaq_b200/.build_stamp 2>/dev/null)" ]; then echo "STALE BUILD"; exit 9; fi
echo "== bench (device only)"; timeout 600 python bench.py --no-e2e --no-cpu --no-extra > gpurun_out/bench_dev.log 2>&1; echo "rc=$?"; tail -c 1500 gpurun_out/bench_dev.log
# ~45 launches of our kernels per step (k_fetch read-backs included); the window below holds at least one whole step after the
# gate step and two warm-up steps (tools/ncu_traffic.py cuts one step out of it)
SKIP=${SKIP:-130}; CNT=${CNT:-110}
echo "== ncu dram counters, full size"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:^k_ -s $SKIP -c $CNT --csv --log-file gpurun_out/traffic_full.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-roofline --no-extra > gpurun_out/ncu_traffic.log 2>&1; echo "rc=$?"
echo "== ncu --set full, every kernel of one step (0.43 GB)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:^k_ -s $SKIP -c $CNT -f -o gpurun_out/prof_all \
    python bench.py --pairs 600000 --steps 1 --warmup 3 --no-e2e --no-cpu --no-roofline --no-extra > gpurun_out/ncu_all.log 2>&1; echo "rc=$?"
ls -la gpurun_out/prof_all.ncu-rep
