#!/bin/bash
# round 2, run A: GPU tests, device bench, launch list with DRAM counters, --set full of every kernel of a step (0.43 GB)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
WANT=$(cat $(ls repaq_b200/csrc/*.cu repaq_b200/csrc/*.cuh repaq_b200/csrc/*.h repaq_b200/csrc/*.inc repaq_b200/csrc/*.cpp include/repaq_b200.h | sort) | sha1sum | cut -c1-40)
if [ "$WANT" != "$(cat repaq_b200/.build_stamp 2>/dev/null)" ]; then echo "STALE BUILD"; exit 9; fi
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt; free -g >> gpurun_out/nproc.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
echo "== bench (device only)"; timeout 600 python bench.py --no-e2e --no-cpu --no-extra > gpurun_out/bench_dev.log 2>&1; echo "rc=$?"; tail -c 1800 gpurun_out/bench_dev.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
SKIP=${SKIP:-130}; CNT=${CNT:-110}
echo "== ncu dram counters, full size"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:^k_ -s $SKIP -c $CNT --csv --log-file gpurun_out/traffic_full.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-roofline --no-extra > gpurun_out/ncu_traffic.log 2>&1; echo "rc=$?"
echo "== ncu --set full, every kernel of one step (0.43 GB)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:^k_ -s $SKIP -c $CNT -f -o gpurun_out/prof_all \
    python bench.py --pairs 600000 --steps 1 --warmup 3 --no-e2e --no-cpu --no-roofline --no-extra > gpurun_out/ncu_all.log 2>&1; echo "rc=$?"
ls -la gpurun_out/prof_all.ncu-rep
