#!/bin/bash
# One GPU-box session: smoke, GPU parity tests, a short and a full bench, the ncu launch list.  Logs -> gpurun_out/
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt; free -g >> gpurun_out/nproc.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== bench small"; timeout 600 python bench.py --pairs 600000 --steps 2 --warmup 3 > gpurun_out/bench_small.log 2>&1; echo "bench small rc=$?"; tail -c 3000 gpurun_out/bench_small.log
if [ "$1" != "quick" ]; then
echo "== bench full"; timeout 1500 python bench.py > gpurun_out/bench_full.log 2>&1; echo "bench full rc=$?"; tail -c 4000 gpurun_out/bench_full.log
echo "== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --pairs 600000 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
fi
