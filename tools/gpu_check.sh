#!/bin/bash
# One GPU-box session: smoke, GPU parity tests, a short and a full bench, the ncu launch list + full captures.  Logs -> gpurun_out/
# usage: gpu_check.sh [quick|full|prof]
cd "$(dirname "$0")/.."
MODE=${1:-full}
mkdir -p gpurun_out
# refuse to measure a stale library: the stamp is the hash of the sources the .so was built from (csrc/Makefile)
WANT=$(cat $(ls repaq_b200/csrc/*.cu repaq_b200/csrc/*.cuh repaq_b200/csrc/*.h repaq_b200/csrc/*.inc repaq_b200/csrc/*.cpp include/repaq_b200.h | sort) | sha1sum | cut -c1-40)
if [ "$WANT" != "$(cat repaq_b200/.build_stamp 2>/dev/null)" ]; then echo "STALE BUILD: librepaq_b200.so does not match the sources; run make -C repaq_b200/csrc"; exit 9; fi
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt; free -g >> gpurun_out/nproc.txt
if [ "$MODE" != "prof" ]; then
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== bench small"; timeout 600 python bench.py --pairs 600000 --steps 2 --warmup 3 > gpurun_out/bench_small.log 2>&1; echo "bench small rc=$?"; tail -c 1500 gpurun_out/bench_small.log
fi
if [ "$MODE" != "quick" ]; then
echo "== bench full"; timeout 1500 python bench.py > gpurun_out/bench_full.log 2>&1; echo "bench full rc=$?"; tail -c 4500 gpurun_out/bench_full.log
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?"; tail -c 1200 gpurun_out/bench_ref.log
echo "== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --pairs 600000 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
for K in ${NCU_KERNELS:-k_streams k_meta k_dec_format k_dec_streams}; do
  echo "== ncu full $K"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/prof_$K python bench.py --pairs 600000 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_$K.log 2>&1; echo "rc=$?"
done
fi
