#!/bin/bash
# The measurements one round's profiles/ entries are made from: smoke, GPU tests, the full bench line, the reference arm, the other
# shapes, the DRAM-traffic launch list at full size and `--set full` captures of the top kernels.  Logs -> gpurun_out/
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
WANT=$(cat $(ls repaq_b200/csrc/*.cu repaq_b200/csrc/*.cuh repaq_b200/csrc/*.h repaq_b200/csrc/*.inc repaq_b200/csrc/*.cpp include/repaq_b200.h | sort) | sha1sum | cut -c1-40)
if [ "$WANT" != "$(cat repaq_b200/.build_stamp 2>/dev/null)" ]; then echo "STALE BUILD"; exit 9; fi
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt; free -g >> gpurun_out/nproc.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== bench full"; timeout 1500 python bench.py > gpurun_out/bench_full.log 2>&1; echo "bench full rc=$?"; tail -c 700 gpurun_out/bench_full.log
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?"; tail -c 500 gpurun_out/bench_ref.log
echo "== other shapes"; timeout 600 python tools/shape_probe.py 4000000 > gpurun_out/shapes.log 2>&1; echo "rc=$?"; cut -c1-400 gpurun_out/shapes.log
echo "== ncu dram counters, full size"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:^k_ -s ${SKIP:-150} -c ${CNT:-130} --csv --log-file gpurun_out/traffic_full.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-roofline --no-extra > gpurun_out/ncu_traffic.log 2>&1; echo "rc=$?"
for K in ${NCU_KERNELS:-k_streams4 k_dec_format3 k_meta3 k_dec_streams}; do
  echo "== ncu full $K"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/prof_$K python bench.py --pairs 600000 --steps 1 --warmup 3 --no-e2e --no-cpu --no-roofline --no-extra > gpurun_out/ncu_$K.log 2>&1; echo "rc=$?"
done
