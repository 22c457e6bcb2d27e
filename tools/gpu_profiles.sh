#!/bin/bash
# profiles of a build: the ncu launch list of one full-size step with DRAM byte counters (-> profiles/r02_traffic.json, kernel shares),
# `--set full` captures with source of the top kernels (NovaSeq-shape bench step at 0.43 GB; BGI-shape probe), CSV summaries only
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
WANT=$(cat $(ls repaq_b200/csrc/*.cu repaq_b200/csrc/*.cuh repaq_b200/csrc/*.h repaq_b200/csrc/*.inc repaq_b200/csrc/*.cpp include/repaq_b200.h | sort) | sha1sum | cut -c1-40)
if [ "$WANT" != "$(cat repaq_b200/.build_stamp 2>/dev/null)" ]; then echo "STALE BUILD"; exit 9; fi
echo "== ncu dram counters, full size"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:^k_ -s ${SKIP:-160} -c ${CNT:-130} --csv --log-file gpurun_out/traffic_full.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-roofline --no-extra > gpurun_out/ncu_traffic.log 2>&1; echo "rc=$?"
python tools/ncu_traffic.py gpurun_out/traffic_full.csv gpurun_out/r02_traffic.json 4758618 | head -40
for K in ${NCU_KERNELS:-k_streams4 k_meta3 k_dec_format4 k_index_lines k_dec_planes k_emit2}; do
  echo "== ncu full $K"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:^$K\$ -s 0 -c 1 -f -o /tmp/prof_$K python bench.py --pairs 600000 --steps 1 --warmup 3 --no-e2e --no-cpu --no-roofline --no-extra > gpurun_out/ncu_$K.log 2>&1; echo "rc=$?"
  ncu -i /tmp/prof_$K.ncu-rep --page raw --csv > gpurun_out/raw_$K.csv 2>/dev/null
  python tools/ncu_summary.py gpurun_out/raw_$K.csv > gpurun_out/sum_$K.txt 2>&1
  python tools/ncu_hot_lines.py /tmp/prof_$K.ncu-rep $K 45 > gpurun_out/hot_$K.txt 2>&1
done
for K in ${NCU_BGI:-k_streams7 k_dec_planes k_dec_qindex}; do
  echo "== ncu full $K (BGI shape)"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:^$K\$ -s 2 -c 1 -f -o /tmp/prof_bgi_$K python tools/shape_probe.py 2000000 bgi > gpurun_out/ncu_bgi_$K.log 2>&1; echo "rc=$?"
  ncu -i /tmp/prof_bgi_$K.ncu-rep --page raw --csv > gpurun_out/raw_bgi_$K.csv 2>/dev/null
  python tools/ncu_summary.py gpurun_out/raw_bgi_$K.csv > gpurun_out/sum_bgi_$K.txt 2>&1
  python tools/ncu_hot_lines.py /tmp/prof_bgi_$K.ncu-rep $K 45 > gpurun_out/hot_bgi_$K.txt 2>&1
done
echo "== shapes"; timeout 600 python tools/shape_probe.py 8000000 > gpurun_out/shapes.log 2>&1; echo "rc=$?"; cut -c1-600 gpurun_out/shapes.log
