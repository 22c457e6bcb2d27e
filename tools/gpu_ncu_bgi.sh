#!/bin/bash
# ncu --set full + source of chosen kernels on the BGI shape (tools/shape_probe.py); CSV summaries only
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for K in ${NCU_KERNELS:-k_streams6}; do
  echo "== ncu full $K"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:^$K\$ -s 2 -c 1 -f -o /tmp/prof_$K python tools/shape_probe.py 2000000 bgi > gpurun_out/ncu_$K.log 2>&1; echo "rc=$?"
  ncu -i /tmp/prof_$K.ncu-rep --page raw --csv > gpurun_out/raw_bgi_$K.csv 2>/dev/null
  python tools/ncu_summary.py gpurun_out/raw_bgi_$K.csv > gpurun_out/sum_bgi_$K.txt 2>&1
  python tools/ncu_hot_lines.py /tmp/prof_$K.ncu-rep $K 45 > gpurun_out/hot_bgi_$K.txt 2>&1
done
