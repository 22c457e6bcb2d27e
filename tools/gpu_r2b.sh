#!/bin/bash
# round 2, run B: --set full + source page of chosen kernels at 0.43 GB; only CSV summaries come back (the .ncu-rep stays on the box)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for K in ${NCU_KERNELS:-k_dec_format4 k_dec_qindex k_streams4 k_meta3}; do
  echo "== ncu full $K"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:^$K\$ -s 3 -c 1 -f -o /tmp/prof_$K python bench.py --pairs 600000 --steps 1 --warmup 3 --no-e2e --no-cpu --no-roofline --no-extra > gpurun_out/ncu_$K.log 2>&1; echo "rc=$?"
  ncu -i /tmp/prof_$K.ncu-rep --page raw --csv > gpurun_out/raw_$K.csv 2>/dev/null
  python tools/ncu_summary.py gpurun_out/raw_$K.csv > gpurun_out/sum_$K.txt 2>&1
  python tools/ncu_hot_lines.py /tmp/prof_$K.ncu-rep $K 60 > gpurun_out/hot_$K.txt 2>&1
  ls -la /tmp/prof_$K.ncu-rep
done
du -sh gpurun_out
