#!/bin/bash
# A/B runs of the device-resident bench under environment knobs, no tests.  usage: gpu_ab_dev.sh "VAR=val VAR=val" ...   ("-" = no knob)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
WANT=$(cat $(ls repaq_b200/csrc/*.cu repaq_b200/csrc/*.cuh repaq_b200/csrc/*.h repaq_b200/csrc/*.inc repaq_b200/csrc/*.cpp include/repaq_b200.h | sort) | sha1sum | cut -c1-40)
if [ "$WANT" != "$(cat repaq_b200/.build_stamp 2>/dev/null)" ]; then echo "STALE BUILD"; exit 9; fi
i=0
for KV in "$@"; do
  i=$((i+1))
  echo "== bench [$KV]"
  if [ "$KV" = "-" ]; then timeout 900 python bench.py --no-e2e --no-cpu ${BENCH_ARGS} > gpurun_out/abd_$i.log 2>&1
  else env $KV timeout 900 python bench.py --no-e2e --no-cpu ${BENCH_ARGS} > gpurun_out/abd_$i.log 2>&1; fi
  echo "rc=$? $KV" >> gpurun_out/abd_$i.log
  python - gpurun_out/abd_$i.log <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l); k = d['roofline']['kernels_ms']
        print('value %.1f enc %.1f dec %.1f ms/step %.2f | %s' % (d['value'], d['encode_gbs'], d['decode_gbs'], d['ms_per_step'], ' '.join('%s=%.2f' % (a.replace('k_', ''), b) for a, b in k.items() if b > 0.12)))
    elif 'rror' in l or 'rc=' in l:
        print(l.strip()[:300])
PY
done
