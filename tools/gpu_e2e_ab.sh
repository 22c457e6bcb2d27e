#!/bin/bash
# A/B runs of the full bench (e2e included) under environment knobs.  usage: gpu_e2e_ab.sh "VAR=val VAR=val" ...   ("-" = no knob)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
WANT=$(cat $(ls repaq_b200/csrc/*.cu repaq_b200/csrc/*.cuh repaq_b200/csrc/*.h repaq_b200/csrc/*.inc repaq_b200/csrc/*.cpp include/repaq_b200.h | sort) | sha1sum | cut -c1-40)
if [ "$WANT" != "$(cat repaq_b200/.build_stamp 2>/dev/null)" ]; then echo "STALE BUILD"; exit 9; fi
if [ "${SKIP_TESTS:-0}" != "1" ]; then
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
fi
i=0
for KV in "$@"; do
  i=$((i+1))
  echo "== bench [$KV]"
  if [ "$KV" = "-" ]; then timeout 900 python bench.py --no-cpu ${BENCH_ARGS} > gpurun_out/e2e_$i.log 2>&1
  else env $KV timeout 900 python bench.py --no-cpu ${BENCH_ARGS} > gpurun_out/e2e_$i.log 2>&1; fi
  echo "rc=$? $KV" >> gpurun_out/e2e_$i.log
  python - gpurun_out/e2e_$i.log <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l); e = d.get('e2e') or {}
        print('value %.1f enc %.1f dec %.1f | e2e %.2f (%.1f ms; enc call %s dec call %s) serial %.2f' % (d['value'], d['encode_gbs'], d['decode_gbs'], e.get('value', 0), e.get('ms_per_step', 0), e.get('encode_call_ms'), e.get('decode_call_ms'), (e.get('serial') or {}).get('value', 0)))
    elif 'rror' in l or 'rc=' in l:
        print(l.strip()[:300])
PY
done
