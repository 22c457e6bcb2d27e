#!/bin/bash
# A/B of one environment knob on the device-resident bench: tools/gpu_ab_env.sh NAME v1 v2 ...   (prints value, encode/decode GB/s, top kernels)
cd "$(dirname "$0")/.."
NAME=$1; shift
for V in "$@"; do
  echo "== $NAME=$V"
  env $NAME=$V python bench.py --no-e2e --no-cpu --no-extra --steps 3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['roofline']['kernels_ms']
print(round(d['value'],1), round(d['encode_gbs'],1), round(d['decode_gbs'],1), dict(list(k.items())[:9]))"
done
