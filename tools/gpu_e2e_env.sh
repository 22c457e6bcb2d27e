#!/bin/bash
# A/B of one environment knob on the end-to-end leg of the bench: tools/gpu_e2e_env.sh NAME v1 v2 ...  ("-" = unset)
cd "$(dirname "$0")/.."
NAME=$1; shift
for V in "$@"; do
  echo "== $NAME=$V"
  if [ "$V" == "-" ]; then E=""; else E="$NAME=$V"; fi
  env $E python bench.py --no-cpu --no-extra --no-roofline --steps ${STEPS:-10} 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']
print('e2e', round(e['value'],2), 'ms/step', round(e['ms_per_step'],1), 'enc call', e['encode_call_ms'], 'dec call', e['decode_call_ms'], 'serial', round(e['serial']['value'],2))"
done
