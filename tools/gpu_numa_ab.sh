#!/bin/bash
# N-GPU e2e with and without binding every rank to the CPUs next to its GPU.  usage: gpu_numa_ab.sh N
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-4}
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
for MODE in 1 0; do
  RPQ_BENCH_NUMA=$MODE timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$MODE bench.py --gpus $N --no-cpu --no-roofline > gpurun_out/numa_${MODE}_n$N.log 2>&1
  echo "numa=$MODE rc=$?"
  python - gpurun_out/numa_${MODE}_n$N.log <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l); e = d['e2e']
        print('value %.1f e2e %.2f (%.1f ms; enc %s dec %s) serial %.2f affinity %s' % (d['value'], e['value'], e['ms_per_step'], e['encode_call_ms'], e['decode_call_ms'], e['serial']['value'], d['config'].get('host_affinity')))
PY
done
head -12 gpurun_out/topo_$N.txt
