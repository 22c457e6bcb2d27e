#!/bin/bash
cd "$(dirname "$0")/.."
python - <<'P'
from tools import fqgen
r1,r2=fqgen.generate(4760000,seed=2,paired=True)
r1.tofile('/dev/shm/a.fq'); r2.tofile('/dev/shm/b.fq')
s1=fqgen.truncate_reads(r1,600000); s2=fqgen.truncate_reads(r2,600000)
s1.tofile('/dev/shm/sa.fq'); s2.tofile('/dev/shm/sb.fq')
P
for f in sa a; do
 echo "== $f"; RPQ_CLI_TIMING=1 repaq_b200/repaq_b200_cli -c -i /dev/shm/$f.fq -I /dev/shm/${f/a/b}.fq -o /dev/shm/$f.rfq 2>&1 | tail -30
 time repaq_b200/repaq_b200_cli -d -i /dev/shm/$f.rfq -o /dev/shm/d1.fq -O /dev/shm/d2.fq
done
