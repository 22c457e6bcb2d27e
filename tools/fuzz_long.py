#!/usr/bin/env python
"""Randomised parity with long reads (up to 65 535 bases: the kernels that stage records in shared memory give way to the warp-per-read
generation), runs of thousands of equal qualities, overlapping long mates - against the oracle.  TEST INFRASTRUCTURE (uses oracle/).
usage: fuzz_long.py first_seed seconds   (RPQ_FUZZ_LIB=<emulation build> for a run without a GPU)"""
import os, sys, random, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from repaq_b200 import codec as K
from tests import parity
from oracle import oracle as O
lib=os.environ.get('RPQ_FUZZ_LIB'); cd=K.Codec(lib_path=lib) if lib else K.Codec(0)
COMP={65:84,84:65,67:71,71:67}
def rc(s): return bytes(COMP.get(c,78) for c in reversed(s))
t0=time.time(); n=0; seed=int(sys.argv[1])
while time.time()-t0 < float(sys.argv[2]):
    rnd=random.Random(seed)
    maxlen=rnd.choice([400, 1200, 3000, 9000, 30000, 65535])
    nreads=max(4, min(400, 600000//maxlen))
    paired=rnd.random()<0.5
    r1=[];r2=[]
    for i in range(nreads):
        L=rnd.randint(1,maxlen) if rnd.random()<0.7 else maxlen
        s=bytes(rnd.choice(b"ACGT") for _ in range(L))
        if rnd.random()<0.2 and i>50: 
            p=rnd.randrange(L); s=s[:p]+b"N"+s[p+1:]
        q=bytearray(rnd.choice(b"FFFFFFFF,:") for _ in range(L))
        if rnd.random()<0.3:
            a=rnd.randrange(L); z=min(L,a+rnd.randint(1,20000)); q[a:z]=bytes([rnd.choice(b",:#F")])*(z-a)
        nm=b"@M:1:FC:%d:%d:%d:%d 1:N:0:AC"%(rnd.randint(1,8),rnd.randint(1,9999),rnd.randint(0,99999),rnd.randint(0,99999))
        r1.append(nm+b"\n"+s+b"\n+\n"+bytes(q)+b"\n")
        if paired:
            how=rnd.randrange(3)
            if how==0:
                o=rnd.randint(1,L); s2=rc(s[L-o:]+bytes(rnd.choice(b"ACGT") for _ in range(rnd.randint(0,200))))
            else: s2=bytes(rnd.choice(b"ACGT") for _ in range(rnd.randint(1,maxlen)))
            s2=s2[:65535]
            q2=bytes(rnd.choice(b"FFFFFFFF,:") for _ in range(len(s2)))
            r2.append(nm.replace(b" 1:",b" 2:")+b"\n"+s2+b"\n+\n"+q2+b"\n")
    b1=b"".join(r1); b2=b"".join(r2) if paired else None
    try:
        parity.check_against_oracle(cd,b1,b2,k=rnd.choice([100,1000]),roundtrip=False)
    except Exception as e:
        print('FAIL seed',seed,type(e).__name__,str(e)[:800]); sys.exit(1)
    n+=1; seed+=1
print('long fuzz ok',n)
