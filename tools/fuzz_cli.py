#!/usr/bin/env python
"""Randomised parity of the command-line driver: `repaq_b200_cli -c` / `-d` with small streaming windows against the oracle (random
shape, pairing, CRLF, mutated quality columns, how the input ends).  TEST INFRASTRUCTURE (uses oracle/).
usage: fuzz_cli.py first_seed seconds [emu|gpu]   (default emu: tests/emu/repaq_emu_cli)"""
import os, sys, random, time, subprocess, tempfile, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import importlib.util
spec=importlib.util.spec_from_file_location('fz', os.path.join(ROOT, 'tools', 'fuzz_parity.py')); fz=importlib.util.module_from_spec(spec); spec.loader.exec_module(fz)
from tools import fqgen
from oracle import oracle as O
GPU = len(sys.argv) > 3 and sys.argv[3] == 'gpu'
CLI = os.path.join(ROOT, 'repaq_b200', 'repaq_b200_cli') if GPU else os.path.join(ROOT, 'tests', 'emu', 'repaq_emu_cli')
env = dict(os.environ, LD_LIBRARY_PATH=os.path.dirname(CLI))
t0=time.time(); n=0; seed=int(sys.argv[1])
tmp=tempfile.mkdtemp()
while time.time()-t0 < float(sys.argv[2]):
    rnd=random.Random(seed)
    shape=rnd.choice([fqgen.NOVA, fqgen.BGI]); paired = shape==fqgen.NOVA and rnd.random()<0.5
    flags=rnd.choice([0, fqgen.VARLEN, fqgen.CRLF, fqgen.VARLEN|fqgen.LONG])
    nreads=rnd.choice([3000, 6000, 9000])
    r1,r2=fqgen.generate(nreads, seed=seed, shape=shape, flags=flags, paired=paired)
    b1=fz.mutate_quality(r1, rnd, 400); b2=fz.mutate_quality(r2, rnd, 400) if paired else None
    nl = b"\r\n" if (flags & fqgen.CRLF) else b"\n"
    t=rnd.randrange(6)
    if t==0 and b1.endswith(b"\n"): b1=b1[:-len(nl)]
    elif t==1: b1+=nl*2
    elif t==2: b1+=b"@ragged"+nl+b"ACGT"
    elif t==3 and b2 is not None:
        p=b2.rfind(b"@",0,len(b2)-1); b2=b2[:p]
    open(tmp+'/a.fq','wb').write(b1)
    cmd=[CLI,'-c','-i',tmp+'/a.fq','-o',tmp+'/o.rfq','-k','100']
    if b2 is not None:
        open(tmp+'/b.fq','wb').write(b2); cmd+=['-I',tmp+'/b.fq']
    e=dict(env, RPQ_CLI_FQ_WINDOW=str(rnd.choice([300000, 700000, 1300000])), RPQ_CLI_RFQ_WINDOW=str(rnd.choice([40000, 200000])))
    if os.path.exists(tmp+'/o.rfq'): os.remove(tmp+'/o.rfq')
    rc=subprocess.call(cmd, env=e, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    exp=O.compress(b1,b2,chunk_bases=100000)
    got=open(tmp+'/o.rfq','rb').read() if os.path.exists(tmp+'/o.rfq') else b''
    if rc!=0 or got!=exp:
        print('FAIL compress seed',seed,'rc',rc,len(got),len(exp)); sys.exit(1)
    if len(exp):
        outs=[tmp+'/d1.fq']+([tmp+'/d2.fq'] if b2 is not None else [])
        cmd=[CLI,'-d','-i',tmp+'/o.rfq','-o',outs[0]]+(['-O',outs[1]] if b2 is not None else [])
        rc=subprocess.call(cmd, env=e, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        expd=O.decompress(exp, pe_out=b2 is not None)
        gotd=tuple(open(o,'rb').read() for o in outs) if b2 is not None else open(outs[0],'rb').read()
        if rc!=0 or gotd!=expd:
            print('FAIL decompress seed',seed,'rc',rc); sys.exit(1)
    n+=1; seed+=1
shutil.rmtree(tmp)
print('cli fuzz ok', n, 'inputs')
