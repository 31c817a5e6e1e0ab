#!/usr/bin/env python
"""Differential fuzzing aimed at hazard H4b: every read ends inside its tandem repeat and the 2-6 reads of a
file come in falling or rising length, so that what polish_repeat reads past the end of a read is whatever earlier reads
left there.  Product host sources + engine on the simulated device (tests/hostsim) against oracle/_ref/mTR_ref_det.

    python tools/fuzz_read_end.py <first seed> <seconds>        # round 2: 4757 files, no difference
"""
import sys, os, subprocess, time
import numpy as np
ref='/root/repo/oracle/_ref/mTR_ref_det'; sim='/root/repo/tests/hostsim/_build/mTR_hostsim'
seed0=int(sys.argv[1]); budget=float(sys.argv[2])
os.makedirs('/tmp/fuzz', exist_ok=True)
t0=time.time(); n=0; bad=0; crashed=0
while time.time()-t0<budget:
    seed=seed0+n; n+=1
    rng=np.random.default_rng(seed)
    recs=[]
    lens=sorted([int(rng.integers(60,2500)) for _ in range(int(rng.integers(2,7)))], reverse=bool(rng.integers(0,2)))
    for r,L in enumerate(lens):
        ul=int(rng.choice([1,2,3,5,7,11,20,35,60,110]))
        unit=rng.integers(0,4,ul)
        flank=int(rng.integers(0,max(1,L//2)))
        rep=np.tile(unit,(L-flank)//ul+2)[:L-flank]
        rate=float(rng.choice([0,0.03,0.08]))
        out=[]
        for b in rep:
            u=rng.random()
            if u<rate/3: continue
            if u<2*rate/3: out.append(int(rng.integers(0,4)))
            out.append(int(rng.integers(0,4)) if rng.random()<rate/3 else int(b))
        seq=np.concatenate([rng.integers(0,4,flank),np.array(out,dtype=np.int64)])
        recs.append(">r%d\n%s\n"%(r,"".join("ACGT"[int(x)] for x in seq)))
    path='/tmp/fuzz/h_%d.fa'%seed0
    open(path,'w').write("".join(recs))
    flags=[[],['-a'],['-m','0.8'],['-m','0.4']][int(rng.integers(0,4))]
    a=subprocess.run([ref]+flags+[path],stdout=subprocess.PIPE,stderr=subprocess.PIPE)
    b=subprocess.run([sim]+flags+[path],stdout=subprocess.PIPE,stderr=subprocess.PIPE,timeout=600)
    if a.returncode<0 and b.stdout.startswith(a.stdout): crashed+=1
    elif a.returncode!=b.returncode or a.stdout!=b.stdout:
        bad+=1; os.rename(path,'/tmp/fuzz/HBAD_%d.fa'%seed)
        print('DIFF seed',seed,flags,a.returncode,b.returncode,a.stdout.count(b'\n'),b.stdout.count(b'\n'),flush=True)
print('done',seed0,'cases',n,'refcrash',crashed,'bad',bad,flush=True)
