#!/usr/bin/env python
"""Differential fuzzing on the CPU (test infrastructure): random small FASTA files (reads of tandem repeats with unit lengths
1-300 and 0-15 % noise between random flanks, lower case, wrapped lines), random flags (-a, -p, -m) and random scheduling
knobs through the product's host sources + engine device code on the simulated device (tests/hostsim) and through the
reference binary built from the reference's sources (oracle/_ref/mTR_ref_det, canonical set order); the bytes must be equal.

    python tools/fuzz_vs_reference.py <first seed> <seconds> [long]   # differing inputs are kept as /tmp/fuzz/BAD_<seed>.fa

Round 2: 19 265 files (6 135 of them `long`), no difference; on 5 of them (-p) the reference itself dies (SIGSEGV in the
middle of the file four times -- state left behind by earlier reads: every read of those files passes alone -- and heap
corruption reported at exit once): its output up to there is a prefix of the product's, which equals the oracle's."""
import sys, os, subprocess, time
import numpy as np
sys.path.insert(0,'/root/repo')
ref='/root/repo/oracle/_ref/mTR_ref_det'; sim='/root/repo/tests/hostsim/_build/mTR_hostsim'
seed0=int(sys.argv[1]); budget=float(sys.argv[2])
big=len(sys.argv)>3 and sys.argv[3]=='long'          # long: reads of several kb (units up to 500, flanks up to 3 kb)
os.makedirs('/tmp/fuzz', exist_ok=True)
t0=time.time(); n=0; bad=0; crashed=0
while time.time()-t0<budget:
    seed=seed0+n; n+=1
    rng=np.random.default_rng(seed)
    nreads=int(rng.integers(1,7))
    recs=[]
    for r in range(nreads):
        parts=[]
        for seg in range(int(rng.integers(1,4))):
            if rng.random()<0.7:
                ul=int(rng.choice([1,2,3,4,5,7,10,13,20,33,50,80,120,200,300]+([400,500] if big else [])))
                unit=rng.integers(0,4,ul)
                copies=int(rng.integers(2,max(3,min(200 if big else 60,(9000 if big else 1500)//ul))))
                rep=np.tile(unit,copies)
                rate=float(rng.choice([0,0.02,0.05,0.1,0.15]))
                out=[]
                for b in rep:
                    u=rng.random()
                    if u<rate/3: continue
                    if u<2*rate/3: out.append(int(rng.integers(0,4)))
                    out.append(int(rng.integers(0,4)) if rng.random()<rate/3 else int(b))
                parts.append(np.array(out,dtype=np.int64))
            else:
                parts.append(rng.integers(0,4,int(rng.integers(0,3000 if big else 400))))
        seq=np.concatenate(parts) if parts else np.zeros(0,dtype=np.int64)
        if len(seq)==0: seq=rng.integers(0,4,5)
        s="".join("ACGT"[int(x)] for x in seq)
        if rng.random()<0.2: s=s.lower()
        lw=int(rng.choice([0,60,80,7]))
        if lw: s="\n".join(s[i:i+lw] for i in range(0,len(s),lw))
        recs.append(">r%d some text\n%s\n"%(r,s))
    path='/tmp/fuzz/f_%d.fa'%seed0
    open(path,'w').write("".join(recs))
    flags=[[],['-a'],['-p'],['-p','-m','0.7'],['-m','0.9'],['-m','0.3']][int(rng.integers(0,6))]
    env=dict(os.environ)
    if rng.random()<0.5: env.update({'MTR_GROUP_READS':str(int(rng.integers(1,4))),'MTR_ENGINE_SLOTS':str(int(rng.integers(1,5))),'MTR_READ_BLOCK_BYTES':str(int(rng.integers(16,3000)))})
    a=subprocess.run([ref]+flags+[path],stdout=subprocess.PIPE,stderr=subprocess.PIPE)
    b=subprocess.run([sim]+flags+[path],stdout=subprocess.PIPE,stderr=subprocess.PIPE,env=env,timeout=600)
    if a.returncode<0 and b.returncode==0 and b.stdout.startswith(a.stdout):
        crashed+=1          # the reference itself died on a signal (seen with -p only); what it had printed is a prefix of the product's output
        print('REFERENCE CRASHED seed',seed,flags,'signal',-a.returncode,flush=True)
    elif a.returncode!=b.returncode or a.stdout!=b.stdout:
        bad+=1
        keep='/tmp/fuzz/BAD_%d.fa'%seed
        os.rename(path,keep)
        print('DIFF seed',seed,flags,{k:v for k,v in env.items() if k.startswith('MTR_')},'rc',a.returncode,b.returncode,'lines',a.stdout.count(b'\n'),b.stdout.count(b'\n'),b.stderr[:100],flush=True)
print('done seed0',seed0,'cases',n,'reference_crashes',crashed,'bad',bad,flush=True)
