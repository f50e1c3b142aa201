#!/usr/bin/env python
"""fuzz_multirank.py [seed] [batches] [big] -- random multi-rank cases on the CPU emulation: random grids, processor grids on 2-4
ranks, memory orders, forward / backward / C2C, fused derivative, in-place, random chunking of the overlapped pairs; every
rank checks its block against the oracle (tests/mp_worker.py).  Development tool; tests/test_multirank.py runs two batches."""
import json,sys,subprocess,os,signal,random,itertools
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from test_multirank import *
rnd=random.Random(int(sys.argv[1]) if len(sys.argv)>1 else 1)
perms=list(itertools.permutations((0,1,2)))
BIG=[128,192,256,384] if len(sys.argv)>3 and sys.argv[3]=='big' else []  # kernel sizes: TMA-fed, tensor-load and mixed-radix kernels as exchange stages
def rand_case(world):
    grids={2:[[1,1,2],[1,2,1]],3:[[1,1,3],[1,3,1]],4:[[1,1,4],[1,2,2],[1,4,1]]}[world]
    pd=rnd.choice(grids)
    n=(rnd.choice([8,12,16,20,32,64]+BIG), rnd.randint(4,24), rnd.randint(4,24))
    kind=rnd.choice(["fwd","bwd","c2c"])
    kw={}
    if rnd.random()<0.5:
        kw["mo1"]=list(rnd.choice(perms)); kw["mo2"]=list(rnd.choice(perms))
    if kind=="fwd":
        c=fwd(n,pd,**kw)
        if rnd.random()<0.3: c["deriv"]=rnd.randint(0,2)
    elif kind=="bwd":
        if "mo1" in kw: kw["mo1"],kw["mo2"]=kw["mo2"],kw["mo1"]
        c=bwd(n,pd,**kw)
    else:
        c=c2c(n,pd,**kw)
    if rnd.random()<0.3: c["inplace"]=True
    c["reps"]=1; c["expect_pairs"]=False
    return c
bad=0
for it in range(int(sys.argv[2]) if len(sys.argv)>2 else 6):
    world=rnd.choice([2,3,4])
    cases=[rand_case(world) for _ in range(6)]
    env=dict(os.environ)
    if rnd.random()<0.7:
        env.update({"P3DFFT_B200_OVERLAP_ALIGN":str(rnd.choice([1,2,4])),"P3DFFT_B200_OVERLAP_CHUNKS":str(rnd.choice([2,3,5]))})
    cmd=[sys.executable,os.path.join(os.path.dirname(os.path.abspath(__file__)), 'mpirun.py'),'-np',str(world),sys.executable,os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'mp_worker.py'),'emu',json.dumps(cases)]
    p=subprocess.Popen(cmd,stdout=subprocess.PIPE,stderr=subprocess.STDOUT,text=True,start_new_session=True,env=env)
    try:
        out,_=p.communicate(timeout=300)
        ok = p.returncode==0 and out.count(" OK worst")==world
    except subprocess.TimeoutExpired:
        os.killpg(p.pid, signal.SIGKILL); out,_=p.communicate(); ok=False; out="HANG "+out
    print(it, world, "ok" if ok else "FAIL", flush=True)
    if not ok:
        bad+=1
        print(json.dumps(cases)); print({k:v for k,v in env.items() if k.startswith("P3DFFT_B200")}); print(out[-1500:])
print("failures", bad)
sys.exit(1 if bad else 0)
