"""GPU: distribution of Newton systems solved per QP (from the debug side channel)."""
import sys, os
import numpy as np, torch
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,R); sys.path.insert(0,os.path.join(R,"tests"))
import helpers
from ipcdnnwalk_b200 import SRBVecEnv
clips=helpers.fresh_clips()
for prec in (64,32):
    N=4096
    env=SRBVecEnv(clips,N,variant="test",precision=prec,lanes_per_env=8,auto_reset=True,seed=1)
    env.attach(dbg=True); env.reset()
    act=torch.zeros(N,22,device="cuda")
    its=[]
    for k in range(40):
        env.tracking_action(act,sigma=0.1,seed=1234,step=k)
        env.dbg.zero_(); env.step(act); torch.cuda.synchronize()
        d=env.dbg.cpu().numpy().reshape(N,4,64)
        its.append(d[:,:,18])
    its=np.array(its)   # steps, N, 4
    print("fp",prec,"iters per substep mean",its.mean(axis=(0,1)),"max",its.max(),"hist",np.bincount(its.astype(int).ravel())[:12], "total/step mean", its.sum(axis=2).mean())
    env.close()
