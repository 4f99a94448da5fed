"""GPU: per-step kernel times of a short run of the SRBTrack step (series, not the mean) -- shows warm-up and outliers."""
import sys, os
import numpy as np, torch
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,R); sys.path.insert(0,os.path.join(R,"tests"))
import helpers
from ipcdnnwalk_b200 import SRBVecEnv
clips=helpers.fresh_clips()
for prec in (32,64):
    N=4096
    env=SRBVecEnv(clips,N,variant="test",precision=prec,lanes_per_env=8,auto_reset=True,seed=1)
    env.reset()
    act=torch.zeros(N,22,device="cuda")
    ts=[]
    for k in range(60):
        env.tracking_action(act,sigma=0.1,seed=1234,step=k)
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); env.step(act); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1)*1e3)
    print("fp",prec," ".join(f"{t:.0f}" for t in ts))
    env.close()
