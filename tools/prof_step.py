"""Short driver for ncu captures: N envs, a few steps of the step kernel. usage: prof_step.py PREC LANES N STEPS"""
import sys, os
import torch
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,R); sys.path.insert(0,os.path.join(R,"tests"))
import helpers
from ipcdnnwalk_b200 import SRBVecEnv
prec=int(sys.argv[1]); G=int(sys.argv[2]); N=int(sys.argv[3]); steps=int(sys.argv[4])
variant=sys.argv[5] if len(sys.argv)>5 else "test"
env=SRBVecEnv(helpers.fresh_clips(),N,variant=variant,precision=prec,lanes_per_env=G,auto_reset=True,seed=1)
env.reset()
act=torch.zeros(N,22,device="cuda")
for k in range(steps):
    env.tracking_action(act,sigma=0.1,seed=1234,step=k)
    env.step(act)
torch.cuda.synchronize()
print("ok")
