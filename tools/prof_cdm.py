"""Short driver for ncu captures of the AdaptiveSRB step kernel. usage: prof_cdm.py PREC LANES N STEPS"""
import sys, os
import torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import helpers_cdm as hc
from ipcdnnwalk_b200 import CDMVecEnv
prec = int(sys.argv[1]); G = int(sys.argv[2]); N = int(sys.argv[3]); steps = int(sys.argv[4])
tab, skel = hc.load_tables()
env = CDMVecEnv(tab, skel, N, precision=prec, lanes_per_env=G, auto_reset=True, seed=3)
env.reset()
g = torch.Generator(device="cuda"); g.manual_seed(3)
for k in range(steps):
    env.step(torch.rand(N, 10, device="cuda", generator=g) * 2 - 1)
torch.cuda.synchronize()
print("ok")
