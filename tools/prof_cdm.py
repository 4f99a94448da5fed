"""Short driver for ncu captures of the AdaptiveSRB step kernel. usage: prof_cdm.py PREC LANES N STEPS [policy]
`policy`: actions from the reference's trained walk2-v1 policy (tests/golden/cdm_policy_walk2.npz), reset noise scaled to 20 % --
the workload of bench.py's cdm record; capture launches after ~120 steps, when the gait is steady."""
import sys, os
import numpy as np
import torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import helpers, helpers_cdm as hc
from ipcdnnwalk_b200 import CDMVecEnv
prec = int(sys.argv[1]); G = int(sys.argv[2]); N = int(sys.argv[3]); steps = int(sys.argv[4])
policy = len(sys.argv) > 5 and sys.argv[5] == "policy"
tab, skel = hc.load_tables()
params = {"noise_q": 0.1, "noise_pos": 0.01, "noise_dq": 0.1, "noise_vx": 0.01, "noise_vy": 0.1, "noise_vz": 0.01} if policy else None
env = CDMVecEnv(tab, skel, N, precision=prec, lanes_per_env=G, auto_reset=True, seed=3, params=params)
obs = env.reset()
g = torch.Generator(device="cuda"); g.manual_seed(3)
if policy:
    pol = dict(np.load(os.path.join(helpers.GOLDEN, "cdm_policy_walk2.npz")))
    W = [torch.as_tensor(pol[f"W{i}"], dtype=torch.float32, device="cuda") for i in range(int(pol["n_layers"]))]
    B = [torch.as_tensor(pol[f"b{i}"], dtype=torch.float32, device="cuda") for i in range(int(pol["n_layers"]))]
    mean = torch.as_tensor(pol["obs_mean"], dtype=torch.float32, device="cuda")
    istd = torch.as_tensor(1.0 / np.sqrt(pol["obs_var"] + 1e-8), dtype=torch.float32, device="cuda")
for k in range(steps):
    if policy:
        h = torch.clamp((obs - mean) * istd, -10.0, 10.0)
        for w, b in zip(W, B):
            h = torch.tanh(h @ w.T + b)
        obs = env.step(h)[0]
    else:
        obs = env.step(torch.rand(N, 10, device="cuda", generator=g) * 2 - 1)[0]
torch.cuda.synchronize()
print("ok")
