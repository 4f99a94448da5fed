"""Diagnostic: device AdaptiveSRB step vs the golden oracle roll-out, printing the largest error per quantity and step
(no assertions; used while bringing the kernel up)."""
import os, sys
import numpy as np
import torch
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import helpers_cdm as hc
from ipcdnnwalk_b200 import CDMVecEnv

prec = int(sys.argv[1]) if len(sys.argv) > 1 else 64
lanes = int(sys.argv[2]) if len(sys.argv) > 2 else 4
T = int(sys.argv[3]) if len(sys.argv) > 3 else 30
tab, skel = hc.load_tables()
roll = dict(np.load(hc.ROLLOUT))
n = roll["plans0"].shape[0]
env = CDMVecEnv(tab, skel, n, precision=prec, lanes_per_env=lanes, auto_reset=False)
env.attach(dbg=True)
obs = env.reset(plan=roll["plans0"])
torch.cuda.synchronize()
np.set_printoptions(precision=3, linewidth=220)


def err(a, b):
    a = np.asarray(a, float); b = np.asarray(b, float)
    e = np.abs(a - b) / (1e-3 + np.abs(b))
    i = np.unravel_index(np.argmax(e), e.shape)
    return f"{e.max():.2e}@{tuple(int(x) for x in i)}"


print("reset obs", err(obs.cpu().numpy(), roll["obs0"]))
st = np.stack([env.get_state(i)["reals"] for i in range(n)])
print("reset state", err(st, roll["reals0"]))
for t in range(T):
    obs, rew, done, _ = env.step(torch.as_tensor(roll["acts"][t], device="cuda"))
    torch.cuda.synchronize()
    dbg = env.dbg.cpu().numpy()
    st = np.stack([env.get_state(i)["reals"] for i in range(n)])
    it = np.stack([env.get_state(i)["ints"] for i in range(n)])
    sel = np.ones(61, bool); sel[22] = False
    print(f"t={t} done {done.cpu().numpy()} vs {roll['done'][t].astype(int)} flags {it[:, 0]} vs {roll['ints'][t][:, 0]} nrep {it[:, 1]} vs {roll['ints'][t][:, 1]} iters {dbg[:, 75:77].astype(int).tolist()}")
    print("   qdd", err(dbg[:, 0:12].reshape(n, 2, 6), roll["qdd"][t]), "scal", err(dbg[:, 64:68], roll["scal"][t]), "obs", err(obs.cpu().numpy(), roll["obs"][t]),
          "rew", err(rew.cpu().numpy(), roll["rew"][t]), "state", err(st[:, sel], roll["reals"][t][:, sel]), "lam", err(dbg[:, 24:64].reshape(n, 2, 20), roll["lam"][t]))
    dn = np.nonzero(roll["done"][t])[0]
    if len(dn):
        env.reset(env_idx=dn, plan=roll["plans"][t][dn])
        torch.cuda.synchronize()
        print("   reset obs", err(env.obs.cpu().numpy()[dn], roll["reset_obs"][t][dn]))
    if not (done.cpu().numpy().astype(bool) == roll["done"][t]).all():
        print("   DONE FLAGS DIVERGED; stopping"); break
