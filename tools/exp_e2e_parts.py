"""Experiment: wall-clock anatomy of one end-to-end step (4096 envs): launch + kernel + sync with every combination of
device / mapped-host placement of the action and result arrays, and the staged-copy pieces one by one."""
import ctypes as C, os, sys, time
import numpy as np, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import helpers
from ipcdnnwalk_b200 import SRBVecEnv
from ipcdnnwalk_b200._lib import check
N = 4096
env = SRBVecEnv(helpers.fresh_clips(), N, variant="test", precision=32, auto_reset=True, seed=1)
env.reset()
act = torch.zeros(N, 22, device="cuda")
pin = env.pinned()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def P(a): return a.ctypes.data_as(C.c_void_p) if isinstance(a, np.ndarray) else C.c_void_p(a.data_ptr())
stream = torch.cuda.Stream()
sp = C.c_void_p(stream.cuda_stream)
h_act = torch.from_numpy(pin["act"]); h_obs = torch.from_numpy(pin["obs"]); h_rew = torch.from_numpy(pin["rew"]); h_done = torch.from_numpy(pin["done"]); h_ri = torch.from_numpy(pin["rinfo"])
def step(a, o, r, d, ri):
    check(env.lib.srb_step(env._h, P(a), P(o), P(r), P(d), P(ri), sp))
modes = {
    "all device (launch+kernel+sync)": lambda: step(act, env.obs, env.rew, env.done, env.rinfo),
    "act host": lambda: step(pin["act"], env.obs, env.rew, env.done, env.rinfo),
    "obs host": lambda: step(act, pin["obs"], env.rew, env.done, env.rinfo),
    "rew/done/info host": lambda: step(act, env.obs, pin["rew"], pin["done"], pin["rinfo"]),
    "all host": lambda: step(pin["act"], pin["obs"], pin["rew"], pin["done"], pin["rinfo"]),
}
def copy_in():
    with torch.cuda.stream(stream): act.copy_(h_act, non_blocking=True)
def copy_obs():
    with torch.cuda.stream(stream): h_obs.copy_(env.obs, non_blocking=True)
def copy_small():
    with torch.cuda.stream(stream):
        h_rew.copy_(env.rew, non_blocking=True); h_done.copy_(env.done, non_blocking=True); h_ri.copy_(env.rinfo, non_blocking=True)
modes["H2D act + all device"] = lambda: (copy_in(), step(act, env.obs, env.rew, env.done, env.rinfo))
modes["all device + D2H obs"] = lambda: (step(act, env.obs, env.rew, env.done, env.rinfo), copy_obs())
modes["all device + D2H obs + 3 small"] = lambda: (step(act, env.obs, env.rew, env.done, env.rinfo), copy_obs(), copy_small())
modes["act host + D2H obs + small host"] = lambda: (step(pin["act"], env.obs, pin["rew"], pin["done"], pin["rinfo"]), copy_obs())
for name, fn in modes.items():
    tt = 0.0; K = 150
    for k in range(K + 3):
        env.tracking_action(act, sigma=0.1, seed=7, step=k)
        pin["act"][...] = act.cpu().numpy()
        flush.fill_(k & 0xff)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        stream.synchronize()
        if k >= 3: tt += time.perf_counter() - t0
    print(f"{name:40s} {tt / K * 1e6:7.1f} us/step", flush=True)
