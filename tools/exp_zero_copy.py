"""Experiment: end-to-end step through staged copies (srb_step_host) vs the kernel reading actions from / writing results to
mapped pinned host memory directly (zero copy)."""
import ctypes as C, os, sys, time
import numpy as np, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import helpers
from ipcdnnwalk_b200 import SRBVecEnv
from ipcdnnwalk_b200._lib import check
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
env = SRBVecEnv(helpers.fresh_clips(), N, variant="test", precision=32, auto_reset=True, seed=1)
env.reset()
act = torch.zeros(N, 22, device="cuda")
pin = env.pinned()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def P(a): return a.ctypes.data_as(C.c_void_p)
stream = torch.cuda.Stream()
def run(mode, K=200):
    tt = 0.0
    for k in range(K + 3):
        env.tracking_action(act, sigma=0.1, seed=7, step=k)
        pin["act"][...] = act.cpu().numpy()
        flush.fill_(k & 0xff)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if mode == "staged":
            env.set_host_mode(0); env.step_host()
        elif mode == "host1":
            env.set_host_mode(1); env.step_host()
        elif mode == "zero":
            check(env.lib.srb_step(env._h, P(pin["act"]), P(pin["obs"]), P(pin["rew"]), P(pin["done"]), P(pin["rinfo"]), C.c_void_p(stream.cuda_stream)))
            stream.synchronize()
        elif mode == "zero_in":    # actions read from host, results written to device + one copy each
            check(env.lib.srb_step(env._h, P(pin["act"]), C.c_void_p(env.obs.data_ptr()), C.c_void_p(env.rew.data_ptr()), C.c_void_p(env.done.data_ptr()), C.c_void_p(env.rinfo.data_ptr()), C.c_void_p(stream.cuda_stream)))
            stream.synchronize()
        if k >= 3:
            tt += time.perf_counter() - t0
    return tt / K * 1e6
for mode in ("staged", "host1", "staged", "host1"):
    us = run(mode)
    print(mode, f"{us:.1f} us/step  {N / us * 1e6:.3e} env-steps/s", flush=True)
