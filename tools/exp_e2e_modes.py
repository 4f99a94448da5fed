"""Wall-clock time of srb_step_host per host mode (0 staged copies, 1 zero copy, 2 split) at the headline batch.
   python tools/exp_e2e_modes.py [envs] [steps]"""
import json, os, sys, time
import numpy as np
import torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
from ipcdnnwalk_b200 import SRBVecEnv
from ipcdnnwalk_b200.srb_env import build_clip_tables
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
clips = build_clip_tables(["ref_contact/stand1.motion.npz", "ref_contact/aiming_show.motion.npz"], data_root=os.path.join(R, "tests", "golden"))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for mode in (3, 1, 2, 0, 3, 1):
    env = SRBVecEnv(clips, n, variant="test", precision=32, auto_reset=True, seed=1)
    env.set_host_mode(mode)
    env.reset()
    act = torch.zeros(n, 22, device="cuda")
    pin = env.pinned()
    for k in range(5):
        env.step_host()
    if mode == 2:
        env.host_timeline(True)
    tls = []
    ts = []
    for k in range(steps):
        env.tracking_action(act, sigma=0.1, seed=7, step=k)
        pin["act"][...] = act.cpu().numpy()
        flush.fill_(k & 0xFF)
        torch.cuda.synchronize()
        t0 = time.perf_counter(); env.step_host(); ts.append(time.perf_counter() - t0)
        if mode == 2:
            tls.append(list(env.host_timeline(True).values()))
    ts = np.array(ts) * 1e6
    if tls:
        print("timeline us (median): window_gather %.1f obs_copy %.1f step_kernel %.1f patch %.1f" % tuple(1e3 * np.median(np.array(tls), axis=0)[1:]))
    print(json.dumps({"mode": mode, "envs": n, "us_mean": float(ts.mean()), "us_median": float(np.median(ts)), "env_steps_per_s": n / ts.mean() * 1e6}), flush=True)
    env.close()
