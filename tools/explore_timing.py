"""GPU exploration: step-kernel time vs precision / lanes-per-env / batch size (not the bench contract)."""
import sys, os, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import helpers
from ipcdnnwalk_b200 import SRBVecEnv

clips = helpers.fresh_clips()
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
def run(N, prec, G, variant="test", steps=60, warm=10, do_flush=True, sigma=0.1):
    env = SRBVecEnv(clips, N, variant=variant, precision=prec, lanes_per_env=G, auto_reset=True, seed=1)
    env.reset()
    act = torch.zeros(N, 22, dtype=torch.float32, device="cuda")
    ts = []
    for k in range(warm + steps):
        env.tracking_action(act, sigma=sigma, seed=1234, step=k)
        if do_flush: flush.fill_(k & 255)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); env.step(act); e1.record()
        torch.cuda.synchronize()
        if k >= warm: ts.append(e0.elapsed_time(e1))
    ep = env.episode_stats()
    ts = np.array(ts)
    print(f"N={N:7d} fp{prec} G={G:2d} {variant:8s} flush={int(do_flush)} step ms: median {np.median(ts):.4f} min {ts.min():.4f} max {ts.max():.4f}"
          f" -> {N/np.median(ts)*1e3:.3e} env-steps/s | episodes {int(ep[2])} mean len {ep[1]/max(ep[2],1):.1f} mean ret {ep[0]/max(ep[2],1):.1f}", flush=True)
    env.close()

which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "small"):
    for prec in (64, 32):
        for G in (1, 2, 4, 8, 16, 32):
            run(4096, prec, G)
    run(4096, 64, 8, do_flush=False); run(4096, 32, 8, do_flush=False)
    run(4096, 64, 8, variant="training"); run(4096, 32, 8, variant="training")
if which in ("all", "big"):
    for prec in (64, 32):
        for G in (1, 2, 4, 8):
            run(262144, prec, G, steps=20, warm=5)
