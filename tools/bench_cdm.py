"""GPU throughput of the batched AdaptiveSRB step (BASELINE config 4: gym_cdm2 deepmimiccdm-v1, walk2-v1 constants,
16384 environments, actions U(-1,1) (scaled by 0.2 inside the step), seed 3, fused auto-reset).
One JSON line per (precision, lanes) with env-steps/s, the HBM-roofline fraction (1116 B / env-step) and the
end-to-end number through cdm_step_host (pinned host buffers, H2D + kernel + D2H per step).

  python tools/bench_cdm.py [--envs 16384] [--steps 200] [--lanes 1,2,4,8] [--precision 32,64]"""
import argparse, json, os, sys, time
import numpy as np
import torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import helpers_cdm as hc
from ipcdnnwalk_b200 import CDMVecEnv

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=16384)
ap.add_argument("--steps", type=int, default=200)
ap.add_argument("--warmup", type=int, default=16)
ap.add_argument("--lanes", default="4")
ap.add_argument("--precision", default="32,64")
ap.add_argument("--e2e", action="store_true")
ap.add_argument("--epw", default="0", help="environments per warp (comma list; 0 = automatic)")
ap.add_argument("--terrain", action="store_true", help="walkterrain-v1: the gym_cdm2 height field attached (31-float observation); with --policy the reference's walkterrain-v1 policy")
ap.add_argument("--policy", action="store_true", help="drive the environments with the reference's trained walk2-v1 policy (tests/golden/cdm_policy_walk2.npz), evaluated in torch outside the timed events, reset noise scaled to 20 %")
a = ap.parse_args()
peak = 6547.8
try:
    peak = json.load(open(os.path.join(R, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
tab, skel = hc.load_tables()
BYTES = {32: 108 + 40 + 8 + 2 * (61 * 4 + 16) + 2 * 32, 64: 108 + 40 + 8 + 2 * (61 * 8 + 16) + 2 * 64}   # obs + act + rew/done + state r/w + replay entry r/w
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for prec in [int(x) for x in a.precision.split(",")]:
    for lanes, epw in [(int(x), int(y)) for x in a.lanes.split(",") for y in a.epw.split(",")]:
        params = {"noise_q": 0.1, "noise_pos": 0.01, "noise_dq": 0.1, "noise_vx": 0.01, "noise_vy": 0.1, "noise_vz": 0.01} if a.policy else None
        if a.policy and os.environ.get("CDM_NOISE_SCALE"):          # experiment: 0 = all environments identical (perfectly coherent warps)
            params = {k: v * float(os.environ["CDM_NOISE_SCALE"]) for k, v in params.items()}
        params = dict(params or {}, envs_per_warp=epw)
        env = CDMVecEnv(tab, skel, a.envs, precision=prec, lanes_per_env=lanes, auto_reset=True, seed=3, params=params, **({"terrain": hc.load_terrain()} if a.terrain else {}))
        obs = env.reset()
        if a.policy:
            import helpers
            pol = dict(np.load(os.path.join(helpers.GOLDEN, "cdm_policy_walkterrain.npz" if a.terrain else "cdm_policy_walk2.npz")))
            W = [torch.as_tensor(pol[f"W{i}"], dtype=torch.float32, device="cuda") for i in range(int(pol["n_layers"]))]
            Bs = [torch.as_tensor(pol[f"b{i}"], dtype=torch.float32, device="cuda") for i in range(int(pol["n_layers"]))]
            mean = torch.as_tensor(pol["obs_mean"], dtype=torch.float32, device="cuda")
            istd = torch.as_tensor(1.0 / np.sqrt(pol["obs_var"] + 1e-8), dtype=torch.float32, device="cuda")

            def policy_action(o):
                h = torch.clamp((o - mean) * istd, -10.0, 10.0)
                for w, b in zip(W, Bs):
                    h = torch.tanh(h @ w.T + b)
                return h
        g = torch.Generator(device="cuda"); g.manual_seed(3)
        acts = [(torch.rand(a.envs, 10, device="cuda", generator=g) * 2 - 1) for _ in range(32)]
        for i in range(a.warmup if not a.policy else max(a.warmup, 120)):     # policy: let the gait reach its steady state
            obs = env.step(policy_action(obs) if a.policy else acts[i % 32])[0]
        torch.cuda.synchronize()
        env.episode_stats(clear=True)
        ts = []
        for i in range(a.steps):
            flush.fill_(i & 0xff)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            act = policy_action(obs) if a.policy else acts[i % 32]
            e0.record(); obs = env.step(act)[0]; e1.record()
            ts.append((e0, e1))
        torch.cuda.synchronize()
        ms = np.array([x.elapsed_time(y) for x, y in ts])
        st = env.episode_stats()
        out = {"kernel": "cdm_step_kernel", "workload": "gym_cdm2 deepmimiccdmterrain-v1 (walkterrain-v1), height field" if a.terrain else "gym_cdm2 deepmimiccdm-v1 (walk2-v1), flat ground", "actions": ("reference walkterrain-v1 policy (torch, untimed)" if a.terrain else "reference walk2-v1 policy (torch, untimed)") if a.policy else "U(-1,1)", "envs": a.envs, "dtype": f"f{prec}", "lanes_per_env": lanes, "envs_per_warp": epw,
               "steps": a.steps, "ms_per_step_mean": float(ms.mean()), "ms_per_step_median": float(np.median(ms)), "env_steps_per_s": a.envs / float(ms.mean()) * 1e3,
               "algorithmic_bytes_per_env_step": BYTES[prec], "achieved_GBps": BYTES[prec] * a.envs / float(ms.mean()) / 1e6, "peak_GBps": peak,
               "frac": BYTES[prec] * a.envs / float(ms.mean()) / 1e6 / peak, "episodes_finished": int(st[2]), "mean_episode_len": float(st[1] / max(st[2], 1)),
               "mean_episode_return": float(st[0] / max(st[2], 1)), "l2": "flushed between steps (256 MiB write)"}
        if a.e2e:
            p = env.pinned()
            hacts = [x.cpu().numpy() for x in acts[:8]]
            for i in range(4):
                env.step_host(hacts[i % 8])
            t0 = time.perf_counter()
            n = min(a.steps, 100)
            for i in range(n):
                env.step_host(hacts[i % 8])
            dt = time.perf_counter() - t0
            out["e2e"] = {"value": a.envs * n / dt, "unit": "env-steps/s", "h2d_bytes_per_step": a.envs * 40, "d2h_bytes_per_step": a.envs * (108 + 4 + 1)}
        print(json.dumps(out), flush=True)
        env.close()
