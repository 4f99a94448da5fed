"""Roll-out statistics of the CUDA path against the oracle over >= 1000 episodes each (BASELINE north_star: "rollout reward
statistics must be statistically indistinguishable over 1000 episodes").  Both sides draw their own resets (reference
distributions) and tracking actions + N(0, 0.1^2); episodes end on `done`.  Prints one JSON line with means, standard
deviations and Welch z-scores of episode return and length.

  python tests/tools/stat_compare.py [--episodes 1024] [--precision 32]"""
import argparse, json, math, os, sys, time
for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ.setdefault(_v, "1")
import numpy as np
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
SIGMA = 0.1


def oracle_worker(a):
    wid, n_ep = a
    import helpers
    from oracle.srb_env import OracleSRBEnv
    rng = np.random.default_rng(4000 + wid)
    clips = helpers.fresh_clips()
    e = OracleSRBEnv(helpers.fresh_clips(), "test")
    out = []
    for _ in range(n_ep):
        helpers.oracle_reset(e, helpers.random_plan(rng, clips))
        ret, ln = 0.0, 0
        while True:
            _, r, done, _, _ = e.step(helpers.tracking_action(e, rng, SIGMA))
            ret += r; ln += 1
            if done:
                break
        out.append((ret, ln))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--episodes", type=int, default=1024)
    ap.add_argument("--precision", type=int, default=32)
    a = ap.parse_args()
    import multiprocessing as mp
    import torch
    import helpers
    from ipcdnnwalk_b200 import SRBVecEnv
    cores = max(1, len(os.sched_getaffinity(0)))
    per = (a.episodes + cores - 1) // cores
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(oracle_worker, [(w, per) for w in range(cores)])
    t_oracle = time.perf_counter() - t0
    o = np.array([x for r in res for x in r])
    # device: episodes of 2048 environments with fused auto-reset, per-episode return / length accumulated on the device
    n = 2048
    env = SRBVecEnv(helpers.fresh_clips(), n, variant="test", precision=a.precision, auto_reset=True, seed=31)
    env.reset()
    act = torch.zeros(n, 22, device="cuda")
    ret = torch.zeros(n, dtype=torch.float64, device="cuda"); ln = torch.zeros(n, dtype=torch.float64, device="cuda")
    # unbiased sample: the FIRST episode of every environment, run until each has finished one (<= 512 steps)
    first = torch.zeros(n, dtype=torch.bool, device="cuda")
    out = torch.zeros(n, 2, dtype=torch.float64, device="cuda")
    k = 0
    while not bool(first.all()) and k < 600:
        env.tracking_action(act, sigma=SIGMA, seed=123, step=k)
        _, rew, done, _ = env.step(act)
        ret += torch.where(first, torch.zeros_like(ret), rew.double()); ln += (~first).double()
        newly = done.bool() & ~first
        out[newly, 0] = ret[newly]; out[newly, 1] = ln[newly]
        first |= newly
        k += 1
    g = out[first].cpu().numpy()

    env.close()

    def z(x, y):
        return float((x.mean() - y.mean()) / math.sqrt(x.var(ddof=1) / len(x) + y.var(ddof=1) / len(y)))
    line = {"check": "episode statistics, CUDA path vs oracle", "dtype": f"f{a.precision}", "sigma": SIGMA,
            "oracle": {"episodes": int(len(o)), "return_mean": float(o[:, 0].mean()), "return_std": float(o[:, 0].std(ddof=1)),
                       "length_mean": float(o[:, 1].mean()), "length_std": float(o[:, 1].std(ddof=1)), "seconds": t_oracle, "cores": cores},
            "cuda": {"episodes": int(len(g)), "return_mean": float(g[:, 0].mean()), "return_std": float(g[:, 0].std(ddof=1)),
                     "length_mean": float(g[:, 1].mean()), "length_std": float(g[:, 1].std(ddof=1)), "env_steps": int(k * n)},
            "welch_z": {"return": z(g[:, 0], o[:, 0]), "length": z(g[:, 1], o[:, 1]),
                        "return_per_step": z(g[:, 0] / g[:, 1], o[:, 0] / o[:, 1])},
            "note": "device sample = the first episode of each of 2048 environments (run until every one has finished)"}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
