"""BASELINE.json: "rollout reward statistics must be statistically indistinguishable over 1000 episodes".

Common random numbers: 16384 environments start from the same explicit reset draws and receive the same action noise
(Philox keyed by environment and step) in the CUDA fp64 build, the CUDA fp32 build (--use_fast_math) and -- for a
subset, replayed action by action -- the oracle.  Only the FIRST episode of every environment is used, so the samples
are paired one to one.

  * oracle vs CUDA fp64: the same episodes, so lengths must be identical and returns agree to round-off: the fp64
    statistics ARE the oracle's statistics;
  * CUDA fp32 vs CUDA fp64 (16384 paired episodes): Welch z of return, length and return per step below 3, and the
    paired bias of the return inside an equivalence margin of 1 % of the mean return (reported with its 95 % CI)."""
import json
import multiprocessing as mp
import os

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu

N_ENVS = 16384
N_ORACLE = 256
MAX_STEPS = 512          # an episode ends at current_frame >= 511 at the latest (_get_done :1001)


def _first_episodes(precision, plans, n_record):
    import torch
    from ipcdnnwalk_b200 import SRBVecEnv
    n = plans.shape[0]
    env = SRBVecEnv(helpers.fresh_clips(), n, variant="test", precision=precision, auto_reset=True, seed=5)
    env.reset(plan=plans)
    act = torch.zeros(n, 22, device="cuda")
    ret = torch.zeros(n, dtype=torch.float64, device="cuda"); length = torch.zeros(n, dtype=torch.int32, device="cuda")
    alive = torch.ones(n, dtype=torch.bool, device="cuda")
    acts = []
    for k in range(MAX_STEPS):
        env.tracking_action(act, sigma=0.1, seed=77, step=k)
        if n_record:
            acts.append(act[:n_record].cpu().numpy().copy())
        _, rew, done, _ = env.step(act)
        ret += torch.where(alive, rew.double(), torch.zeros_like(ret))
        length += alive.int()
        alive &= ~done.bool()
        if k % 16 == 15 and not bool(alive.any()):
            break
    assert not bool(alive.any())
    env.close()
    return ret.cpu().numpy(), length.cpu().numpy(), (np.stack(acts) if n_record else None)


def _one_thread():
    try:                                   # forked workers must not each spin up a BLAS thread pool
        import threadpoolctl
        threadpoolctl.threadpool_limits(1)
    except Exception:                      # noqa: BLE001
        pass


def _oracle_episode(args):
    plan, acts = args
    from oracle.srb_env import OracleSRBEnv
    o = OracleSRBEnv(helpers.fresh_clips(), "test")
    helpers.oracle_reset(o, plan)
    ret = 0.0
    for k in range(acts.shape[0]):
        _, r, done, _, _ = o.step(acts[k])
        ret += r
        if done:
            return ret, k + 1
    return ret, acts.shape[0]


def _welch(a, b):
    return (a.mean() - b.mean()) / np.sqrt(a.var(ddof=1) / a.size + b.var(ddof=1) / b.size)


def test_rollout_statistics_common_random_numbers(built_lib):
    clips = helpers.fresh_clips()
    rng = np.random.default_rng(2025)
    plans = np.stack([helpers.random_plan(rng, clips) for _ in range(N_ENVS)])
    r64, l64, acts = _first_episodes(64, plans, N_ORACLE)
    r32, l32, _ = _first_episodes(32, plans, 0)

    # ---- oracle vs CUDA fp64, paired episode by episode
    jobs = [(plans[i], acts[:l64[i], i].copy()) for i in range(N_ORACLE)]
    with mp.get_context("fork").Pool(min(os.cpu_count() or 1, 32), initializer=_one_thread) as pool:
        res = pool.map(_oracle_episode, jobs, chunksize=4)
    ro = np.array([r for r, _ in res]); lo = np.array([l for _, l in res])
    same_len = lo == l64[:N_ORACLE]
    assert same_len.mean() >= 0.99, f"oracle / fp64 episode lengths differ in {(~same_len).sum()} of {N_ORACLE} episodes"
    rel = np.abs(ro - r64[:N_ORACLE])[same_len] / np.maximum(1.0, np.abs(ro[same_len]))
    assert rel.max() <= 1e-4, f"oracle / fp64 episode returns differ by {rel.max():.2e}"

    # ---- CUDA fp32 vs CUDA fp64
    z_ret = _welch(r32, r64); z_len = _welch(l32.astype(float), l64.astype(float)); z_rps = _welch(r32 / l32, r64 / l64)
    d = r32 - r64
    bias, half = d.mean(), 1.96 * d.std(ddof=1) / np.sqrt(d.size)
    out = dict(episodes=int(N_ENVS), oracle_episodes=int(N_ORACLE), oracle_len_identical=float(same_len.mean()), oracle_ret_max_rel=float(rel.max()),
               ret_mean_fp64=float(r64.mean()), ret_std_fp64=float(r64.std()), ret_mean_fp32=float(r32.mean()), len_mean_fp64=float(l64.mean()),
               len_mean_fp32=float(l32.mean()), welch_z_return=float(z_ret), welch_z_length=float(z_len), welch_z_return_per_step=float(z_rps),
               paired_bias_return=float(bias), paired_bias_ci95=[float(bias - half), float(bias + half)], same_length_fraction_fp32=float((l32 == l64).mean()))
    print("stat_compare:", json.dumps(out))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if os.path.isdir(os.path.join(root, "gpurun_out")):
        with open(os.path.join(root, "gpurun_out", "stat_compare_crn.json"), "w") as f:
            json.dump(out, f, indent=1)
    assert abs(z_ret) < 3 and abs(z_len) < 3 and abs(z_rps) < 3, out
    assert abs(bias) + half <= 0.01 * abs(r64.mean()), out
