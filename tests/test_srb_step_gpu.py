"""GPU parity of the CUDA SRBTrack step against the oracle (through the C-ABI)."""
import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu

RTOL64 = 1e-5     # BASELINE.json: single-step state / GRF / QP within 1e-5 relative in fp64


def _golden(variant):
    import os
    return np.load(os.path.join(helpers.GOLDEN, f"rollout_{variant}.npz"))


def _close(a, b, rtol, what):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    scale = np.maximum(1.0, np.abs(b))
    err = np.abs(a - b) / scale
    assert np.nanmax(err) <= rtol, f"{what}: max rel err {np.nanmax(err):.3e} at {np.unravel_index(np.nanargmax(err), err.shape)}"


@pytest.mark.parametrize("variant", ["test", "training"])
@pytest.mark.parametrize("lanes", [8, 32, 1])
def test_golden_rollout_fp64(built_lib, variant, lanes):
    import torch
    from ipcdnnwalk_b200 import SRBVecEnv
    g = _golden(variant)
    n_env, n_steps = g["actions"].shape[:2]
    env = SRBVecEnv(helpers.fresh_clips(), n_env, variant=variant, precision=64, lanes_per_env=lanes, auto_reset=False)
    env.attach(dbg=True)
    obs0 = env.reset(plan=g["plans"]).cpu().numpy()
    _close(obs0, g["obs0"], 2e-6, "reset obs")
    for k in range(n_steps):
        env.dbg.zero_()
        obs, rew, done, rinfo = env.step(torch.as_tensor(g["actions"][:, k], device="cuda"))
        torch.cuda.synchronize()
        st = np.stack([helpers.cuda_state_vector(env.named_state(i)) for i in range(n_env)])
        # integer / boolean outputs: bit exact
        assert (done.cpu().numpy().astype(bool) == g["done"][:, k]).all(), f"done mismatch at step {k}"
        assert (st[:, 23:26] == g["state"][:, k, 23:26]).all(), f"contact flags / frame mismatch at step {k}"
        _close(st[:, :23], g["state"][:, k, :23], RTOL64, f"state step {k}")
        _close(obs.cpu().numpy(), g["obs"][:, k], RTOL64, f"obs step {k}")
        _close(rew.cpu().numpy(), g["rew"][:, k], RTOL64, f"reward step {k}")
        _close(rinfo.cpu().numpy(), g["rinfo"][:, k], RTOL64, f"reward_info step {k}")
        dbg = env.dbg.cpu().numpy().reshape(n_env, 4, 64)
        qdd_ref = g["qdd"][:, k]
        m = ~np.isnan(qdd_ref[..., 0])
        _close(dbg[..., :6][m], qdd_ref[m], RTOL64, f"QP qdd step {k}")
    env.close()
