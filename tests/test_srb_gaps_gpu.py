"""GPU parity of the SRBTrack step against the oracle for the paths and instantiations that the golden roll-outs of
tests/test_srb_step_gpu.py do not reach: contact-probability hysteresis, the external-push schedule with a 6-vector
force, every lane count that is built, the fp32 instantiations behind the benchmark numbers (2 / 4 lanes, terrain with 4
lanes, the coalesced-I/O kernel of the host-buffer path) and samples at the benchmark batch sizes.  Everything goes
through the C-ABI (ipcdnnwalk_b200.SRBVecEnv -> libsrbtrack_b200.so)."""
import os

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu
NPZ = os.path.join(helpers.GOLDEN, "terrain_playground.npz")


def _close(a, b, rtol, what):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    err = np.abs(a - b) / np.maximum(1.0, np.abs(b))
    assert np.nanmax(err) <= rtol, f"{what}: max rel err {np.nanmax(err):.3e} at {np.unravel_index(np.nanargmax(err), err.shape)}"


def _oracles(n, variant, seed, clips=None, noise=True):
    from oracle.srb_env import OracleSRBEnv
    rng = np.random.default_rng(seed)
    clips = clips or helpers.fresh_clips()
    plans = np.stack([helpers.random_plan(rng, clips, noise=noise) for _ in range(n)])
    envs = []
    for i in range(n):
        o = OracleSRBEnv(helpers.fresh_clips(), variant)
        helpers.oracle_reset(o, plans[i])
        envs.append(o)
    return envs, plans, rng


def _compare_step(env, oracles, acts, obs, rew, done, rtol, what, rew_rtol=None):
    dones = []
    for i, o in enumerate(oracles):
        oo, rr, dd, _, _ = o.step(acts[i])
        dones.append(dd)
        st = helpers.cuda_state_vector(env.named_state(i)); ref = helpers.oracle_state_vector(o)
        assert (st[23:26] == ref[23:26]).all(), f"{what}: contact flags / frame must be bit exact (env {i})"
        assert bool(done[i]) == dd, f"{what}: done (env {i})"
        _close(st[:23], ref[:23], rtol, f"{what} state env {i}")
        _close(obs[i], oo, rtol, f"{what} obs env {i}")
        _close(float(rew[i]), rr, rew_rtol or rtol, f"{what} reward env {i}")
    return dones


# ---------------------------------------------------------------------------------------------------------------------
def test_contact_probability_hysteresis_matches_oracle(built_lib):
    """SRBEnv.step :239-275 -- with Policy_prior_posterior_MoE_window._last_contact_prob set, the contact flags follow the
    0.7 / 0.3 hysteresis on the previous flags instead of the arg-max of the one-hot (this is the branch the reference demo
    runs through).  NaN rows = `contact_prob[0] == None` (arg-max path), mixed in the same batch."""
    import torch
    from ipcdnnwalk_b200 import SRBVecEnv
    n = 8
    oracles, plans, rng = _oracles(n, "test", seed=101)
    env = SRBVecEnv(helpers.fresh_clips(), n, variant="test", precision=64, auto_reset=False)
    env.attach(contact_prob=True)
    env.reset(plan=plans)
    switched = 0
    held = 0
    for k in range(30):
        acts = np.stack([helpers.tracking_action(o, rng, 0.1) for o in oracles])
        # probabilities as the policy hands them over: float32 values turned into Python floats (.item())
        prob = rng.uniform(0, 1, (n, 2)).astype(np.float32)
        prob[rng.uniform(size=(n, 2)) < 0.3] = np.float32(0.5)          # inside the dead band: flags are held
        if k % 7 == 3:
            prob[0] = [0.3, 0.7]                                          # exactly on the thresholds (strict comparisons)
        none_rows = (np.arange(n) % 4 == 3)                              # these environments take the arg-max path
        for i, o in enumerate(oracles):
            o.contact_prob = [None, None] if none_rows[i] else [float(prob[i, 0]), float(prob[i, 1])]
        p_dev = prob.copy(); p_dev[none_rows] = np.nan
        env.contact_prob.copy_(torch.as_tensor(p_dev))
        before = [(o.left_foot_in_contact, o.right_foot_in_contact) for o in oracles]
        obs, rew, done, _ = env.step(torch.as_tensor(acts, device="cuda"))
        torch.cuda.synchronize()
        _compare_step(env, oracles, acts, obs.cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy(), 1e-5, f"hysteresis step {k}")
        for i, o in enumerate(oracles):
            if none_rows[i]:
                continue
            now = (o.left_foot_in_contact, o.right_foot_in_contact)
            am = int(np.argmax(acts[i, 18:22]))
            switched += now != before[i]
            held += now != (am in (1, 3), am in (2, 3))                  # differs from what the arg-max path would give
    assert switched >= 10 and held >= 10, "the hysteresis branch must actually be exercised"
    env.close()


@pytest.mark.parametrize("variant", ["test", "training"])
def test_impulse_schedule_with_six_vector_force(built_lib, variant):
    """update_impulse :191-206: `f[0:len(self.force)] = self.force` -> data.xfrc_applied[1]; the push of
    walk_SRBTrack2025_terrain_singlethreaded.py:98-120 is a 6-vector (force, torque = impulse_lpos x force).  The
    schedule (interval / duration / frame_counter across resets) and both halves of the wrench are compared."""
    import torch
    from ipcdnnwalk_b200 import SRBVecEnv
    n = 4
    oracles, plans, rng = _oracles(n, variant, seed=202)
    force = np.array([150.0, -90.0, 40.0, 0.0, 0.0, 0.0])
    lpos = np.array([0.13, -0.18, 0.18]) * 0.5
    force[3:] = np.cross(lpos, force[:3])
    interval, duration = 9, 5
    env = SRBVecEnv(helpers.fresh_clips(), n, variant=variant, precision=64, auto_reset=False)
    env.set_impulse(interval, duration, force)
    env3 = SRBVecEnv(helpers.fresh_clips(), n, variant=variant, precision=64, auto_reset=False)
    env3.set_impulse(interval, duration, force[:3])
    env.reset(plan=plans); env3.reset(plan=plans)
    for o in oracles:
        o.impulse_interval, o.impulse_duration, o.force = interval, duration, list(force)
    diff = 0.0
    pushed = 0
    for k in range(14):
        acts = np.stack([helpers.tracking_action(o, rng, 0.05) for o in oracles])
        obs, rew, done, _ = env.step(torch.as_tensor(acts, device="cuda"))
        env3.step(torch.as_tensor(acts, device="cuda"))
        torch.cuda.synchronize()
        if variant == "training":          # update_impulse is not called by the training env's step (:225-419)
            for o in oracles:
                assert o.frame_counter == 0
        _compare_step(env, oracles, acts, obs.cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy(), 1e-5, f"impulse step {k}")
        for i, o in enumerate(oracles):
            s = env.named_state(i)
            assert s["frame_counter"] == o.frame_counter and s["impulse"] == o.impulse, "impulse schedule counters"
            pushed += o.impulse > 0
            diff = max(diff, np.abs(s["qvel"][3:] - env3.named_state(i)["qvel"][3:]).max())
        if k == 6:                         # the schedule persists across a reset (frame_counter / impulse are not cleared)
            for i, o in enumerate(oracles):
                helpers.oracle_reset(o, plans[i])
            env.reset(plan=plans); env3.reset(plan=plans)
    if variant == "test":
        assert pushed > 0 and diff > 1e-3, "the torque half of the 6-vector must change the angular velocity"
    env.close(); env3.close()


@pytest.mark.parametrize("variant", ["test", "training"])
@pytest.mark.parametrize("lanes", [2, 4, 16])
def test_golden_rollout_fp64_other_lane_counts(built_lib, variant, lanes):
    """The golden roll-outs of test_srb_step_gpu.py for the remaining lane counts (8, 32, 1 are covered there)."""
    import torch
    from ipcdnnwalk_b200 import SRBVecEnv
    g = np.load(os.path.join(helpers.GOLDEN, f"rollout_{variant}.npz"))
    n_env, n_steps = g["actions"].shape[:2]
    env = SRBVecEnv(helpers.fresh_clips(), n_env, variant=variant, precision=64, lanes_per_env=lanes, auto_reset=False)
    env.attach(dbg=True)
    _close(env.reset(plan=g["plans"]).cpu().numpy(), g["obs0"], 2e-6, "reset obs")
    for k in range(n_steps):
        env.dbg.zero_()
        obs, rew, done, rinfo = env.step(torch.as_tensor(g["actions"][:, k], device="cuda"))
        torch.cuda.synchronize()
        st = np.stack([helpers.cuda_state_vector(env.named_state(i)) for i in range(n_env)])
        assert (done.cpu().numpy().astype(bool) == g["done"][:, k]).all(), f"done mismatch at step {k}"
        assert (st[:, 23:26] == g["state"][:, k, 23:26]).all(), f"contact flags / frame mismatch at step {k}"
        _close(st[:, :23], g["state"][:, k, :23], 1e-5, f"state step {k}")
        _close(obs.cpu().numpy(), g["obs"][:, k], 1e-5, f"obs step {k}")
        _close(rew.cpu().numpy(), g["rew"][:, k], 1e-5, f"reward step {k}")
        _close(rinfo.cpu().numpy(), g["rinfo"][:, k], 1e-5, f"reward_info step {k}")
        dbg = env.dbg.cpu().numpy().reshape(n_env, 4, 64)
        m = ~np.isnan(g["qdd"][:, k][..., 0])
        _close(dbg[..., :6][m], g["qdd"][:, k][m], 1e-5, f"QP qdd step {k}")
    env.close()


@pytest.mark.parametrize("lanes,host", [(2, False), (4, False), (16, False), (8, True), (4, True)])
def test_single_step_fp32_other_instantiations(built_lib, lanes, host):
    """fp32 at 1e-3 for the instantiations behind the benchmark numbers that test_single_step_fp32_within_1e3 (8 lanes,
    device buffers) does not run: 2 lanes (>16384 envs per GPU), 4 lanes (<=16384), and the coalesced-I/O kernel that
    srb_step_host launches on mapped pinned buffers (the e2e number).  State reloaded from the oracle every step."""
    import torch
    from ipcdnnwalk_b200 import SRBVecEnv
    n = 7 if host else 6                                   # 7: a ragged last warp in the coalesced-I/O kernel
    oracles, plans, rng = _oracles(n, "test", seed=300 + lanes)
    env = SRBVecEnv(helpers.fresh_clips(), n, variant="test", precision=32, lanes_per_env=lanes, auto_reset=False)
    env.reset(plan=plans)
    torch.cuda.synchronize()
    for k in range(20):
        for i, o in enumerate(oracles):
            env.set_state(i, helpers.oracle_to_cuda_state(o))
        acts = np.stack([helpers.tracking_action(o, rng, 0.1) for o in oracles])
        if host:
            obs, rew, done, _ = env.step_host(acts)
            obs, rew, done = obs.copy(), rew.copy(), done.copy()
        else:
            obs, rew, done, _ = env.step(torch.as_tensor(acts, device="cuda"))
            torch.cuda.synchronize()
            obs, rew, done = obs.cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy()
        dones = _compare_step(env, oracles, acts, obs, rew, done, 1e-3, f"fp32 lanes {lanes} host {host} step {k}", rew_rtol=2e-3)
        for i, o in enumerate(oracles):
            if dones[i]:
                helpers.oracle_reset(o, plans[i])
    env.close()


@pytest.mark.parametrize("lanes", [4, 8])
@pytest.mark.parametrize("planar", [(22.0, 24.0), (-42.0, 42.0)])
def test_single_step_fp32_terrain(built_lib, lanes, planar):
    """The fp32 terrain kernel (4 lanes = the instantiation behind the config-3 numbers; 8 lanes) on the pyramid and on the
    hill height field: state / obs within 1e-3 with the state reloaded from the oracle every step; flags, frame indices and
    done bit exact.  Terrain heights enter the state, so the bound covers the fp32 height-field interpolation."""
    import torch
    from ipcdnnwalk_b200 import SRBVecEnv, TerrainModel
    from oracle.terrain import Terrain
    from oracle.srb_env import OracleSRBEnv, lift_clip_to_terrain
    T = Terrain.load(NPZ)
    n = 4
    rng = np.random.default_rng(17)
    frames = [0, 40, 120, 233]
    env = SRBVecEnv(helpers.fresh_clips(("aiming_show",)), n, variant="terrain", precision=32, lanes_per_env=lanes, auto_reset=False,
                    terrain=TerrainModel.from_npz(NPZ), initial_pos_planar=np.tile(np.array(planar), (n, 1)))
    plans = np.stack([np.concatenate([[0, f], rng.uniform(-.05, .05, 3), [1.0, 0, 0, 0]]) for f in frames])
    env.reset(plan=plans)
    oracles = []
    for i in range(n):
        o = OracleSRBEnv([lift_clip_to_terrain(helpers.load_clip("aiming_show"), T, planar)], "terrain", terrain=T)
        helpers.oracle_reset(o, plans[i])
        oracles.append(o)
    for k in range(15):
        for i, o in enumerate(oracles):
            st = helpers.oracle_to_cuda_state(o)
            st["reals"][44:46] = planar
            env.set_state(i, st)
        acts = np.stack([helpers.tracking_action(o, rng, 0.05) for o in oracles])
        if k % 5 == 2:
            acts[:, 18:22] = 0; acts[:, 18 + 2] = 1
        if k % 5 == 4:
            acts[:, 18:22] = 0; acts[:, 18 + 3] = 1
        obs, rew, done, _ = env.step(torch.as_tensor(acts, device="cuda"))
        torch.cuda.synchronize()
        obs = obs.cpu().numpy()
        for i, o in enumerate(oracles):
            oo, rr, dd, _, _ = o.step(acts[i])
            st = helpers.cuda_state_vector(env.named_state(i)); ref = helpers.oracle_state_vector(o)
            assert (st[23:26] == ref[23:26]).all() and bool(done[i]) == dd
            _close(st[:23], ref[:23], 1e-3, f"fp32 terrain state step {k} env {i}")
            _close(obs[i], oo, 1e-3, f"fp32 terrain obs step {k} env {i}")
    env.close()


# ---------------------------------------------------------------------------------------------------------------------
def _sample_check(env, clips, plans, sample, variant, steps, rtol, sigma, make_oracle=None):
    """Drive the whole batch with the device tracking controller; replay the sampled environments in the oracle with the
    very same float32 actions and compare after every step."""
    import torch
    from oracle.srb_env import OracleSRBEnv
    oracles = {}
    for i in sample:
        o = make_oracle(i) if make_oracle else OracleSRBEnv(helpers.fresh_clips(), variant)
        helpers.oracle_reset(o, plans[i])
        oracles[i] = o
    act = torch.zeros(env.num_envs, 22, device="cuda")
    alive = set(sample)
    for k in range(steps):
        env.tracking_action(act, sigma=sigma, seed=99, step=k)
        a = act[torch.as_tensor(sample, device="cuda")].cpu().numpy()
        obs, rew, done, _ = env.step(act)
        torch.cuda.synchronize()
        idx = torch.as_tensor(sample, device="cuda")
        o_s = obs[idx].cpu().numpy(); r_s = rew[idx].cpu().numpy(); d_s = done[idx].cpu().numpy()
        for j, i in enumerate(sample):
            if i not in alive:
                continue
            o = oracles[i]
            oo, rr, dd, _, _ = o.step(a[j])
            st = helpers.cuda_state_vector(env.named_state(i)); ref = helpers.oracle_state_vector(o)
            assert (st[23:26] == ref[23:26]).all() and bool(d_s[j]) == dd, (k, i)
            _close(st[:23], ref[:23], rtol, f"state step {k} env {i}")
            _close(o_s[j], oo, rtol, f"obs step {k} env {i}")
            _close(float(r_s[j]), rr, 2 * rtol, f"reward step {k} env {i}")
            if dd:
                alive.discard(i)                       # auto_reset is off: a finished environment is not compared further
    return len(alive)


@pytest.mark.parametrize("precision,rtol", [(64, 1e-5), (32, 5e-3)])
def test_sample_at_bench_size_flat_4096(built_lib, precision, rtol):
    """BASELINE configs[1] batch (4096 envs, default lane count): 64 random environments of the full batch against the
    oracle over 8 steps.  fp64 at the single-step tolerance (1e-5); fp32 free-running for 8 steps, so the bound is the
    accumulated one (5e-3), with flags / frame indices / done bit exact."""
    from ipcdnnwalk_b200 import SRBVecEnv
    n = 4096
    clips = helpers.fresh_clips()
    rng = np.random.default_rng(4096)
    plans = np.stack([helpers.random_plan(rng, clips) for _ in range(n)])
    env = SRBVecEnv(clips, n, variant="test", precision=precision, auto_reset=False)
    env.reset(plan=plans)
    sample = sorted(rng.choice(n, 64, replace=False).tolist() + [0, n - 1])
    left = _sample_check(env, clips, plans, sample, "test", 8, rtol, 0.1)
    assert left >= 32
    env.close()


def test_sample_at_bench_size_two_lanes_32768(built_lib):
    """32768 envs per GPU with the 2-lane fp32 instantiation (the round-1 8-GPU 1.8e9 figure; since round 2 the default at this size
    is 4 lanes, covered by test_sample_at_bench_size_default_lanes_32768): 32 sampled environments, 6 steps."""
    from ipcdnnwalk_b200 import SRBVecEnv
    n = 32768
    clips = helpers.fresh_clips()
    rng = np.random.default_rng(32768)
    plans = np.stack([helpers.random_plan(rng, clips) for _ in range(n)])
    env = SRBVecEnv(clips, n, variant="test", precision=32, auto_reset=False, lanes_per_env=2)
    env.reset(plan=plans)
    sample = sorted(set(rng.choice(n, 32, replace=False).tolist() + [0, n - 1]))
    _sample_check(env, clips, plans, sample, "test", 6, 5e-3, 0.1)
    env.close()


def test_sample_at_bench_size_default_lanes_32768(built_lib):
    """32768 envs per GPU with the library's default lane count (4: the structured 4-lane QP): 32 sampled environments, 6 steps."""
    from ipcdnnwalk_b200 import SRBVecEnv
    n = 32768
    clips = helpers.fresh_clips()
    rng = np.random.default_rng(32769)
    plans = np.stack([helpers.random_plan(rng, clips) for _ in range(n)])
    env = SRBVecEnv(clips, n, variant="test", precision=32, auto_reset=False)
    env.reset(plan=plans)
    sample = sorted(set(rng.choice(n, 32, replace=False).tolist() + [0, n - 1]))
    _sample_check(env, clips, plans, sample, "test", 6, 5e-3, 0.1)
    env.close()


@pytest.mark.parametrize("precision,rtol", [(64, 1e-5), (32, 5e-3)])
def test_sample_at_bench_size_terrain_8192(built_lib, precision, rtol):
    """BASELINE configs[2] per-GPU share (8192 terrain envs spread over the four playground quadrants, default lanes):
    32 sampled environments against the oracle over 6 steps."""
    from ipcdnnwalk_b200 import SRBVecEnv, TerrainModel
    from oracle.terrain import Terrain
    from oracle.srb_env import OracleSRBEnv, lift_clip_to_terrain
    T = Terrain.load(NPZ)
    n = 8192
    clips = helpers.fresh_clips(("aiming_show",))
    rng = np.random.default_rng(8192)
    quad = rng.integers(0, 4, n)
    centre = np.array([[42.0, 42.0], [-42.0, 42.0], [42.0, -42.0], [-42.0, -42.0]])[quad]
    planar = centre + rng.uniform(-30, 30, (n, 2))
    planar[:8] = [[22.0, 24.0], [20.5, 21.0], [-42.0, 42.0], [42.0, -40.0], [0.0, 0.0], [30.0, 12.0], [-60.0, 50.0], [50.0, -20.0]]
    plans = np.stack([np.concatenate([[0, rng.integers(0, 300)], rng.uniform(-.05, .05, 3), [1.0, 0, 0, 0]]) for _ in range(n)])
    env = SRBVecEnv(clips, n, variant="terrain", precision=precision, auto_reset=False, terrain=TerrainModel.from_npz(NPZ), initial_pos_planar=planar)
    env.reset(plan=plans)
    sample = sorted(set(list(range(8)) + rng.choice(n, 24, replace=False).tolist()))

    def make(i):
        return OracleSRBEnv([lift_clip_to_terrain(helpers.load_clip("aiming_show"), T, tuple(planar[i]))], "terrain", terrain=T)
    _sample_check(env, clips, plans, sample, "terrain", 6, rtol, 0.05, make_oracle=make)
    env.close()
