"""Every kernel family once at ragged batch sizes (last warp / CTA partly filled), all lane counts, device and host paths.
Written for compute-sanitizer (memcheck / racecheck), which is closed on this GPU pool; runs plainly as a crash test."""
import sys, os, numpy as np, torch
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import helpers, helpers_mmsto as hm
from ipcdnnwalk_b200 import SRBVecEnv, ShortTrajOpt, CDMVecEnv, Skeleton, vrml_skeleton, TerrainModel
from ipcdnnwalk_b200.short_traj_opt import extractSingleFrameFullbody
import helpers_cdm as hc
skel = hm.skeleton()
# MMSTO: 5 windows (ragged warp), 3 iterations, plus evaluate, comslide, extract
ps = [hm.make_problem(s, skel=skel) for s in range(5)]
opt = ShortTrajOpt(skel, hm.NF, hm.MARKERS); opt.set("max_iterations", 3)
K = ("gpos_target", "x_original", "dx", "ddx", "com", "momentum")
x, info = opt.solveEndEffectors_euler(hm.stack(ps, "euler_mot"), *[hm.stack(ps, k) for k in K], con=hm.stack(ps, "con"), velcon=hm.stack(ps, "velcon"))
opt.evaluate(hm.stack(ps, "x_original"), *[hm.stack(ps, k) for k in K])
poses = torch.as_tensor(opt.eulerToPoseDOF(x.cpu().numpy()), device="cuda").contiguous()
opt.removeCOMsliding_online(hm.stack(ps, "com"), hm.stack(ps, "com"), poses)
extractSingleFrameFullbody(np.random.default_rng(0).normal(size=(37, 29)), np.random.default_rng(1).normal(size=(37, 214)), 60.0, np.eye(3))
torch.cuda.synchronize(); print("mmsto ok", info[:, 1].tolist())
# SRB: flat (device + host path), terrain, 33 envs (ragged warp), all lane counts
clips = helpers.fresh_clips()
for lanes in (8, 4, 2, 1):
    env = SRBVecEnv(clips, 33, variant="test", precision=32, lanes_per_env=lanes, auto_reset=True, seed=1); env.reset()
    act = torch.zeros(33, 22, device="cuda")
    for k in range(3):
        env.tracking_action(act, sigma=0.1, seed=3, step=k); env.step(act)
    env.pack_obs_extra()
    if lanes in (8, 4):
        env.pinned()["act"][...] = act.cpu().numpy(); env.step_host()
    torch.cuda.synchronize(); env.close()
npz = os.path.join(helpers.GOLDEN, "terrain_playground.npz")
env = SRBVecEnv(helpers.fresh_clips(("aiming_show",)), 33, variant="terrain", precision=32, auto_reset=True, terrain=TerrainModel.from_npz(npz),
                initial_pos_planar=np.tile(np.array([[22.0, 24.0]]), (33, 1))); env.reset()
act = torch.zeros(33, 22, device="cuda")
for k in range(3):
    env.tracking_action(act, sigma=0.1, seed=3, step=k); env.step(act)
torch.cuda.synchronize(); env.close(); print("srb ok")
tab, sk = hc.load_tables()
for prec in (32, 64):
    env = CDMVecEnv(tab, sk, 70, precision=prec, lanes_per_env=1, auto_reset=True, seed=3); env.reset()
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    for k in range(4):
        env.step(torch.rand(70, 10, device="cuda", generator=g) * 2 - 1)
    torch.cuda.synchronize(); env.close()
print("cdm ok")
s = Skeleton(vrml_skeleton.builtin())
for dt in (torch.float32, torch.float64):
    pose = torch.randn(200, 60, device="cuda", dtype=dt); pose[:, 3:7] /= pose[:, 3:7].norm(dim=1, keepdim=True)
    s.forward(pose.contiguous(), torch.randn(200, 59, device="cuda", dtype=dt))
torch.cuda.synchronize(); s.close(); print("fk ok")
