"""Shared test infrastructure of the MMSTO path: synthetic windows in the shapes walk_SRBTrack2025.py:505-587 builds."""
import numpy as np

from ipcdnnwalk_b200 import vrml_skeleton
from oracle import mmsto as mm

NF = 7
# hanyang_lowdof_T.con_config.lua: toes / ankles both on LeftAnkle (bone 3) / RightAnkle (bone 6); marker weight 100 (walk_SRBTrack2025.py:338)
MARKERS = [dict(bone=3, lpos=[0, -0.12, 0.14], weight=100.0), dict(bone=3, lpos=[0, -0.12, -0.07], weight=100.0),
           dict(bone=6, lpos=[0, -0.12, 0.14], weight=100.0), dict(bone=6, lpos=[0, -0.12, -0.07], weight=100.0)]


def skeleton():
    return vrml_skeleton.builtin()


def make_problem(seed, nf=NF, target_noise=0.02, velcon=True, skel=None):
    """One window: a smooth desired Euler motion, targets taken from a perturbed copy of it -> dict of arrays (no batch dimension)."""
    skel = skel or skeleton()
    rng = np.random.default_rng(seed)
    t = mm.EulerTree(skel)
    nd = t.ndof
    q0 = rng.uniform(-.3, .3, nd); q0[1] += 0.95
    vel = rng.uniform(-.5, .5, nd); vel[0] = 1.0
    acc = rng.uniform(-1, 1, nd)
    X = np.stack([q0 + vel * (i / 30) + 0.5 * acc * (i / 30) ** 2 for i in range(nf)])
    Xt = X + rng.normal(0, target_noise, X.shape)
    tr = [mm.EulerTree(skel) for _ in range(nf)]
    for i in range(nf):
        tr[i].set_q(Xt[i])
    gpos = np.stack([np.concatenate([tr[i].global_pos(m["bone"], m["lpos"]) for m in MARKERS]) for i in range(nf)])
    com = np.stack([tr[i].calc_com() for i in range(nf)])
    mom = []
    for i in range(nf - 1):
        tr[i].set_dq((Xt[i + 1] - Xt[i]) * 30)
        mom.append(tr[i].calc_momentum_com())
    dx = (X[1:] - X[:-1]) * 30
    p = dict(x_original=X, dx=dx, ddx=(dx[1:] - dx[:-1]) * 30, gpos_target=gpos, com=com, momentum=np.stack(mom),
             con=rng.uniform(size=(4, nf)) > 0.5, euler_mot=X + rng.normal(0, 0.01, X.shape))
    vc = np.zeros((4, 10)); vc[:, 0] = -1
    if velcon:
        vc[0] = [2, 13, 0.8, 0, 1.0, 0, 0.05, 0, 0, 0]              # left wrist must move up at 0.8 m/s at frame 2
    p["velcon"] = vc
    return p


def oracle_opt(p, skel=None, dof_weights=None, **weights):
    skel = skel or skeleton()
    nf = p["x_original"].shape[0]
    o = mm.ShortTrajOpt(skel, nf, MARKERS, **weights)
    o.x_original = p["x_original"].copy(); o.dx = p["dx"].copy(); o.ddx = p["ddx"].copy()
    o.gpos_target = p["gpos_target"]; o.com = p["com"]; o.momentum = p["momentum"]
    o.con = [p["con"][m] for m in range(4)]
    o.hs_vel = {}
    for row in p["velcon"]:
        if row[0] >= 0:
            o.hs_vel.setdefault(int(row[0]), []).append((row[2], row[3:6], int(row[1]), row[6:9]))
    if dof_weights is not None:
        o.dof_weights = np.asarray(dof_weights, dtype=float).copy()
    return o


def ankle_dof_weights(skel=None):
    """walk_SRBTrack2025.py:370-376: ankle joints more flexible."""
    t = mm.EulerTree(skel or skeleton())
    w = np.ones(t.ndof)
    w[t.bone_dofs[3]] = 0.1; w[t.bone_dofs[6]] = 0.1
    return w


def stack(problems, key):
    return np.stack([p[key] for p in problems])
