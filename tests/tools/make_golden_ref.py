"""Record what the REFERENCE's own taesooLib code computes (oracle/_ref, built from /root/reference by oracle/ref_build/Makefile)
as committed fixtures, so that the pins travel to boxes without /root/reference:

    make -C oracle/ref_build && python tests/tools/make_golden_ref.py      ->  tests/golden/ref_taesoolib.npz

Contents (inputs + reference outputs; nothing here is produced by oracle/ or by the CUDA kernels):
  fk_*      BoneForwardKinematics frames, quaternion-root LoaderToTree frames / CoM / centroidal momentum / link velocities  (a-22, a-23)
  eu_*      Euler-root LoaderToTree (ShortTrajOpt's tree): frames, CoM, momentum, CoM / momentum / marker / rotation Jacobians,
            updateGrad_S_JT                                                                                              (f-1)
  q_*       quater / vector3 / transf / Liegroup primitives                                                              (a-19, a-22, f-2)
  tr_*      OBJloader::Terrain::height on the gym_cdm2 height map (64 x 64 after the loader's decimation), tiled and not (a-21)
  hull_*    GrahamScan hulls (PhysicsLib/convexhull/graham.cpp) of height profiles: what math.projectTrajectoryToItsHull builds on   (f-2)
  dof_*     MotionDOFcontainer read of hanyang_lowdof_T_lafan1_walk_2915.dof                                              (f-4)
  (ref_taesoolib_signal.npz, `python tests/tools/make_golden_ref.py signal`)
  sg_*      Filter::GetGaussFilter / gaussFilter, NonuniformSpline (zero end accelerations) curve + second derivative, matrixn::sampleRow --
            the signal operators under fc.CDMTraj, on the CoM track of the walking clip (33 frames) and on a random walk; MotionDOF::samplePose on the clip   (a-20, f-3)
  lb_*      libLBFGS (reference lbfgs.cpp, parameters as optimize.h:426-470 sets them) on its Rosenbrock sample and on MMSTO windows
            whose objective / gradient callback is oracle/mmsto.py (the missing Lua wrapper's role): status code, iteration and
            evaluation counts, per-iteration fx / step / line-search count, final x                                      (f-1)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "tools")):
    sys.path.insert(0, p)
import ref_harness as rh                      # noqa: E402
import helpers_mmsto as hm                    # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
EU_MARKERS = [(4, (0.0, -0.12, 0.14)), (7, (0.0, -0.12, -0.07)), (14, (0.1, 0.02, 0.03)), (20, (0.0, 0.1, 0.0))]   # tree indices (1-based)
TERRAIN_RAW = "/root/reference/work/taesooLib/Resource/motion/crowdEditingScripts/terrain/heightmap1_256_256_2bytes.raw"
DOF = "/root/reference/work/taesooLib/Resource/motion/MOB1/hanyang_lowdof_T_lafan1_walk_2915.dof"


def rosenbrock(x):
    """The sample objective shipped with libLBFGS (sample/sample.c)."""
    g = np.zeros_like(x); f = 0.0
    for i in range(0, len(x), 2):
        t1 = 1.0 - x[i]; t2 = 10.0 * (x[i + 1] - x[i] * x[i])
        g[i + 1] = 20.0 * t2; g[i] = -2.0 * (x[i] * g[i + 1] + t1); f += t1 * t1 + t2 * t2
    return f, g


def mmsto_window(seed):
    """x0 (free part), evaluate(z) of one MMSTO window exactly as ShortTrajOpt.solve builds them (oracle/mmsto.py:324-338)."""
    p = hm.make_problem(seed)
    o = hm.oracle_opt(p, dof_weights=hm.ankle_dof_weights())
    o.mcache = None; o.n_eval = 0
    x0 = o.initial_guess(np.asarray(p["euler_mot"], dtype=float)).ravel()
    nfix = 2 * o.ndof

    def ev(z):
        x = x0.copy(); x[nfix:] = z
        f, g = o.evaluate(x)
        return f, g[nfix:]
    return x0, nfix, ev


def pack_lbfgs(prefix, res, out):
    out[prefix + "ret"] = res["ret"]; out[prefix + "x"] = res["x"]; out[prefix + "fx"] = res["fx"]; out[prefix + "evaluations"] = res["evaluations"]
    out[prefix + "iter_fx"] = np.array([i["fx"] for i in res["iterations"]]); out[prefix + "iter_step"] = np.array([i["step"] for i in res["iterations"]])
    out[prefix + "iter_ls"] = np.array([i["ls"] for i in res["iterations"]]); out[prefix + "iter_gnorm"] = np.array([i["gnorm"] for i in res["iterations"]])


def main():
    assert rh.available(), "build oracle/_ref first: make -C oracle/ref_build"
    out = {}
    rng = np.random.default_rng(20251020)
    # ---- fk
    n = 48
    pose = np.zeros((n, 60)); pose[:, :3] = rng.uniform(-5, 5, (n, 3))
    q = rng.standard_normal((n, 4)); pose[:, 3:7] = q / np.linalg.norm(q, axis=1, keepdims=True)
    pose[:, 7:] = rng.uniform(-1, 1, (n, 53)); pose[0, 7:] = 0.0; pose[1, 3:7] = [1, 0, 0, 0]
    dq = rng.uniform(-2, 2, (n, 59)); dq[2] = 0.0
    r = rh.fk(pose, dq)
    out.update(fk_pose=pose, fk_dq=dq, fk_bfk=r["bfk"], fk_tree=r["tree"], fk_com=r["com"], fk_mom=r["mom"], fk_angvel=r["angvel"], fk_linvel=r["linvel"])
    # ---- euler tree
    n = 6
    eq = rng.uniform(-1, 1, (n, 59)); eq[:, :3] *= 3; edq = rng.uniform(-2, 2, (n, 59)); dS = rng.uniform(-1, 1, (n, 3))
    res = rh.euler(eq, edq, dS, EU_MARKERS)
    out.update(eu_q=eq, eu_dq=edq, eu_dS=dS, eu_marker_bone=np.array([b for b, _ in EU_MARKERS]), eu_marker_lpos=np.array([l for _, l in EU_MARKERS]),
               eu_frames=np.stack([d["frames"] for d in res]), eu_com=np.stack([d["com"] for d in res]), eu_mom=np.stack([d["mom"] for d in res]),
               eu_JT_com=np.stack([d["JT_com"] for d in res]), eu_JT_mom=np.stack([d["JT_mom"] for d in res]),
               eu_mk_pos=np.stack([[m["pos"] for m in d["markers"]] for d in res]), eu_mk_JT=np.stack([[m["JT"] for m in d["markers"]] for d in res]),
               eu_mk_JT_rot=np.stack([[m["JT_rot"] for m in d["markers"]] for d in res]), eu_mk_grad=np.stack([[m["grad"] for m in d["markers"]] for d in res]))
    # ---- primitives
    n = 64
    rows = np.concatenate([rng.standard_normal((n, 8)), rng.uniform(-1.5, 1.5, (n, 3)), rng.uniform(0.05, 0.95, (n, 1))], axis=1)
    rows[0, 0:4] = [1, 0, 0, 0]; rows[1, 0:4] = [0.99999999, 1e-9, 0, 0]; rows[2, 4:8] = -rows[2, 0:4] + 1e-3
    Q = rh.quat(rows)
    out["q_rows"] = rows
    for k, v in Q.items():
        out["q_" + k] = v
    # ---- terrain
    img = np.fromfile(TERRAIN_RAW, dtype="<u2", count=256 * 256).reshape(256, 256)
    dec = img[::4, ::4][:64, :64].copy()                    # the loader's BAD_NOTEBOOK decimation (Terrain.cpp:488-503)
    dec[-1] = dec[0]                                        # what the loader's tiling branch does last (Terrain.cpp:427-428)
    xz = np.concatenate([rng.uniform(0, 4800, (400, 2)), rng.uniform(-6000, 12000, (200, 2)),
                         np.array([[0, 0], [4800, 4800], [76.19047619047619 * 3, 76.19047619047619 * 5], [100, 4800 - 1e-9], [4799.999, 10]])])
    out.update(tr_image=dec, tr_xz=xz, tr_size=np.array([4800.0, 4800.0, 100.0]))
    for tile in (1, 0):
        hm_, h, h2, nrm = rh.terrain(dec, 4800.0, 4800.0, 100.0, tile, xz)
        out[f"tr_heightmap"] = hm_; out[f"tr_height_tile{tile}"] = h; out[f"tr_height2_tile{tile}"] = h2; out[f"tr_normal_tile{tile}"] = nrm
    # ---- .dof reader
    out["dof_walk_2915"] = rh.dof(DOF)
    # ---- GrahamScan (PhysicsLib/convexhull/graham.cpp) on height profiles over the frame index, incl. ties and collinear runs
    sets = []
    for k in range(160):
        npt = int(rng.integers(3, 10))
        y = np.round(rng.uniform(0, 1, npt) * [1, 4][k % 2]) / [1, 4][k % 2] if k % 3 == 0 else rng.uniform(0, 1, npt)
        sets.append(np.stack([np.arange(npt, dtype=float), y], 1))
    sets.append(np.stack([np.arange(7.0), np.zeros(7)], 1)); sets.append(np.stack([np.arange(7.0), [0, 0, 0.2, 0.2, 0.4, 0.4, 0.4]], 1))
    hulls = rh.hull(sets)
    out["hull_count"] = len(sets)
    for k, (a, b) in enumerate(zip(sets, hulls)):
        out[f"hull_in_{k}"] = a; out[f"hull_out_{k}"] = b
    # ---- libLBFGS
    d = rh.lbfgs_defaults()
    out["lb_defaults_keys"] = np.array(list(d.keys())); out["lb_defaults"] = np.array(list(d.values()))
    x0 = np.zeros(100); x0[0::2] = -1.2; x0[1::2] = 1.0
    pack_lbfgs("lb_rosen_", rh.lbfgs(x0, rosenbrock, epsilon=1e-5), out)
    out["lb_rosen_x0"] = x0
    for seed, max_it in ((3, 12), (21, 10), (22, 0)):       # 0 = until libLBFGS terminates by itself (the online configuration)
        z0, nfix, ev = mmsto_window(seed)
        res = rh.lbfgs(z0[nfix:].copy(), ev, epsilon=1e-9, max_iterations=max_it if max_it else -1)
        pack_lbfgs(f"lb_mmsto{seed}_", res, out)
        out[f"lb_mmsto{seed}_max_iterations"] = max_it
        print("mmsto window", seed, "ret", res["ret"], "iterations", len(res["iterations"]), "evaluations", res["evaluations"], "fx", res["fx"])
    np.savez_compressed(os.path.join(GOLDEN, "ref_taesoolib.npz"), **out)
    print("wrote tests/golden/ref_taesoolib.npz:", {k: np.asarray(v).shape for k, v in out.items()})


def main_signal():
    """Inputs + reference outputs of the signal operators (separate file so that ref_taesoolib.npz stays as recorded)."""
    assert rh.available(), "build oracle/_ref first: make -C oracle/ref_build"
    from ipcdnnwalk_b200 import cdm_tables as ct
    from ipcdnnwalk_b200.vrml_skeleton import parse_wrl
    rng = np.random.default_rng(20251021)
    skel = parse_wrl(rh.WRL)
    mot = ct.read_dof(DOF)                                  # one cycle, 33 frames
    com = np.stack([ct._com(skel, *ct.fk(skel, mot[i])) for i in range(len(mot))])
    walk = np.cumsum(rng.standard_normal((90, 4)), 0)
    out = {}
    for name, M, delta in (("com", com, 7), ("rw", walk, 5)):
        n = M.shape[0]
        ks = [16, 17, 12, 13, 5, 20] + ([63] if n > 63 else [])     # LTIFilter refuses kernels longer than the signal
        t = np.concatenate([rng.uniform(0, n - 1, 40), np.floor(rng.uniform(0, n - 2, 12)) + rng.choice([0.004, 0.0049, 0.0051, 0.006, 0.994, 0.9949, 0.9951, 0.996], 12),
                            [0.0, n - 1.0, n - 1.4, n - 1.003, 3.3, 10.7, 0.1 + 0.2, 1.0 / 3.0]])
        r = rh.signal(M, ks, delta, t)
        out.update({f"sg_{name}_in": M, f"sg_{name}_kernel_sizes": np.array(ks), f"sg_{name}_delta": np.array(delta), f"sg_{name}_times": t,
                    f"sg_{name}_timing": r["timing"], f"sg_{name}_curve": r["curve"], f"sg_{name}_acc": r["acc"], f"sg_{name}_samples": r["samples"]})
        for k, ker, f in zip(ks, r["kernels"], r["filtered"]):
            out[f"sg_{name}_kernel_{k}"] = ker; out[f"sg_{name}_filtered_{k}"] = f
    # MotionDOF::samplePose / MotionDOFinfo::blend on the walking clip (root quaternion by safeSlerp, everything else linear, double fraction)
    nfr = mot.shape[0]
    tp = np.concatenate([rng.uniform(0, nfr - 1, 30), np.floor(rng.uniform(0, nfr - 2, 8)) + rng.choice([0.0049, 0.0051, 0.9949, 0.9951], 8), [0.0, nfr - 1.0, nfr - 1.5, 5.5, 1.0 / 3.0]])
    out["sg_pose_times"] = tp; out["sg_pose_rows"] = rh.samplepose(DOF, tp)
    np.savez_compressed(os.path.join(GOLDEN, "ref_taesoolib_signal.npz"), **out)
    print("wrote tests/golden/ref_taesoolib_signal.npz:", len(out), "arrays")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "signal":
        main_signal()
    else:
        main()
