"""Python driver of oracle/_ref/taesoo_ref and oracle/_ref/liblbfgs_ref.so -- the REFERENCE's own taesooLib code compiled here
by oracle/ref_build/Makefile.  Used by tests/tools/make_golden_ref.py (fixture generation) and by the live pin tests, which skip
when oracle/_ref has not been built (the GPU box has no /root/reference to build it from; the committed fixtures cover it)."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
BIN = os.path.join(REF_DIR, "taesoo_ref")
LBFGS_SO = os.path.join(REF_DIR, "liblbfgs_ref.so")
WRL = "/root/reference/gym_trackSRB/taesooLib/hanyang_lowdof_T_sh.wrl"


def available():
    return os.path.exists(BIN) and os.path.exists(LBFGS_SO)


class _Cursor:
    def __init__(self, a):
        self.a, self.p = a, 0

    def take(self, n, shape=None):
        v = self.a[self.p:self.p + n]
        assert v.shape[0] == n, "reference harness output is shorter than expected"
        self.p += n
        return v.reshape(shape) if shape else v

    def scalar(self):
        return float(self.take(1)[0])

    def mat(self):
        r, c = int(self.scalar()), int(self.scalar())
        return self.take(r * c, (r, c))

    def vec(self):
        return self.take(int(self.scalar()))


def _run(args, inputs):
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.txt"), os.path.join(td, "out.txt")
        np.savetxt(fin, np.asarray(inputs, dtype=np.float64).reshape(-1), fmt="%.17g")
        cmd = [BIN] + [a if a not in ("@in", "@out") else (fin if a == "@in" else fout) for a in args]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, cwd=td, timeout=600)
        if r.returncode != 0:
            raise RuntimeError(f"taesoo_ref failed ({r.returncode}): {r.stdout.decode()[-2000:]}")
        return _Cursor(np.loadtxt(fout, dtype=np.float64).reshape(-1))


def fk(pose, dq, wrl=WRL):
    """BoneForwardKinematics + quaternion-root LoaderToTree of the reference.  pose [n,60], dq [n,59]."""
    pose = np.asarray(pose, dtype=np.float64); dq = np.asarray(dq, dtype=np.float64)
    c = _run(["fk", wrl, "@in", "@out"], np.concatenate([pose, dq], axis=1))
    n, nb, nd = int(c.scalar()), int(c.scalar()), int(c.scalar())
    out = dict(bfk=np.zeros((n, nb, 7)), tree=np.zeros((n, nb, 7)), com=np.zeros((n, 3)), mom=np.zeros((n, 6)), angvel=np.zeros((n, nb, 3)), linvel=np.zeros((n, nb, 3)))
    for r in range(n):
        out["bfk"][r] = c.take(nb * 7, (nb, 7)); out["tree"][r] = c.take(nb * 7, (nb, 7))
        out["com"][r] = c.take(3); out["mom"][r] = c.take(6)
        wv = c.take(nb * 6, (nb, 6))
        out["angvel"][r] = wv[:, :3]; out["linvel"][r] = wv[:, 3:]
    return out


def euler(q, dq, deltaS, markers, wrl=WRL):
    """Euler-root LoaderToTree (ShortTrajOpt's tree).  q, dq [n,59]; deltaS [n,3]; markers = [(bone tree index, (lx,ly,lz)), ...]."""
    q = np.asarray(q, dtype=np.float64)
    args = ["euler", wrl, "@in", "@out", str(len(markers))]
    for b, l in markers:
        args += [str(int(b))] + ["%.17g" % x for x in l]
    c = _run(args, np.concatenate([q, dq, deltaS], axis=1))
    n, nb, nq, nm = int(c.scalar()), int(c.scalar()), int(c.scalar()), int(c.scalar())
    res = []
    for r in range(n):
        d = dict(frames=c.take(nb * 7, (nb, 7)), com=c.take(3), mom=c.take(6), JT_com=c.mat(), JT_mom=c.mat(), markers=[])
        for k in range(nm):
            d["markers"].append(dict(pos=c.take(3), JT=c.mat(), JT_rot=c.mat(), grad=c.vec()))
        d["q_back"] = c.take(nq); d["dq_back"] = c.take(nq)
        res.append(d)
    return res


QUAT_FIELDS = (("q1", 4), ("q2", 4), ("mult", 4), ("inverse", 4), ("rotate", 3), ("euler_YZX", 4), ("euler_XYZ", 4), ("euler_ZXY", 4), ("euler_X", 4),
               ("axis_angle", 4), ("exp", 4), ("log", 3), ("rotation_vector", 3), ("rotation_angle", 1), ("difference", 4), ("safe_slerp", 4),
               ("rot_y", 4), ("offset", 4), ("no_twist", 4), ("twist_q", 4), ("axis_to_axis", 4), ("matrix", 9), ("from_matrix", 4),
               ("tf_mult", 7), ("tf_inverse", 7), ("tf_apply", 3), ("twist", 6), ("inertia_times_se3", 6), ("dAd", 6), ("inv_dAd", 6), ("se3_log", 6),
               ("se3_exp", 7), ("sop_map", 1), ("sop_clamp_map", 1))


def quat(rows):
    """rows [n,12] = q1 wxyz | q2 wxyz | v xyz | t  -> dict of arrays, one entry per primitive (see oracle/ref_build/ref_harness.cpp)."""
    c = _run(["quat", "@in", "@out"], rows)
    n = int(c.scalar())
    out = {k: np.zeros((n, w)) for k, w in QUAT_FIELDS}
    for r in range(n):
        for k, w in QUAT_FIELDS:
            out[k][r] = c.take(w)
    return out


def terrain(image, size_x, size_z, height_max, tile, xz):
    """OBJloader::Terrain built from a uint16 image (no seam blending) -> (mHeight, height(x) [n], height(x, normal) [n], normals [n,3])."""
    image = np.asarray(image)
    xz = np.asarray(xz, dtype=np.float64).reshape(-1, 2)
    inp = np.concatenate([[image.shape[0], image.shape[1], size_x, size_z, height_max, int(tile)], image.astype(np.float64).reshape(-1), [xz.shape[0]], xz.reshape(-1)])
    c = _run(["terrain", "@in", "@out"], inp)
    n = int(c.scalar())
    hm = c.mat()
    hn = c.take(5 * n, (n, 5))
    return hm, hn[:, 0].copy(), hn[:, 1].copy(), hn[:, 2:].copy()


def hull(point_sets):
    """GrahamScan of the reference on every point set ([n,2] arrays) -> list of closed hull polygons [m,2]."""
    inp = []
    for pts in point_sets:
        pts = np.asarray(pts, dtype=np.float64).reshape(-1, 2)
        inp += [pts.shape[0]] + pts.reshape(-1).tolist()
    c = _run(["hull", "@in", "@out"], inp)
    return [c.mat().copy() for _ in point_sets]


def signal(mat, kernel_sizes, delta, sample_times):
    """Filter::GetGaussFilter / gaussFilter per kernel size, NonuniformSpline (zero end accelerations) through every `delta`-th row
    evaluated at the unit frames (curve, second derivative), matrixn::sampleRow at `sample_times`."""
    mat = np.asarray(mat, dtype=np.float64)
    inp = np.concatenate([[mat.shape[0], mat.shape[1]], mat.reshape(-1), [len(kernel_sizes)], np.asarray(kernel_sizes, dtype=np.float64), [delta],
                          [len(sample_times)], np.asarray(sample_times, dtype=np.float64)])
    c = _run(["signal", "@in", "@out"], inp)
    out = {"kernels": [], "filtered": []}
    for _ in kernel_sizes:
        out["kernels"].append(c.vec().copy()); out["filtered"].append(c.mat().copy())
    out["timing"] = c.vec().copy(); out["curve"] = c.mat().copy(); out["acc"] = c.mat().copy()
    n = int(c.scalar())
    out["samples"] = np.stack([c.vec().copy() for _ in range(n)])
    return out


def samplepose(path, times, wrl=WRL):
    """MotionDOF::samplePose of the clip at `times` -> [n, numDOF]."""
    c = _run(["samplepose", wrl, path, "@in", "@out"], np.asarray(times, dtype=np.float64))
    n, k = int(c.scalar()), int(c.scalar())
    return c.take(n * k, (n, k))


def dof(path, wrl=WRL):
    c = _run(["dof", wrl, path, "@out"], [0.0])
    r, k = int(c.scalar()), int(c.scalar())
    return c.take(r * k, (r, k))


# ---- libLBFGS through ctypes callbacks ------------------------------------------------------------------------------------
_EVAL = C.CFUNCTYPE(C.c_double, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int, C.c_double)
_PROG = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int)


def lbfgs_defaults():
    lib = C.CDLL(LBFGS_SO)
    out = (C.c_double * 14)()
    lib.ref_lbfgs_defaults(out)
    keys = ("m", "epsilon", "past", "delta", "max_iterations", "linesearch", "max_linesearch", "min_step", "max_step", "ftol", "wolfe", "gtol", "xtol", "sizeof_float")
    return dict(zip(keys, list(out)))


def lbfgs(x0, evaluate, epsilon=1e-5, m=0, max_iterations=-1):
    """Run the reference's libLBFGS on `evaluate(x) -> (f, g)`.  Returns dict(ret, x, fx, evaluations, iterations = per-iteration
    records (k, ls, fx, xnorm, gnorm, step, x))."""
    lib = C.CDLL(LBFGS_SO)
    lib.ref_lbfgs.restype = C.c_int
    lib.ref_lbfgs.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), _EVAL, _PROG, C.c_void_p, C.c_double, C.c_int, C.c_int]
    n = len(x0)
    x = (C.c_double * n)(*[float(v) for v in x0])
    nev = [0]
    its = []

    def ev(_inst, xp, gp, nn, _step):
        xv = np.ctypeslib.as_array(xp, shape=(nn,)).copy()
        f, g = evaluate(xv)
        nev[0] += 1
        for i in range(nn):
            gp[i] = float(g[i])
        return float(f)

    def pr(_inst, xp, gp, fx, xnorm, gnorm, step, nn, k, ls):
        its.append(dict(k=int(k), ls=int(ls), fx=float(fx), xnorm=float(xnorm), gnorm=float(gnorm), step=float(step),
                        x=np.ctypeslib.as_array(xp, shape=(nn,)).copy()))
        return 0
    fx = C.c_double(0.0)
    cb_ev, cb_pr = _EVAL(ev), _PROG(pr)
    ret = lib.ref_lbfgs(n, x, C.byref(fx), cb_ev, cb_pr, None, float(epsilon), int(m), int(max_iterations))
    return dict(ret=int(ret), x=np.array(list(x)), fx=float(fx.value), evaluations=nev[0], iterations=its)
