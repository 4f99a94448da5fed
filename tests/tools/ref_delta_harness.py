"""Run the REFERENCE'S OWN full-body mapping code (gym_trackSRB/taesooLib/DeltaExtractor.py: _extractSingleFrameFullbody :577-643,
unpack_obs_extra :1037-1043, pack_obs_extra :1015-1035, removeCOMsliding_online :421-512, removeFootSliding_online / _stair + avoidHull
:650-897) here, so that oracle/fullbody_map.py, the CoM-sliding and the foot-sliding restatements are compared with the reference code instead
of being trusted (rows f-2a ... f-2d).

DeltaExtractor.py is Python over taesooLib's libmainlib (`m = lua.taesooLib()`) and `lua.*` / `RE.*` helpers; libmainlib cannot be built in
this image.  The file is imported as it stands from /root/reference with
  m.vector3 / quater / transf / vectorn / matrixn / intvectorn   NumPy-backed classes offering the accessors these four functions use, each a
        restatement of the BaseLib source it names (vector3.cpp, quater.cpp / quater.h, transf.h / transf.cpp, vectorn.cpp:1107-1120); the
        algebra under them (quaternion product, rotation of a vector, transf product, exp(se3), rotation matrix -> quaternion) is
        oracle/bone_fk.py / oracle/fullbody_map.py, which tests/test_refpins.py pins to the COMPILED BaseLib at round-off;
  vectorn.to_se3, .slice, .range, `|`     binding-level helpers that are not in the tree: (w, v) order, copy, view, concatenation (assumed);
  the SRB tree (setPoseDOF / setDQ / calcMomentumCOM) and the full-body loader (setPoseDOF / calcCOM)   answered by the COMPILED reference
        (oracle/_ref/taesoo_ref fk: IK_sdls::LoaderToTree on a one-body WRL written here from the mass / inertia the test chooses, and on
        hanyang_lowdof_T_sh.wrl) -- SRB4.xml, where the reference takes the body from, is not in the tree;
  lua.new('QuadraticFunctionHardCon')     the Lua wrapper (module.lua:2090-2199) over OperatorStitch.cpp:468-620 restated: squared terms
        and hard constraints collected as written, the KKT system built as buildSystem builds it, solved by numpy (the reference: LU);
  lua.F('fc.getContactPositions_v2' / '_stair' / 'math.projectTrajectoryToItsHull')   see _lua_F: the first two are NOT in the reference tree
        (oracle stand-ins W5 / W6), the third is pure Lua restated over the pinned GrahamScan; what IS pinned for row f-2d is everything the
        Python function does with them -- the four x three quadratic programmes, their hard constraints, the ground clamp, avoidHull;
  RE.toVector3 / toQuater / ZtoY          the definitions of work/rendermodule.py:505-675 restated (three lines each).
Test infrastructure only."""
import importlib.util
import math
import os
import sys
import tempfile
import types

import numpy as np

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)
from oracle import bone_fk as fk                     # noqa: E402
from oracle import fullbody_map as fm                # noqa: E402
import ref_harness as rh                             # noqa: E402


def available():
    return os.path.exists(os.path.join(REF, "gym_trackSRB", "taesooLib", "DeltaExtractor.py")) and rh.available()


# ------------------------------------------------------------------------------------------------ libmainlib value types
class vector3:
    def __init__(self, x=0.0, y=0.0, z=0.0):
        self.x, self.y, self.z = float(x), float(y), float(z)

    def arr(self):
        return np.array([self.x, self.y, self.z])

    @staticmethod
    def of(a):
        return vector3(a[0], a[1], a[2])

    def __add__(self, o):
        return vector3(self.x + o.x, self.y + o.y, self.z + o.z)

    def __sub__(self, o):
        return vector3(self.x - o.x, self.y - o.y, self.z - o.z)

    def __mul__(self, s):
        return vector3(self.x * s, self.y * s, self.z * s)

    def cross(self, o):                                   # vector3::cross(b): this x b
        return vector3(self.y * o.z - self.z * o.y, self.z * o.x - self.x * o.z, self.x * o.y - self.y * o.x)

    def length(self):
        return math.sqrt(self.x * self.x + self.y * self.y + self.z * self.z)

    def normalize(self):                                  # vector3.cpp:271-279 (a zero vector stays zero)
        l = self.length()
        inv = 0.0 if l == 0 else 1 / l
        self.x *= inv; self.y *= inv; self.z *= inv

    def copy(self):
        return vector3(self.x, self.y, self.z)

    def normalized(self):
        v = self.copy(); v.normalize()
        return v

    def assign(self, o):
        self.x, self.y, self.z = o.x, o.y, o.z

    def __call__(self, i):
        return (self.x, self.y, self.z)[i]

    def ZtoY(self):                                       # rendermodule.py:633
        return vector3(self.y, self.z, self.x)


class quater:
    def __init__(self, *a):
        if len(a) == 4:
            self.w, self.x, self.y, self.z = (float(v) for v in a)
        elif len(a) == 2:                                 # quater(angle, axis) -> setRotation(axis, angle), quater.cpp:327-340
            ang, ax = a
            s = math.sin(0.5 * ang)
            self.w, self.x, self.y, self.z = math.cos(0.5 * ang), s * ax.x, s * ax.y, s * ax.z
        else:
            self.w, self.x, self.y, self.z = 1.0, 0.0, 0.0, 0.0

    def arr(self):
        return np.array([self.w, self.x, self.y, self.z])

    @staticmethod
    def of(a):
        return quater(a[0], a[1], a[2], a[3])

    def __mul__(self, o):
        if isinstance(o, quater):
            return quater.of(fk.qmult(self.arr(), o.arr()))
        return vector3.of(fk.vrotate(self.arr(), o.arr()))

    def ZtoY(self):                                       # rendermodule.py:631
        return quater(self.w, self.y, self.z, self.x)


class transf:
    def __init__(self, *a):
        self.rotation, self.translation = quater(), vector3()
        if len(a) == 2:
            self.rotation, self.translation = a
        elif len(a) == 1 and isinstance(a[0], vector3):
            self.translation = a[0]
        elif len(a) == 1:
            self.rotation = a[0]

    def __mul__(self, o):                                 # transf::mult (transf.cpp:98-103)
        return transf(self.rotation * o.rotation, (self.rotation * o.translation) + self.translation)

    def toGlobalPos(self, v):                             # transf.cpp:178-183
        return (self.rotation * v) + self.translation

    def toVector(self):
        v = vectorn(7)
        v.setTransf(0, self)
        return v

    def ZtoY(self):                                       # rendermodule.py:635
        return transf(self.rotation.ZtoY(), self.translation.ZtoY())


class se3:
    def __init__(self, a):
        self.a = np.asarray(a, dtype=float).copy()

    def exp(self):
        R, p = fm.se3_exp(self.a)
        return transf(quater.of(fm.mat_to_quat(R)), vector3.of(p))


class vectorn:
    def __init__(self, n=0, _a=None):
        self.a = np.zeros(int(n)) if _a is None else _a

    def size(self):
        return self.a.shape[0]

    def ref(self):
        return self.a

    def set(self, i, v):
        self.a[i] = v

    def __call__(self, i):
        return float(self.a[i])

    def __getitem__(self, i):
        return float(self.a[i])

    def __setitem__(self, i, v):
        self.a[i] = v

    def setAllValue(self, v):
        self.a[:] = v

    def rmult(self, s):
        self.a *= s

    def sum(self):
        return float(self.a.sum())

    def range(self, s, e):
        return vectorn(_a=self.a[s:e])

    def slice(self, s, e):
        return vectorn(_a=self.a[_span(s, e, self.a.shape[0])].copy())

    def __or__(self, o):
        return vectorn(_a=np.concatenate([self.a, o.a]))

    def toVector3(self, i=0):
        return vector3.of(self.a[i:i + 3])

    def toQuater(self, i=0):
        return quater.of(self.a[i:i + 4])

    def toTransf(self, i=0):                              # 7-dim format (p, (w, x, y, z)), vectorn.h:329
        return transf(self.toQuater(i + 3), self.toVector3(i))

    def setVec3(self, i, v):
        self.a[i:i + 3] = v.arr()

    def setTransf(self, i, t):
        self.a[i:i + 3] = t.translation.arr(); self.a[i + 3:i + 7] = t.rotation.arr()

    def to_se3(self):
        return se3(self.a[0:6])


class matrixn:
    def __init__(self, r=0, c=0, _a=None):
        self.a = np.zeros((int(r), int(c))) if _a is None else _a

    def rows(self):
        return self.a.shape[0]

    def cols(self):
        return self.a.shape[1]

    def row(self, i):
        return vectorn(_a=self.a[i])

    def column(self, j):
        return vectorn(_a=self.a[:, j])

    def __getitem__(self, i):
        return vectorn(_a=self.a[i])

    def __call__(self, i, j):
        return float(self.a[i, j])

    def matView(self):
        return self

    def ref(self):
        return self.a

    def set(self, i, j, v):
        self.a[i, j] = v

    def sub(self, r0, r1, c0, c1):                        # matrixn:sub: negative = from the end, end 0 = to the end (view)
        return matrixn(_a=self.a[_span(r0, r1, self.a.shape[0]), _span(c0, c1, self.a.shape[1])])

    def assign(self, o):
        self.a[:, :] = o.a

    def __or__(self, o):
        return matrixn(_a=np.concatenate([self.a, o.a], axis=1))

    def __sub__(self, o):
        return matrixn(_a=self.a - o.a)

    def derivative_forward(self, rate):                   # forward difference, last row repeated (binding-level helper, assumed)
        d = np.zeros_like(self.a)
        d[:-1] = (self.a[1:] - self.a[:-1]) * rate
        d[-1] = d[-2]
        return matrixn(_a=d)


def _span(s, e, n):
    if s < 0:
        s += n
    if e <= 0:
        e += n
    return slice(s, e)


class vector3N:
    """rows of vector3 sharing one [n, 3] array; v(i) is a reference to row i."""

    def __init__(self, n=0, _a=None):
        self.a = np.zeros((int(n), 3)) if _a is None else _a

    def size(self):
        return self.a.shape[0]

    rows = size

    class _Row(vector3):
        """v(i): a vector3 that lives in row i (assignments write through)"""

        def __init__(self, a):
            self._a = a
        x = property(lambda s: float(s._a[0]), lambda s, v: s._a.__setitem__(0, v))
        y = property(lambda s: float(s._a[1]), lambda s, v: s._a.__setitem__(1, v))
        z = property(lambda s: float(s._a[2]), lambda s, v: s._a.__setitem__(2, v))

    def __call__(self, i):
        return vector3N._Row(self.a[i])

    def slice(self, s, e):
        return vector3N(_a=self.a[_span(s, e, self.a.shape[0])])

    def matView(self):
        return matrixn(_a=self.a)

    def __add__(self, o):
        return vector3N(_a=self.a + o.a)

    def __sub__(self, o):
        return vector3N(_a=self.a - o.a)

    def __mul__(self, s):
        return vector3N(_a=self.a * s)


class boolN:
    def __init__(self, a):
        self.a = np.asarray(a, dtype=bool)

    def __call__(self, i):
        return bool(self.a[i])

    def slice(self, s, e):
        return boolN(self.a[_span(s, e, self.a.shape[0])])

    def __or__(self, o):
        return boolN(self.a | o.a)

    def count(self):
        return int(self.a.sum())

    def size(self):
        return self.a.shape[0]


class intIntervals:
    def __init__(self, runs):
        self.runs = list(runs)

    def size(self):
        return len(self.runs)

    def startI(self, i):
        return self.runs[i][0]

    def endI(self, i):
        return self.runs[i][1]


class intvectorn:
    def __init__(self, n=0):
        self.a = np.zeros(int(n), dtype=int)

    def colon(self, start, end, step=1):
        self.a = np.arange(start, end, step, dtype=int)

    def size(self):
        return self.a.shape[0]


class QuadraticFunctionHardCon:
    """module.lua:2090-2199 over OperatorStitch.cpp:176-260 (squared terms), :468-500 (constraints), :578-620 (buildSystem)."""

    def __init__(self, nvar, ncon):
        self.numVar, self.numCon = int(nvar), int(ncon)
        self.sq, self.cons = [], []

    @staticmethod
    def _pairs(p):
        n = len(p) // 2
        return np.array([int(p[2 * i + 1]) for i in range(n)], dtype=int), np.array([float(p[2 * i]) for i in range(n)] + [float(p[2 * n])])

    def addWeighted(self, w, *p):
        idx, coef = self._pairs(p)
        self.sq.append((idx, coef * math.sqrt(w)))

    def add(self, *p):
        self.sq.append(self._pairs(p))

    def addSquared(self, index, coef):
        self.sq.append((index.a.copy(), coef.a.copy()))

    def con(self, *p):
        self.cons.append(self._pairs(p))

    def addCon(self, index, coef):
        self.cons.append((index.a.copy(), coef.a.copy()))

    def solve(self):
        n, nc = self.numVar, self.numCon
        assert len(self.cons) == nc
        Aug = np.zeros((n + nc, n + nc)); d = np.zeros(n + nc)
        for idx, c in self.sq:                              # d/dx of (sum c_i x_i + c_n)^2 = 0  ->  H x = d0
            k = len(idx)
            for i in range(k):
                for j in range(k):
                    Aug[idx[i], idx[j]] += 2.0 * c[i] * c[j]
                d[idx[i]] -= 2.0 * c[k] * c[i]
        for r, (idx, c) in enumerate(self.cons):
            for i in range(len(idx)):
                Aug[n + r, idx[i]] = c[i]
            d[n + r] = c[len(idx)] * -1.0
        Aug[:n, n:] = Aug[n:, :n].T
        return vectorn(_a=np.linalg.solve(Aug, d)[:n].copy())


# ------------------------------------------------------------------------------------------------ loaders / trees answered by the compiled reference
def _one_body_wrl(path, mass, inertia, com):
    I = np.asarray(inertia, dtype=float).reshape(3, 3)
    with open(path, "w") as f:
        f.write('DEF SampleRobot Humanoid {\n name "srb"\n humanoidBody [\n  DEF root Joint\n  {\n    jointType "free"\n    jointAxis "Z"\n'
                '    translation 0.000000 0.000000 0.000000\n    children [\n      Segment\n      {\n'
                f'        centerOfMass {com[0]:.17g} {com[1]:.17g} {com[2]:.17g}\n        mass {mass:.17g}\n'
                '        momentsOfInertia [' + " ".join(f"{v:.17g}" for v in I.ravel()) + ']\n        material "shiny"\n'
                '        children   Transform { rotation 0.000000 0.000000 0.000000 0.000000 translation 0.000000 0.000000 0.000000  children Shape {geometry Box { size 0.300000 0.130000 0.150000 }}}\n'
                '      }\n    ]\n  }\n ]\n}\n')


class _Momentum:
    def __init__(self, h):
        self.h = h

    def M(self):
        return vector3.of(self.h[0:3])

    def F(self):
        return vector3.of(self.h[3:6])


class SRBTree:
    """srb_tree / srb_loader of walk_SRBTrack2025.py:243-251 for one free body of the given mass / inertia / local CoM."""

    def __init__(self, mass, inertia, com):
        self._td = tempfile.TemporaryDirectory()
        self.wrl = os.path.join(self._td.name, "srb_one_body.wrl")
        _one_body_wrl(self.wrl, mass, inertia, com)
        self.dofInfo = None
        self.pose = None; self.dq = None

    def setPoseDOF(self, dofInfo, pose):
        self.pose = pose.a.copy()

    def setDQ(self, dq):
        self.dq = dq.a.copy()

    def calcMomentumCOM(self, loader):
        r = rh.fk(self.pose[None], self.dq[None], wrl=self.wrl)
        return _Momentum(r["mom"][0])


class FullBodyLoader:
    """mLoader_full: setPoseDOF + calcCOM by the compiled reference on hanyang_lowdof_T_sh.wrl (all poses of a window in one call, cached)."""

    def __init__(self, poses):
        poses = np.asarray(poses, dtype=float)
        self._com = {tuple(np.round(p, 12)): c for p, c in zip(poses, rh.fk(poses, np.zeros((poses.shape[0], 59)))["com"])}
        self._cur = None

    def setPoseDOF(self, pose):
        self._cur = tuple(np.round(pose.a, 12))

    def calcCOM(self):
        return vector3.of(self._com[self._cur])


def _lua_F(name, *a):
    """The three Lua functions removeFootSliding_online(_stair) calls.  fc.getContactPositions_v2 / _stair are NOT in the reference tree
    (stand-ins W5 / W6 of oracle/foot_sliding.py, restated from their call sites); math.projectTrajectoryToItsHull is pure Lua
    (gym_cdm2/ConvexHull2D.lua:362-455, restated in oracle/foot_sliding.py over the GrahamScan that tests/test_refpins.py pins to graham.cpp)
    and changes the heights of its argument in place."""
    from oracle import foot_sliding as fs
    if name == "fc.getContactPositions_v2":
        foot_len, con_toe, con_heel, toe, heel = a
        pos, dr, ii = fs.get_contact_positions_v2(foot_len, con_toe.a, con_heel.a, toe.a, heel.a)
        return vector3N(_a=np.array(pos, dtype=float).reshape(-1, 3).copy()), vector3N(_a=np.array(dr, dtype=float).reshape(-1, 3).copy()), intIntervals(ii)
    if name == "fc.getContactPositions_stair":
        con_toe, con_heel, center = a
        pos, ii = fs.get_contact_positions_stair(con_toe.a, con_heel.a, center.a)
        return vector3N(_a=np.array(pos, dtype=float).reshape(-1, 3).copy()), intIntervals(ii)
    if name == "math.projectTrajectoryToItsHull":
        tpos, = a
        changed, y = fs.project_trajectory_to_its_hull(tpos.a[:, 1].copy())
        tpos.a[:, 1] = y
        return changed
    raise KeyError(name)


# ------------------------------------------------------------------------------------------------ import of the reference file
def load():
    """-> the reference's DeltaExtractor module."""
    lua = types.ModuleType("work.luamodule")
    mm = types.SimpleNamespace(vector3=vector3, quater=quater, transf=transf, vectorn=vectorn, matrixn=matrixn, intvectorn=intvectorn, vector3N=vector3N)
    lua.taesooLib = lambda: mm
    lua.zeros = lambda n, c=None: vectorn(n) if c is None else matrixn(n, c)
    lua.vec = lambda v: vectorn(_a=np.asarray(v, dtype=float).copy())
    lua.new = lambda name, *a: {"QuadraticFunctionHardCon": QuadraticFunctionHardCon}[name](*a)
    lua.ivec = lambda a: boolN(a)
    lua.F = _lua_F
    RE = types.ModuleType("work.rendermodule")
    RE.toVector3 = lambda v: vector3(v[0], v[1], v[2])                                         # rendermodule.py:550-563

    def toQuater(v):                                                                           # rendermodule.py:505-533
        v = np.asarray(v, dtype=float)
        if v.size == 9:
            return quater.of(fm.mat_to_quat(v.reshape(3, 3)))
        return quater(v[0], v[1], v[2], v[3])
    RE.toQuater = toQuater
    work = types.ModuleType("work"); work.__path__ = []
    work.luamodule = lua; work.rendermodule = RE
    easydict = types.ModuleType("easydict")
    easydict.EasyDict = lambda **k: types.SimpleNamespace(**k)
    names = ("work", "work.luamodule", "work.rendermodule", "easydict", "mmsto_config")
    saved = {k: sys.modules.get(k) for k in names}
    d = os.path.join(REF, "gym_trackSRB", "taesooLib")
    try:
        sys.modules.update({"work": work, "work.luamodule": lua, "work.rendermodule": RE, "easydict": easydict})
        sys.modules.pop("mmsto_config", None)
        sys.path.insert(0, d)
        spec = importlib.util.spec_from_file_location("_ref_DeltaExtractor", os.path.join(d, "DeltaExtractor.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    finally:
        sys.path.remove(d)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def extract(mod, extra, row, srb_mass, srb_inertia, srb_local_com):
    """unpack_obs_extra + _extractSingleFrameFullbody of the reference on one (obs_extra[29], decoder row[214]) pair -> the seven outputs
    as arrays, keyed like oracle.fullbody_map.extract_single_frame_fullbody."""
    tree = SRBTree(srb_mass, srb_inertia, srb_local_com)
    pose_tl, dq, foot = mod.unpack_obs_extra(vectorn(_a=np.asarray(extra, dtype=float).copy()))
    human, dxddx, feet, com, mom, con_bin, con = mod._extractSingleFrameFullbody(tree, tree, pose_tl, dq, foot, vectorn(_a=np.asarray(row, dtype=float).copy()), None)
    return dict(humanPose=human.a.copy(), dx_ddx=dxddx.a.copy(), feetpos=feet.a.copy(), com=com.arr(), momentum=mom.a.copy(), con_bin=con_bin.a.copy(), con=con.a.copy())


def remove_com_sliding(mod, com_srb, desired_com, poses, fix_y_also=True, ground_h=None):
    """removeCOMsliding_online of the reference on one window (poses [n, 60] are modified) -> new poses.  ground_h[n]: what the caller's
    projectToGround(cc) would set cc.y to for the i-th current CoM (the reference calls it once per frame, in frame order)."""
    out = matrixn(_a=np.asarray(poses, dtype=float).copy())
    project = None
    if ground_h is not None:
        it = iter(np.asarray(ground_h, dtype=float))

        def project(cc):
            cc.y = float(next(it))
    mod.removeCOMsliding_online(FullBodyLoader(poses), matrixn(_a=np.asarray(com_srb, dtype=float).copy()), matrixn(_a=np.asarray(desired_com, dtype=float).copy()), out, fix_y_also, project)
    return out.a


def remove_foot_sliding(mod, foot_len, con_all, initial_feet, marker_traj, marker, foot_center=None, ground=None, desired_vel=None):
    """removeFootSliding_online (foot_center None) / removeFootSliding_online_stair of the reference -> optMarkerTraj [nvar, 12].
    ground(p[3]) -> height under p (what projectToGround sets p.y to; the reference's call sites always pass one)."""
    def project(v):
        v.y = float(ground(np.array([v.x, v.y, v.z])))
    mt = [vector3N(_a=np.asarray(t, dtype=float).copy()) for t in marker_traj]
    dv = None if desired_vel is None else matrixn(_a=np.asarray(desired_vel, dtype=float).copy())
    args = [foot_len, matrixn(_a=np.asarray(con_all, dtype=float).copy()), vectorn(_a=np.asarray(initial_feet, dtype=float).copy()), mt,
            matrixn(_a=np.asarray(marker, dtype=float).copy())]
    if foot_center is None:
        out = mod.removeFootSliding_online(*args, dv, project)
    else:
        out = mod.removeFootSliding_online_stair(*args, [vector3N(_a=np.asarray(c, dtype=float).copy()) for c in foot_center], dv, project)
    return out.a
