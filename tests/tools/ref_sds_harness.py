"""Run the REFERENCE'S OWN swing-foot filter (work/controlmodule.py: NonlinearController :911-998, SDS :999-1086) here, so that
oracle/sds_filter.py is compared with the reference code instead of being trusted.

controlmodule.py is Python on top of taesooLib's libmainlib (`m = lua.taesooLib()`: matrixn / vectorn) and a few `lua.*` helpers; neither
can be built in this image.  This module answers exactly the accessors those two classes use with NumPy-backed objects of the same
meaning (views share memory, `mult` / `radd` / `assign` write in place) and imports the reference file as it stands from
/root/reference.  Stand-ins that carry semantics of their own, and where those are pinned:
  matrixn.sampleRow        -> oracle.tl_math.sample_row, bit exact against the compiled matrixn.cpp (tests/test_refpins.py)
  lua.F('sop.piecewiseLinearMap', ...) -> oracle.sds_filter.piecewise_linear_map (module.lua:5203-5233 restated; pure Lua, not runnable here)
Everything else -- the controller equations, the order of the semi-implicit update, the constant `G`, the gain table, the fall-back from
`ct.lqr` to `useKpreset` -- is the reference's code.  Test infrastructure only (tests/test_refenv_pins.py, tests/tools/make_golden_refenv.py)."""
import importlib.util
import os
import sys
import types

import numpy as np

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def available():
    return os.path.exists(os.path.join(REF, "work", "controlmodule.py"))


class vectorn:
    def __init__(self, n=0, _a=None):
        self.a = np.zeros(int(n)) if _a is None else _a

    def size(self):
        return self.a.shape[0]

    def set(self, i, v):
        self.a[i] = v

    def __call__(self, i):
        return float(self.a[i])

    def assign(self, vals):
        vals = np.asarray(vals.a if isinstance(vals, vectorn) else vals, dtype=float)
        if self.a.shape[0] != vals.shape[0]:
            assert self.a.base is None, "cannot resize a view"
            self.a = np.zeros(vals.shape[0])
        self.a[:] = vals

    def setAllValue(self, v):
        self.a[:] = v

    def range(self, s, e):
        return vectorn(_a=self.a[s:e])

    def column(self):
        return matrixn(_a=self.a.reshape(-1, 1))

    def ref(self):
        return self.a


class matrixn:
    def __init__(self, r=0, c=0, _a=None):
        self.a = np.zeros((int(r), int(c))) if _a is None else _a

    def rows(self):
        return self.a.shape[0]

    def cols(self):
        return self.a.shape[1]

    def set(self, i, j, v):
        self.a[i, j] = v

    def __call__(self, i, j):
        return float(self.a[i, j])

    def ref(self):
        return self.a

    def setAllValue(self, v):
        self.a[:, :] = v

    def range(self, r0, r1, c0, c1):
        return matrixn(_a=self.a[r0:r1, c0:c1])

    def diag(self):
        return vectorn(_a=np.einsum("ii->i", self.a))          # writable view of the diagonal

    def row(self, i):
        return vectorn(_a=self.a[i])

    def column(self, j):
        return vectorn(_a=self.a[:, j])

    def mult(self, A, B):
        self.a[:, :] = A.a @ B.a

    def radd(self, x):
        self.a += x.a if isinstance(x, matrixn) else x

    def assign(self, x):
        self.a[:, :] = x.a if isinstance(x, matrixn) else np.asarray(x, dtype=float).reshape(self.a.shape)

    def sampleRow(self, t, out):
        from oracle.tl_math import sample_row
        out.a[:] = sample_row(self.a, float(t))

    def __mul__(self, x):
        return matrixn(_a=self.a @ x.a) if isinstance(x, matrixn) else matrixn(_a=self.a * x)

    def __sub__(self, x):
        return matrixn(_a=self.a - x.a)

    def __add__(self, x):
        return matrixn(_a=self.a + x.a)


def _fake_lua():
    from oracle.sds_filter import piecewise_linear_map
    lua = types.ModuleType("work.luamodule")
    mm = types.SimpleNamespace(matrixn=matrixn, vectorn=vectorn)
    lua.taesooLib = lambda: mm
    lua.zeros = lambda r, c=None: matrixn(r, c) if c is not None else vectorn(r)
    lua.vec = lambda v: vectorn(_a=np.asarray(v, dtype=float).copy())
    lua.eye = lambda n: matrixn(_a=np.eye(n))

    def mat(*a):
        if len(a) == 1:
            return matrixn(_a=np.array(a[0], dtype=float).reshape(np.asarray(a[0]).shape if np.asarray(a[0]).ndim == 2 else (1, -1)))
        r, c = int(a[0]), int(a[1])
        vals = np.asarray(a[2:], dtype=float)
        return matrixn(_a=(np.full((r, c), vals[0]) if vals.size == 1 else vals.reshape(r, c)).copy())
    lua.mat = mat

    def F(name, *args):
        assert name == "sop.piecewiseLinearMap", name
        return piecewise_linear_map(*args)
    lua.F = F
    return lua


def load():
    """-> the reference's controlmodule (imported from /root/reference/work/controlmodule.py with the stand-ins above)."""
    saved = {k: sys.modules.get(k) for k in ("luamodule", "rendermodule", "work", "work.luamodule", "work.rendermodule")}
    try:
        lua = _fake_lua()
        sys.modules["luamodule"] = lua
        sys.modules["rendermodule"] = types.ModuleType("rendermodule")
        spec = importlib.util.spec_from_file_location("_ref_controlmodule", os.path.join(REF, "work", "controlmodule.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def drive(SDS, plan, steps_per_update=4):
    """The call pattern of SRBEnv.step (SRBEnv5_mocap_list_window.py:1228-1299) on one filter: construct, set x / xd, then per RL step
    updateCoef(B, k, Q), new target, `steps_per_update` singleSteps; plan rows = (Q, target, reset_velocity flag).  Returns x after
    every singleStep [n, 2] and the gain after every updateCoef [n_updates, 2]."""
    get = (lambda mth, i: mth(i, 0)) if not isinstance(getattr(SDS, "__name__", ""), int) else None
    s = SDS(60.0, 0.001, 1, float(plan[0][0]), 0, 1 / 120.0, True)
    is_ref = hasattr(s.x, "set")
    setv = (lambda mth, i, v: mth.set(i, 0, v)) if is_ref else (lambda arr, i, v: arr.__setitem__(i, v))
    getv = (lambda mth, i: mth(i, 0)) if is_ref else (lambda arr, i: float(arr[i]))
    setv(s.x, 0, float(plan[0][1])); setv(s.xd, 0, float(plan[0][1]))
    xs, Ks = [], []
    for Q, target, zero_vel in plan:
        s.updateCoef(0.001, 1, float(Q))
        K = s.K
        Ks.append([K(0, 0), K(0, 1)] if is_ref else [float(K[0]), float(K[1])])
        setv(s.xd, 0, float(target))
        if zero_vel:
            setv(s.x, 1, 0.0); setv(s.xd, 1, 0.0)
        for _ in range(steps_per_update):
            s.singleStep()
            xs.append([getv(s.x, 0), getv(s.x, 1)])
    return np.array(xs), np.array(Ks)
