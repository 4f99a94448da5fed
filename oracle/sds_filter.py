"""ORACLE (test infrastructure only).  CPU restatement of the swing-foot LQR filter `SDS`.

Follows, line by line:
  work/controlmodule.py:911-998   NonlinearController (oneStep / oneStepRaw / calcControlForce)
  work/controlmodule.py:999-1086  SDS (useKpreset, updateCoef, singleStep)
  work/taesooLib/Resource/scripts/module.lua:5203-5233   sop.piecewiseLinearMap
  work/taesooLib/BaseLib/math/Operator.cpp:16-28         sop::interpolate / sop::map
  work/taesooLib/BaseLib/math/matrixn.cpp:781-804        matrixn::sampleRow (float32 rounding of t)

Only the code path the SRBTrack env takes is restated: usePreset=True so `self.ct is None`, hence
`updateCoef` always raises inside the try block and falls back to `useKpreset(Q)`.
PARITY PINNED against the reference's own class: tests/tools/ref_sds_harness.py imports work/controlmodule.py from /root/reference over
NumPy-backed matrixn / vectorn objects (libmainlib cannot be built here) and tests/test_refenv_pins.py compares gains (exact) and filter
states (1e-13) with it -- fixture tests/golden/refenv_sds.npz; `sample_row` is bit exact against the compiled matrixn.cpp
(tests/test_refpins.py).  Not pinned: `piecewise_linear_map` (pure Lua in the reference, restated).
"""
import math
import numpy as np

# work/controlmodule.py:1047-1054
K_PRESET = np.array([
    [262.685, 127.410],      # 1e5
    [942.739, 281.657],      # 1e6
    [3103.828, 553.238],     # 1e7
    [9941.174, 1033.866],    # 1e8
    [31563.832, 1887.117],   # 1e9
    [99941.017, 3403.601],   # 1e10
    [316168.772, 6099.859],  # 1e11
])


def sop_interpolate(t, s, e):
    return s * (1.0 - t) + e * t


def sop_map(t, mn, mx, v1, v2):
    t = (t - mn) / (mx - mn)
    return sop_interpolate(t, v1, v2)


def piecewise_linear_map(t_global, keytimes, values):
    """module.lua:5203-5233"""
    eps = 1e-10
    iseg = -1
    w = None
    for i in range(len(keytimes) - 1):
        t = keytimes[i + 1]
        d = t - keytimes[i]
        if t >= t_global - eps:
            iseg = i
            w = sop_map(t_global, t - d, t, 0, 1)
            break
    if iseg == -1:
        return values[-1]
    return sop_map(w, 0, 1, values[iseg], values[iseg + 1])


def sample_row(mat, critical_time):
    """matrixn::sampleRow -- `float t` is single precision in the C++ source."""
    a = int(math.floor(critical_time))
    t = float(np.float32(critical_time - float(np.float32(a))))
    if t < 0.005:
        return mat[a].copy()
    if t > 0.995:
        return mat[a + 1].copy()
    if a < 0:
        return sop_interpolate(t - 1.0, mat[a + 1], mat[a + 2])
    if a + 1 >= mat.shape[0]:
        return sop_interpolate(t + 1.0, mat[a - 1], mat[a])
    return sop_interpolate(t, mat[a], mat[a + 1])


def k_preset(q):
    """SDS.useKpreset (controlmodule.py:1041-1057)."""
    src = [math.pow(10, i) for i in range(5, 13)]
    tgt = [float(i) for i in range(0, 8)]
    index = piecewise_linear_map(q, src, tgt)
    return sample_row(K_PRESET, index)


class SDS:
    """1-dof spring-damper tracked by a preset-gain LQR (state x=[theta, dtheta], target xd)."""

    def __init__(self, mass, B, k, q, qd, dt, usePreset=True):
        assert usePreset
        self.M = mass
        self.b = B
        self.k = k
        self.dt = dt
        self.t = 0
        self.K = k_preset(q)
        self.x = np.zeros(2)
        self.xd = np.zeros(2)

    def updateCoef(self, B, k, Q):
        self.b = B
        self.k = k
        self.K = k_preset(Q)

    def singleStep(self):
        invM = 1 / self.M
        C = self.b * invM          # updateSDRE: C = b/M, G = k/M (constant, not G*theta: Q4), H = 1/M
        G = self.k * invM
        H = 1 * invM
        U = self.K[0] * self.xd[0] + self.K[1] * self.xd[1]                # U.assign(K*xd)
        cf = U - (self.K[0] * self.x[0] + self.K[1] * self.x[1])          # U - K*x
        Dddtheta = H * cf - G - C * self.x[1]
        ddtheta = 1.0 * Dddtheta                                          # invD = 1
        self.x[1] += ddtheta * self.dt
        self.x[0] += self.x[1] * self.dt
        self.t = self.t + self.dt
