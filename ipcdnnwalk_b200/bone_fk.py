"""Batched look-alikes of the taesooLib python-bound kinematics classes over the CUDA FK kernel.

Reference interface mirrored (work/taesooLib/MainLib/python/MainlibPython.hpp):
  :3171-3197  BoneForwardKinematics(loader): setPoseDOF(vectorn), forwardKinematics(), globalFrame(i) -> transf
  :3496-3550  LoaderToTree(loader, useEulerRoot, useFixedRootPos): setPoseDOF(dofInfo, pose), setDQ(dq),
              calcCOM(loader) -> vector3, calcMomentumCOM(loader) -> dse3, globalFrame(i)
as used by gym_trackSRB/taesooLib/DeltaExtractor.py:577-643.  Here every call takes a batch: `pose` is
[N, 7+#angles] (root xyz, root quaternion w x y z, Euler angles in tree order), `dq` is [N, 6+#angles]
(root angular velocity world, root linear velocity world, joint rates).  Bone index 0 is the root (taesooLib
treeIndex 1).  torch is used for device memory only."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check

AX = {"X": 0, "Y": 1, "Z": 2}


def _ptr(t):
    if t is None:
        return None
    if isinstance(t, np.ndarray):
        return t.ctypes.data_as(C.c_void_p)
    return C.c_void_p(t.data_ptr())


class Skeleton:
    """A VRMLloader resident on one GPU (tables from ipcdnnwalk_b200.vrml_skeleton)."""

    def __init__(self, skel, device=0):
        import torch
        self.torch = torch
        if not torch.cuda.is_available():
            raise RuntimeError("ipcdnnwalk_b200.bone_fk needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.skel = skel
        self.device = torch.device("cuda", device)
        nb = len(skel["parent"])
        parent = np.ascontiguousarray(skel["parent"], dtype=np.int32)
        offset = np.ascontiguousarray(skel["offset"], dtype=np.float64)
        nchan = np.ascontiguousarray([len(c) for c in skel["channels"]], dtype=np.int32)
        axes = np.zeros((nb, 3), dtype=np.int32)
        for b, ch in enumerate(skel["channels"]):
            for k, c in enumerate(ch):
                axes[b, k] = AX[c]
        mass = np.ascontiguousarray(skel["mass"], dtype=np.float64)
        com = np.ascontiguousarray(skel["com"], dtype=np.float64)
        inertia = np.ascontiguousarray(skel["inertia"], dtype=np.float64)
        self._h = C.c_void_p()
        check(self.lib.fk_create(int(device), nb, _ptr(parent), _ptr(offset), _ptr(nchan), _ptr(axes), _ptr(mass), _ptr(com), _ptr(inertia),
                                 C.byref(self._h)))
        d = [C.c_int32() for _ in range(3)]
        check(self.lib.fk_dims(self._h, *[C.byref(x) for x in d]))
        self.num_bones, self.n_pose, self.n_dq = (x.value for x in d)

    def numBone(self):
        return self.num_bones

    def set_generic(self, force=True):
        """Force the structure-agnostic kernel even when the skeleton-specialised one applies (A/B runs, tests)."""
        check(self.lib.fk_set_generic(self._h, int(bool(force))))

    def set_tma(self, use=True):
        """Specialised path: -1 / True = automatic choice, 0 / False = rows accessed directly, 1 = rolled loop on TMA tiles,
        3 = unrolled sweep on TMA tiles, 4 = chain-wise unrolled bodies on TMA tiles (A/B runs, tests)."""
        if use is True:
            use = -1
        check(self.lib.fk_set_tma(self._h, int(use)))

    def is_specialised(self):
        return bool(self.lib.fk_is_specialised(self._h))

    def forward(self, pose, dq=None, want_xform=True, want_com=True, want_momentum=True, out=None):
        t = self.torch
        assert pose.is_cuda and pose.dim() == 2 and pose.shape[1] == self.n_pose and pose.is_contiguous()
        n = pose.shape[0]
        dt = pose.dtype
        assert dt in (t.float32, t.float64)
        if dq is not None:
            assert dq.shape == (n, self.n_dq) and dq.dtype == dt and dq.is_contiguous()
        out = out or {}
        xf = out.get("xform") if want_xform else None
        if want_xform and xf is None:
            xf = t.empty(n, self.num_bones, 7, dtype=dt, device=pose.device)
        com = out.get("com") if want_com else None
        if want_com and com is None:
            com = t.empty(n, 3, dtype=dt, device=pose.device)
        mom = out.get("momentum") if (want_momentum and dq is not None) else None
        if want_momentum and dq is not None and mom is None:
            mom = t.empty(n, 6, dtype=dt, device=pose.device)
        fn = self.lib.fk_forward_f32 if dt == t.float32 else self.lib.fk_forward_f64
        stream = C.c_void_p(t.cuda.current_stream(pose.device).cuda_stream)
        check(fn(self._h, n, _ptr(pose), _ptr(dq), _ptr(xf), _ptr(com), _ptr(mom), stream))
        return dict(xform=xf, com=com, momentum=mom)

    def close(self):
        if self._h:
            self.lib.fk_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class BoneForwardKinematics:
    def __init__(self, skeleton):
        self.skeleton = skeleton
        self._pose = None
        self._xform = None

    def setPoseDOF(self, pose):
        self._pose = pose
        self.forwardKinematics()

    def forwardKinematics(self):
        self._xform = self.skeleton.forward(self._pose, None, want_xform=True, want_com=False, want_momentum=False)["xform"]

    def getPoseDOF(self):
        return self._pose

    def globalFrame(self, bone):
        """[N,7] = quaternion (w,x,y,z) + translation of bone `bone` (0-based, or a bone name)."""
        if isinstance(bone, str):
            bone = self.skeleton.skel["names"].index(bone)
        return self._xform[:, bone]

    def global_frames(self):
        return self._xform


class LoaderToTree:
    def __init__(self, skeleton, useEulerRoot=False, useFixedRootPos=False):
        if useEulerRoot or useFixedRootPos:
            raise NotImplementedError("only the quaternion-root tree used by DeltaExtractor.py:577-643 is implemented")
        self.skeleton = skeleton
        self._pose = self._dq = None
        self._res = None

    def setPoseDOF(self, pose):
        self._pose = pose
        self._res = None

    def setDQ(self, dq):
        self._dq = dq
        self._res = None

    def _run(self):
        if self._res is None:
            self._res = self.skeleton.forward(self._pose, self._dq)
        return self._res

    def globalFrame(self, bone):
        return self._run()["xform"][:, bone]

    def calcCOM(self):
        return self._run()["com"]

    def calcMomentumCOM(self):
        """[N,6] dse3: angular momentum about the CoM (M), linear momentum (F), world frame."""
        return self._run()["momentum"]
