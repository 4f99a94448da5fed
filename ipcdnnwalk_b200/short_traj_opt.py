"""Batched look-alike of gym_cdm2's `ShortTrajOpt` (shortTermOptimizer.lua) over the CUDA MMSTO kernel.

Reference interface mirrored (the calls walk_SRBTrack2025.py:293,370-376,529-587 makes on one optimiser):
  ShortTrajOpt(loader, motion)                     -> ShortTrajOpt(skel, nf, markers, device)
  trajOpt.set('weightMomentum', w) ...             -> set(name, value)
  trajOpt.get('dof_weights').range(s, s+3).setAllValue(0.1)   -> dof_weights[s:s+3] = 0.1; set_dof_weights()
  trajOpt:calculateOriginalEulerPoses(mot, {poseOnly=true})   -> calculateOriginalEulerPoses(pose)  (host side, numpy)
  trajOpt:solveEndEffectors(mot, markers, gpos_target, nil, nil, hsVelConInfo)   -> solveEndEffectors_euler(...)
Every array carries a leading batch dimension: one independent window per character.  The online configuration only
(onlineOpt = true, no collision half-spaces, weightEEvel = 0): anything else raises.  torch is used for device memory."""
import ctypes as C
import math

import numpy as np

from . import _lib

AX = {"X": 0, "Y": 1, "Z": 2}
ROOT_ROT_CHANNELS = "YZX"          # VRMLloader.cpp:16 DEFAULT_ROT_CHANNELS
INFO_KEYS = ("fx", "ret", "iterations", "evaluations")


class MmstoError(RuntimeError):
    pass


def _check(lib, rc):
    if rc != 0:
        raise MmstoError(f"libsrbtrack_b200 mmsto error {rc}: {lib.mmsto_last_error().decode()}")


def _ptr(t):
    if t is None:
        return None
    if isinstance(t, np.ndarray):
        return t.ctypes.data_as(C.c_void_p)
    return C.c_void_p(t.data_ptr())


def _quat_to_yzx(q):
    """Euler angles (ry, rz, rx) with R = Ry Rz Rx from quaternions [.., 4] (w x y z); rz in [-pi/2, pi/2]."""
    q = q / np.linalg.norm(q, axis=-1, keepdims=True)
    w, x, y, z = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    r00 = 1 - 2 * (y * y + z * z); r10 = 2 * (x * y + w * z); r20 = 2 * (x * z - w * y)
    r11 = 1 - 2 * (x * x + z * z); r12 = 2 * (y * z - w * x)
    rz = np.arcsin(np.clip(r10, -1.0, 1.0))
    ry = np.arctan2(-r20, r00)
    rx = np.arctan2(-r12, r11)
    return np.stack([ry, rz, rx], axis=-1)


def _yzx_to_quat(e):
    def ax(axis, a):
        q = np.zeros(a.shape + (4,)); q[..., 0] = np.cos(a / 2); q[..., 1 + axis] = np.sin(a / 2)
        return q

    def mul(a, b):
        w1, x1, y1, z1 = np.moveaxis(a, -1, 0); w2, x2, y2, z2 = np.moveaxis(b, -1, 0)
        return np.stack([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                         w1 * y2 + y1 * w2 + z1 * x2 - x1 * z2, w1 * z2 + z1 * w2 + x1 * y2 - y1 * x2], axis=-1)
    return mul(mul(ax(1, e[..., 0]), ax(2, e[..., 1])), ax(0, e[..., 2]))


def align_euler_angles(a):
    """vectorn::alignEulerAngles along the last axis: remove 2 pi jumps between consecutive frames."""
    a = np.array(a, dtype=np.float64)
    for i in range(1, a.shape[-1]):
        d = a[..., i] - a[..., i - 1]
        a[..., i] -= np.round(d / (2 * math.pi)) * (2 * math.pi)
    return a


class ShortTrajOpt:
    def __init__(self, skel, nf, markers, device=0):
        import torch
        self.torch = torch
        if not torch.cuda.is_available():
            raise RuntimeError("ipcdnnwalk_b200.short_traj_opt needs a CUDA device (sm_100a); there is no CPU fallback")
        if len(markers) != 4:
            raise ValueError("four markers (l toe, l heel, r toe, r heel) as in walk_SRBTrack2025.py:336-347")
        self.lib = _lib.load()
        self.device = torch.device("cuda", device)
        self.nf = int(nf)
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
        mb = np.ascontiguousarray([m["bone"] for m in markers], dtype=np.int32)
        ml = np.ascontiguousarray([m["lpos"] for m in markers], dtype=np.float64)
        mw = np.ascontiguousarray([m["weight"] for m in markers], dtype=np.float64)
        self._h = C.c_void_p()
        _check(self.lib, self.lib.mmsto_create(int(device), nb, _ptr(parent), _ptr(offset), _ptr(nchan), _ptr(axes), _ptr(mass), _ptr(com),
                                               _ptr(inertia), self.nf, _ptr(mb), _ptr(ml), _ptr(mw), C.byref(self._h)))
        self.ndof = self.lib.mmsto_ndof(self._h)
        self.dof_weights = np.ones(self.ndof)

    # ---- parameters
    def set(self, name, value):
        if name in ("weightEEvel", "weightHalfSpace", "weightRelativeHalfSpace", "weightAngleInput"):
            if name == "weightEEvel" and float(value) != 0.0 or name == "weightAngleInput" and float(value) != 0.0:
                raise NotImplementedError(f"{name} != 0 is outside the online configuration (walk_SRBTrack2025.py:529-537)")
            return                         # collision half-space weights have no effect without constraints
        _check(self.lib, self.lib.mmsto_set_param(self._h, name.encode(), float(value)))

    def get(self, name):
        v = C.c_double()
        _check(self.lib, self.lib.mmsto_get_param(self._h, name.encode(), C.byref(v)))
        return v.value

    def set_dof_weights(self, w=None):
        if w is not None:
            self.dof_weights = np.ascontiguousarray(w, dtype=np.float64)
        _check(self.lib, self.lib.mmsto_set_dof_weights(self._h, _ptr(np.ascontiguousarray(self.dof_weights, dtype=np.float64))))

    # ---- host-side conversions (shortTermOptimizer.lua:4-49; LoaderToTree::setPoseDOF / getPoseDOF with an Euler root)
    def calculateOriginalEulerPoses(self, pose, prev_ry=None):
        """pose [.., nframes, 7 + #angles] (root xyz, quaternion wxyz, angles) -> q [.., nframes, ndof]; ry aligned along the frames."""
        pose = np.asarray(pose, dtype=np.float64)
        q = np.concatenate([pose[..., 0:3], _quat_to_yzx(pose[..., 3:7]), pose[..., 7:]], axis=-1)
        ry = q[..., 3]
        if prev_ry is not None:
            ry = np.concatenate([np.asarray(prev_ry, dtype=np.float64)[..., None], ry], axis=-1)
        ry = align_euler_angles(ry)
        q[..., 3] = ry[..., 1:] if prev_ry is not None else ry
        return q

    def eulerToPoseDOF(self, q):
        q = np.asarray(q, dtype=np.float64)
        return np.concatenate([q[..., 0:3], _yzx_to_quat(q[..., 3:6]), q[..., 6:]], axis=-1)

    # ---- the solve
    def _dev(self, a, shape, dtype=None):
        torch = self.torch
        t = a if torch.is_tensor(a) else torch.as_tensor(np.ascontiguousarray(a))
        t = t.to(self.device, dtype or torch.float64).contiguous()
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f"expected shape {tuple(shape)}, got {tuple(t.shape)}")
        return t

    def pack_velcon(self, n, hsVelConInfo):
        """hsVelConInfo: per problem a dict {lua frame index (1-based): (min_speed[k], normals[k,3], tree_indices[k], local_positions[k,3])}
        (walk_SRBTrack2025.py:561-577; tree index = bone index + 1) -> [n, 4, 10]."""
        v = np.zeros((n, 4, 10)); v[:, :, 0] = -1
        for i, info in enumerate(hsVelConInfo or []):
            k = 0
            for lua_frame, (speed, normals, tis, lps) in (info or {}).items():
                for s, nr, ti, lp in zip(speed, normals, tis, lps):
                    if k >= 4:
                        raise ValueError("at most 4 velocity constraints per window")
                    v[i, k, 0] = lua_frame - 1; v[i, k, 1] = ti - 1; v[i, k, 2] = s; v[i, k, 3:6] = nr; v[i, k, 6:9] = lp
                    k += 1
        return v

    def solveEndEffectors_euler(self, euler_mot, gpos_target, x_original, dx, ddx, com, momentum, con=None, velcon=None):
        """All arrays [n, ...] as in include/mmsto_b200.h -> (x [n, nf, ndof] torch.float64 on the device, info [n, 4]).  Runs on torch's
        current stream (the staging tensors are created on it, so their copies are ordered before the kernel)."""
        torch = self.torch
        nf, nd = self.nf, self.ndof
        n = int(x_original.shape[0])
        xo = self._dev(x_original, (n, nf, nd)); dxd = self._dev(dx, (n, nf - 1, nd)); ddxd = self._dev(ddx, (n, nf - 2, nd))
        em = self._dev(euler_mot, (n, nf, nd)); gt = self._dev(gpos_target, (n, nf, 12)); cm = self._dev(com, (n, nf, 3))
        mo = self._dev(momentum, (n, nf - 1, 6))
        cn = None
        if con is not None:
            c = np.asarray(con.cpu() if torch.is_tensor(con) else con).astype(bool)          # [n, 4, nf]
            bits = np.zeros(n, dtype=np.int32)
            for m in range(4):
                for f in range(nf):
                    bits |= (c[:, m, f].astype(np.int32) << (m * 8 + f))
            cn = torch.as_tensor(bits, device=self.device)
        vc = self._dev(velcon, (n, 4, 10)) if velcon is not None else None
        x = torch.empty(n, nf, nd, dtype=torch.float64, device=self.device)
        info = torch.empty(n, 4, dtype=torch.float64, device=self.device)
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _check(self.lib, self.lib.mmsto_solve(self._h, n, _ptr(xo), _ptr(dxd), _ptr(ddxd), _ptr(em), _ptr(gt), _ptr(cm), _ptr(mo), _ptr(cn), _ptr(vc),
                                              _ptr(x), _ptr(info), st))
        return x, info

    def evaluate(self, x, gpos_target, x_original, dx, ddx, com, momentum, con=None, velcon=None):
        """Objective [n] and gradient [n, nf, ndof] of LBFGS_opta's function at x (test / diagnostic entry)."""
        torch = self.torch
        nf, nd = self.nf, self.ndof
        n = int(x_original.shape[0])
        xd = self._dev(x, (n, nf, nd))
        xo = self._dev(x_original, (n, nf, nd)); dxd = self._dev(dx, (n, nf - 1, nd)); ddxd = self._dev(ddx, (n, nf - 2, nd))
        gt = self._dev(gpos_target, (n, nf, 12)); cm = self._dev(com, (n, nf, 3)); mo = self._dev(momentum, (n, nf - 1, 6))
        cn = None
        if con is not None:
            c = np.asarray(con).astype(bool)
            bits = np.zeros(n, dtype=np.int32)
            for m in range(4):
                for f in range(nf):
                    bits |= (c[:, m, f].astype(np.int32) << (m * 8 + f))
            cn = torch.as_tensor(bits, device=self.device)
        vc = self._dev(velcon, (n, 4, 10)) if velcon is not None else None
        obj = torch.empty(n, dtype=torch.float64, device=self.device)
        grad = torch.empty(n, nf, nd, dtype=torch.float64, device=self.device)
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _check(self.lib, self.lib.mmsto_evaluate(self._h, n, _ptr(xd), _ptr(xo), _ptr(dxd), _ptr(ddxd), _ptr(gt), _ptr(cm), _ptr(mo), _ptr(cn), _ptr(vc),
                                                 _ptr(obj), _ptr(grad), st))
        return obj, grad

    # ---- the step after the solve (DeltaExtractor.py:421-512)
    def removeCOMsliding_online(self, com_srb, desired_com, output, fix_y_also=True, ground_h=None):
        """output [n, nvar, ndof + 1] poses (torch.float64 on the device, updated in place); com_srb, desired_com [n, nvar, 3];
        ground_h [n, nvar] terrain height under the current CoMs or None -> (currCOM, optCOM) [n, nvar, 3]."""
        torch = self.torch
        assert torch.is_tensor(output) and output.is_cuda and output.dtype == torch.float64 and output.is_contiguous()
        n, nvar, npose = output.shape
        if npose != self.ndof + 1:
            raise ValueError(f"poses have {npose} columns, expected {self.ndof + 1}")
        cs = self._dev(com_srb, (n, nvar, 3)); dc = self._dev(desired_com, (n, nvar, 3))
        gh = self._dev(ground_h, (n, nvar)) if ground_h is not None else None
        curr = torch.empty(n, nvar, 3, dtype=torch.float64, device=self.device); opt = torch.empty_like(curr)
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _check(self.lib, self.lib.mmsto_remove_com_sliding(self._h, n, nvar, _ptr(output), _ptr(cs), _ptr(dc), int(bool(fix_y_also)), _ptr(gh),
                                                           _ptr(curr), _ptr(opt), st))
        return curr, opt

    def windowCOM(self, poses):
        torch = self.torch
        p = self._dev(poses, tuple(poses.shape))
        n, nvar, _ = p.shape
        com = torch.empty(n, nvar, 3, dtype=torch.float64, device=self.device)
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _check(self.lib, self.lib.mmsto_window_com(self._h, n, nvar, _ptr(p), _ptr(com), st))
        return com

    def launch_count(self):
        return int(self.lib.mmsto_launch_count(self._h))

    def close(self):
        if self._h:
            self.lib.mmsto_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def extractSingleFrameFullbody(obs_extra, recon, srb_mass, srb_inertia, srb_local_com=(0.0, 0.0, 0.0), device=0):
    """Batched `DeltaExtractor._extractSingleFrameFullbody` (DeltaExtractor.py:577-643).  obs_extra [n, 29] (pack_obs_extra
    records), recon [n, 214] (decoder rows), the SRB segment's mass / 3x3 inertia / local CoM -> dict of torch.float64 device
    tensors: humanPose [n, 60], dx_ddx [n, 118], feetpos [n, 24], com [n, 3], momentum [n, 6], con_bin [n, 4], con [n, 4]."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("ipcdnnwalk_b200.short_traj_opt needs a CUDA device (sm_100a); there is no CPU fallback")
    lib = _lib.load()
    dev = torch.device("cuda", device)

    def to_dev(a, w):
        t = a if torch.is_tensor(a) else torch.as_tensor(np.ascontiguousarray(a))
        t = t.to(dev, torch.float64).contiguous()
        if t.dim() != 2 or t.shape[1] != w:
            raise ValueError(f"expected [n, {w}], got {tuple(t.shape)}")
        return t
    ex = to_dev(obs_extra, 29); rw = to_dev(recon, 214)
    n = ex.shape[0]
    if rw.shape[0] != n:
        raise ValueError("obs_extra and recon must have the same number of rows")
    out = {k: torch.empty(n, w, dtype=torch.float64, device=dev) for k, w in
           (("humanPose", 60), ("dx_ddx", 118), ("feetpos", 24), ("com", 3), ("momentum", 6), ("con_bin", 4), ("con", 4))}
    I = np.ascontiguousarray(np.asarray(srb_inertia, dtype=np.float64).reshape(3, 3))
    lc = np.ascontiguousarray(np.asarray(srb_local_com, dtype=np.float64).reshape(3))
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    _check(lib, lib.fullbody_extract(int(device), n, _ptr(ex), _ptr(rw), float(srb_mass), _ptr(I), _ptr(lc), _ptr(out["humanPose"]),
                                     _ptr(out["dx_ddx"]), _ptr(out["feetpos"]), _ptr(out["com"]), _ptr(out["momentum"]),
                                     _ptr(out["con_bin"]), _ptr(out["con"]), st))
    return out


def removeFootSliding_online(footLen, conAll, initialFeet, markerTraj, marker, desiredVel=None, projectToGround=None, footCenterTraj=None, device=0):
    """Batched `DeltaExtractor.removeFootSliding_online` (DeltaExtractor.py:650-760) and, when `footCenterTraj` is given,
    `removeFootSliding_online_stair` (:763-893).  conAll [n, nvar, 4], initialFeet [n, 12], markerTraj [n, nvar, 12] (the four smooth
    marker trajectories side by side: toe L, heel L, toe R, heel R), marker [n, nvar, 12], desiredVel [n, nvar, 12] or None,
    footCenterTraj [n, 2, nvar, 3] or None.  projectToGround: None (no projection), a float (flat ground at that height, what the flat
    driver's `v.y = 0` does) or an SRBVecEnv of the terrain variant (the playground under the character).  -> optMarkerTraj [n, nvar, 12]
    (torch.float64 on the device)."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("ipcdnnwalk_b200.short_traj_opt needs a CUDA device (sm_100a); there is no CPU fallback")
    lib = _lib.load()
    dev = torch.device("cuda", device)

    def to_dev(a, shape):
        if a is None:
            return None
        t = a if torch.is_tensor(a) else torch.as_tensor(np.ascontiguousarray(a))
        t = t.to(dev, torch.float64).contiguous()
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f"expected {tuple(shape)}, got {tuple(t.shape)}")
        return t
    ca = conAll if torch.is_tensor(conAll) else torch.as_tensor(np.ascontiguousarray(conAll))
    n, nvar = int(ca.shape[0]), int(ca.shape[1])
    ca = to_dev(ca, (n, nvar, 4)); fi = to_dev(initialFeet, (n, 12)); mt = to_dev(markerTraj, (n, nvar, 12)); mk = to_dev(marker, (n, nvar, 12))
    dv = to_dev(desiredVel, (n, nvar, 12)); fc = to_dev(footCenterTraj, (n, 2, nvar, 3))
    mode, gy, henv = 0, 0.0, None
    if projectToGround is not None:
        if hasattr(projectToGround, "_h"):
            mode, henv = 2, projectToGround._h
        else:
            mode, gy = 1, float(projectToGround)
    out = torch.empty(n, nvar, 12, dtype=torch.float64, device=dev)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    _check(lib, lib.mmsto_remove_foot_sliding(int(device), n, nvar, float(footLen), _ptr(ca), _ptr(fi), _ptr(mt), _ptr(mk), _ptr(dv), _ptr(fc),
                                              mode, gy, henv, _ptr(out), st))
    return out
