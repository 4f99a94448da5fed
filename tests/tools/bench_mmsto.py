"""GPU throughput of the batched MMSTO solve: windows per second and objective evaluations per second at a
fixed iteration budget, plus the numpy oracle on one host core for scale.  Usage: python tests/tools/bench_mmsto.py [n] [max_iterations]"""
import json, os, sys, time
import numpy as np
import torch
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import helpers_mmsto as hm
from ipcdnnwalk_b200 import ShortTrajOpt

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 30
skel = hm.skeleton()
base = [hm.make_problem(s, skel=skel) for s in range(16)]
ps = [base[i % 16] for i in range(n)]
keys = ("gpos_target", "x_original", "dx", "ddx", "com", "momentum")
dev = [torch.as_tensor(hm.stack(ps, k), device="cuda") for k in keys]
em = torch.as_tensor(hm.stack(ps, "euler_mot"), device="cuda"); vc = torch.as_tensor(hm.stack(ps, "velcon"), device="cuda")
opt = ShortTrajOpt(skel, hm.NF, hm.MARKERS)
opt.set_dof_weights(hm.ankle_dof_weights(skel))
opt.set("max_iterations", iters)
for _ in range(2):
    x, info = opt.solveEndEffectors_euler(em, *dev, velcon=vc)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); x, info = opt.solveEndEffectors_euler(em, *dev, velcon=vc); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = sorted(ts)[len(ts) // 2]
ev = float(info[:, 3].sum().item())
out = {"kernel": "mmsto_kernel<false>", "windows": n, "max_iterations": iters, "ms": ms, "windows_per_s": n / ms * 1e3,
       "evaluations": ev, "evaluations_per_s": ev / ms * 1e3, "evals_per_window": ev / n, "dtype": "f64"}
if "--cpu" in sys.argv:
    o = hm.oracle_opt(base[0], skel, dof_weights=hm.ankle_dof_weights(skel))
    t = time.perf_counter(); X, inf = o.solve(base[0]["euler_mot"], max_iterations=iters); dt = time.perf_counter() - t
    out["oracle_numpy_1core"] = {"s_per_window": dt, "evaluations": inf["evaluations"], "evaluations_per_s": inf["evaluations"] / dt}
print(json.dumps(out))
if "--comslide" in sys.argv:
    # removeCOMsliding_online on windows of 8 poses: algorithmic bytes = poses in + root out + CoM targets in
    from ipcdnnwalk_b200.short_traj_opt import _yzx_to_quat
    nw = 1 << 17
    q = np.stack([b["x_original"] for b in base])                           # [16, 7, 59]
    q = np.concatenate([q, q[:, -1:]], 1)                                    # 8 frames
    pose = np.concatenate([q[..., 0:3], _yzx_to_quat(q[..., 3:6]), q[..., 6:]], -1)
    pose = torch.as_tensor(pose, device="cuda").repeat(nw // 16, 1, 1).contiguous()
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    cs = (torch.randn(nw, 8, 3, device="cuda", generator=g, dtype=torch.float64) * 0.02).cumsum(1)
    dc = (torch.randn(nw, 8, 3, device="cuda", generator=g, dtype=torch.float64) * 0.02).cumsum(1)
    work = pose.clone()
    for _ in range(3):
        work.copy_(pose); opt.removeCOMsliding_online(cs, dc, work)
    ts = []
    for _ in range(10):
        work.copy_(pose)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); opt.removeCOMsliding_online(cs, dc, work); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    by = nw * 8 * (60 + 7 + 6 + 6) * 8
    print(json.dumps({"kernel": "comslide_kernel", "windows": nw, "poses_per_window": 8, "ms": ms, "windows_per_s": nw / ms * 1e3,
                      "algorithmic_GBps": by / ms / 1e6, "note": "timing includes the two small output allocations of the Python mirror"}))
if "--extract" in sys.argv:
    from ipcdnnwalk_b200.short_traj_opt import extractSingleFrameFullbody
    ni = 1 << 20
    g = torch.Generator(device="cuda"); g.manual_seed(2)
    ex = torch.randn(ni, 29, device="cuda", generator=g, dtype=torch.float64) * 0.5
    for c0 in (3, 16, 23):
        ex[:, c0:c0 + 4] /= ex[:, c0:c0 + 4].norm(dim=1, keepdim=True)
    rw = torch.randn(ni, 214, device="cuda", generator=g, dtype=torch.float64) * 0.3
    I = np.diag([3.0, 1.5, 2.5])
    for _ in range(3):
        extractSingleFrameFullbody(ex, rw, 60.0, I)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); extractSingleFrameFullbody(ex, rw, 60.0, I); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    by = ni * (29 + 214 + 60 + 118 + 24 + 3 + 6 + 4 + 4) * 8
    print(json.dumps({"kernel": "fullbody_extract_kernel", "items": ni, "ms": ms, "items_per_s": ni / ms * 1e3, "algorithmic_GBps": by / ms / 1e6,
                      "note": "timing includes the seven output allocations of the Python mirror"}))
opt.close()
