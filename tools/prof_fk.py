"""Short driver for ncu captures of the FK kernels. usage: prof_fk.py DTYPE(32|64) [generic|direct|tma|tma1|tma3|tma4]"""
import os, sys
import torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
from ipcdnnwalk_b200 import vrml_skeleton, Skeleton
dt = torch.float32 if (len(sys.argv) < 2 or sys.argv[1] == "32") else torch.float64
mode = sys.argv[2] if len(sys.argv) > 2 else "tma"
sk = Skeleton(vrml_skeleton.builtin())
sk.set_generic(mode == "generic"); sk.set_tma(int(mode[3:]) if mode.startswith("tma") and len(mode) > 3 else (mode == "tma"))
n = 1 << 20
g = torch.Generator(device="cuda"); g.manual_seed(4)
q = torch.randn(n, 4, device="cuda", generator=g, dtype=dt); q = q / q.norm(dim=1, keepdim=True)
pose = torch.cat([torch.rand(n, 3, device="cuda", generator=g, dtype=dt) * 10 - 5, q, torch.rand(n, 53, device="cuda", generator=g, dtype=dt) * 2 - 1], 1).contiguous()
dq = (torch.rand(n, 59, device="cuda", generator=g, dtype=dt) * 2 - 1).contiguous()
out = sk.forward(pose, dq)
for _ in range(3):
    sk.forward(pose, None, out=out)
    sk.forward(pose, dq, out=out)
torch.cuda.synchronize()
print("ok")
