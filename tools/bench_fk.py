"""GPU throughput of the batched FK / CoM / momentum kernel (BASELINE config 5: 1M humanoid poses).
Prints poses/s and the HBM-roofline fraction (836 B/pose fp32, 1672 B/pose fp64: pose in, 20x7 transforms + CoM + momentum out;
dq input adds 236/472 B and is reported separately)."""
import json, os, sys
import torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
from ipcdnnwalk_b200 import vrml_skeleton, Skeleton

# Under torchrun (BASELINE config 5: 1M poses over 1/2/4/8 GPUs) the poses are sharded over the ranks (strong scaling, no
# collective on the data path); times are the max over ranks and rank 0 prints whole-job poses/s.
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
dist = None
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
skel = vrml_skeleton.builtin()
sk = Skeleton(skel, device=local)
n_total = 1 << 20
n = n_total // world
peak = 6547.8
try:
    peak = json.load(open(os.path.join(R, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
g = torch.Generator(device="cuda"); g.manual_seed(4)
for dt, name in ((torch.float32, "f32"), (torch.float64, "f64")):
    q = torch.randn(n, 4, device="cuda", generator=g, dtype=dt); q = q / q.norm(dim=1, keepdim=True)
    pose = torch.cat([torch.rand(n, 3, device="cuda", generator=g, dtype=dt) * 10 - 5, q, torch.rand(n, 53, device="cuda", generator=g, dtype=dt) * 2 - 1], 1).contiguous()
    dq = (torch.rand(n, 59, device="cuda", generator=g, dtype=dt) * 2 - 1).contiguous()
    out = sk.forward(pose, dq)
    for generic in (("auto",) if world > 1 else ("auto", "tma1", "tma3", "tma4", "direct", "generic")):
      sk.set_generic(generic == "generic"); sk.set_tma(int(generic[3:]) if generic.startswith("tma") else (-1 if generic == "auto" else 0))
      for mode, d in (("fk+com (no velocities)", None), ("fk+com+momentum", dq)):
        for _ in range(3):
            sk.forward(pose, d, out=out)
        ts = []
        for _ in range(10):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); sk.forward(pose, d, out=out); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        if dist is not None:
            tm = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            ms = float(tm.item())
            if rank != 0:
                continue
        es = pose.element_size()
        alg = (60 + 140 + 3 + (6 if d is not None else 0)) * es
        tot = alg + (59 * es if d is not None else 0)
        print(json.dumps({"kernel": {"generic": "fk_forward_kernel (structure-agnostic)", "tma1": "fk_hanyang_tma_rolled_kernel (64-pose TMA tiles)", "auto": "default dispatch", "tma3": "fk_hanyang_tma_unrolled_kernel (64-pose TMA tiles)", "tma4": "fk_hanyang_tma_chain_kernel (64-pose TMA tiles)", "direct": "fk_hanyang_kernel (skeleton-specialised, direct rows)"}[generic], "dtype": name, "mode": mode, "poses": n * world, "n_gpus": world, "ms": ms, "poses_per_s": n * world / ms * 1e3,
                          "algorithmic_bytes_per_pose": alg, "achieved_GBps_algorithmic_per_gpu": alg * n / ms / 1e6,
                          "achieved_GBps_incl_dq_per_gpu": tot * n / ms / 1e6, "peak_GBps": peak, "frac": alg * n / ms / 1e6 / peak}))
sk.close()
if dist is not None:
    dist.destroy_process_group()
