#!/usr/bin/env python
"""bench.py -- SRB env-steps/s of the batched SRBTrack step (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA step through the C-ABI)
  python bench.py --impl reference --gpus N --steps K ...  reference arm: the CPU oracle port of the
                                                           reference's SRBEnv.step on all host cores

Workload (config.workload): BASELINE.json configs[1] -- SRBEnv5_mocap_list_window flat ground,
4096 environments per GPU (weak scaling: N GPUs step N*4096 environments, no data-path collective;
the per-rollout NCCL all-reduce of observation-normaliser statistics is inside the timed region).
One "step" = one SRBEnv.step of every environment = 4 physics sub-steps + obs/reward/done (+ fused
auto-reset).  Actions come from the device-side reference-tracking controller + N(0, 0.1^2) noise and
are produced outside the timed events, like a policy would.

Timing: CUDA events on the launching stream around every step launch; L2 is flushed (256 MiB write)
between steps because one step's working set (~6 MB) is far below the 126 MB L2; max over ranks.
The path's one collective -- observation-normaliser statistics + episode counters, one NCCL all-reduce per
rollout chunk -- runs every ROLLOUT steps and once more at the end of the timed loop, so it is inside the timed
total whatever --steps is; its time is reported as `collective_ms`.

The same JSON line carries `secondary` records (each timed the same way, after the headline loop): the fp64
(reference precision) headline, BASELINE configs[2] (terrain, 8192 envs per GPU), configs[3] (gym_cdm2, 16384
envs), configs[4] (FK of 1 M poses, split over the GPUs) and the batched MMSTO solve (4096 windows).
Reference tables are built by the product's own ClipBuilder from the raw clips under tests/golden/gym_trackSRB.
"""
import argparse
import json
import os

# one BLAS/OpenMP thread per process, as the reference's trainer sets (walk_SRBTrack2025_train_delta.py:6-10);
# must happen before numpy is imported.  The CPU arm parallelises over forked workers instead.
for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ.setdefault(_v, "1")
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

RAW_CLIPS = ["ref_contact/stand1.motion.npz", "ref_contact/aiming_show.motion.npz"]   # tests/golden/gym_trackSRB/... (raw mocap + contact flags)
ENVS_PER_GPU = 4096
SIGMA = 0.1
ROLLOUT = 128                 # steps between observation-statistics all-reduces (one per PPO rollout chunk)
BYTES_PER_STEP = {32: 1008, 64: 1320}   # algorithmic HBM bytes per env-step, SURVEY.md 8(d) / DESIGN.md
METRIC = "SRB env-steps/sec"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=512)
    ap.add_argument("--warmup", type=int, default=16)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=ENVS_PER_GPU)
    ap.add_argument("--precision", type=int, default=int(os.environ.get("SRB_BENCH_PRECISION", "32")), choices=[32, 64])
    ap.add_argument("--lanes", type=int, default=0)
    ap.add_argument("--variant", default="test", choices=["test", "training", "terrain"],
                    help="terrain = BASELINE config 3 (SRBEnv5_mocap_list_terrain_window on the playground height fields); not the headline workload")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--collective", default="fused", choices=["fused", "nccl"],
                    help="the per-rollout exchange: fused = statistics + all-reduce in ONE kernel over NVLink peer memory (srb_rollout_stats_allreduce); "
                         "nccl = statistics kernel + torch.distributed all-reduce")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary records (fp64, terrain, gym_cdm2 flat + terrain, FK, MMSTO)")
    ap.add_argument("--flush", default="write", choices=["write", "read", "none"],
                    help="L2 flush between steps: write a 256 MiB buffer (default, the timing rule), read it, or nothing (diagnostic)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU arm
TERRAIN_NPZ = os.path.join(ROOT, "tests", "golden", "terrain_playground.npz")
# where the environments of the terrain workload stand on Characters/SRB5_playground.xml: flat plane, pyramid, hill and rough
# height fields (the four regions the parity tests cover), jittered per environment
TERRAIN_SPOTS = ((0.0, 0.0), (22.0, 24.0), (-42.0, 42.0), (42.0, -40.0))


# The terrain env's `done` (absolute height < 0.5 or > 1.2 m, env/SRBEnv5_mocap_list_terrain_window.py:642-700) is true on
# every elevated part of the playground and walk_SRBTrack2025_terrain_singlethreaded.py:421,724 ignores it, so the terrain
# workload ignores it too: no auto-reset, all environments are re-seeded every TERRAIN_EPISODE steps outside the timed events.
TERRAIN_EPISODE = 32


def terrain_planar(n, seed):
    rng = np.random.default_rng(seed)
    spots = np.array(TERRAIN_SPOTS)[np.arange(n) % len(TERRAIN_SPOTS)]
    return spots + rng.uniform(-1.5, 1.5, (n, 2))


def usable_cores():
    """Host threads this process may really use: affinity mask capped by the cgroup CPU quota."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    try:
        q = open("/sys/fs/cgroup/cpu.max").read().split()
        if q[0] != "max":
            n = max(1, min(n, int(float(q[0]) / float(q[1]))))
    except Exception:
        pass
    return n


def _oracle_worker(args):
    """One host core: `n_env` oracle environments; `warmup` passes, then `steps` passes (or until
    `budget_s` seconds when steps is None).  Returns (env-steps, seconds)."""
    wid, n_env, warmup, steps, budget_s, variant = args
    import helpers
    from oracle.srb_env import OracleSRBEnv
    rng = np.random.default_rng(1000 + wid)
    clips = helpers.fresh_clips()
    envs = []
    if variant == "terrain":
        from oracle.terrain import Terrain
        from oracle.srb_env import lift_clip_to_terrain
        T = Terrain.load(TERRAIN_NPZ)
        planar = terrain_planar(n_env, 500 + wid)
        clips = helpers.fresh_clips(("aiming_show",))
    for i in range(n_env):
        if variant == "terrain":
            e = OracleSRBEnv([lift_clip_to_terrain(helpers.load_clip("aiming_show"), T, tuple(planar[i]))], "terrain", terrain=T)
        else:
            e = OracleSRBEnv(helpers.fresh_clips(), variant)
        helpers.oracle_reset(e, helpers.random_plan(rng, clips))
        envs.append(e)

    passes = [0]

    def one_pass():
        passes[0] += 1
        for e in envs:
            a = helpers.tracking_action(e, rng, SIGMA)
            _, _, done, _, _ = e.step(a)
            # terrain: `done` is ignored as the reference's terrain script does (see TERRAIN_EPISODE)
            if (done and variant != "terrain") or (variant == "terrain" and passes[0] % TERRAIN_EPISODE == 0):
                helpers.oracle_reset(e, helpers.random_plan(rng, clips))
    for _ in range(warmup):
        one_pass()
    t0 = time.perf_counter()
    n = 0
    while (steps is not None and n < steps) or (steps is None and time.perf_counter() - t0 < budget_s):
        one_pass()
        n += 1
    return n_env * n, time.perf_counter() - t0


def cpu_oracle_throughput(cores, n_env_per_core, warmup, steps, variant, budget_s=None):
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        res = pool.map(_oracle_worker, [(w, n_env_per_core, warmup, steps, budget_s, variant) for w in range(cores)])
    total = sum(r[0] for r in res)
    tmax = max(r[1] for r in res)
    return total / tmax, total, tmax


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = usable_cores()
    # calibrate one pass (1 env per core), then size the per-step sample so that warmup+steps passes take ~40 s
    _, _, t1 = cpu_oracle_throughput(cores, 1, 1, 2, args.variant)
    per_pass = max(t1 / 2, 1e-3)
    per_core = int(max(1, min(16, 40.0 / (per_pass * max(args.steps + args.warmup, 1)))))
    value, total, tmax = cpu_oracle_throughput(cores, per_core, args.warmup, args.steps, args.variant)
    sample = (f"{cores} forked single-thread workers x {per_core} envs x {args.steps} steps of the oracle port of SRBEnv.step "
              f"(variant {args.variant}, clips stand1+aiming_show, tracking actions + N(0,{SIGMA}^2)); {total} env-steps in {tmax:.1f} s")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tmax / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": (f"SRBEnv5_mocap_list_terrain_window on the playground terrain, bounded sample of {cores * per_core} envs per step on {cores} host cores"
                                if args.variant == "terrain" else
                                f"SRBEnv5_mocap_list_window flat ground ({args.variant} variant), bounded sample of {cores * per_core} envs per step on {cores} host cores")},
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args, rank, local_rank, world):
    import torch
    import helpers
    from ipcdnnwalk_b200 import SRBVecEnv, OBS_DIM, ACT_DIM, RINFO_DIM
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    N = args.envs
    # reference tables: the product's ClipBuilder (device FK of the MJCF humanoid) over the raw clips, SRBEnv.loadMocap's job
    from ipcdnnwalk_b200.srb_env import build_clip_tables
    all_clips = build_clip_tables(RAW_CLIPS, data_root=os.path.join(ROOT, "tests", "golden"), device=local_rank)
    if args.variant == "terrain":
        from ipcdnnwalk_b200 import TerrainModel
        clips = all_clips[1:]
        env = SRBVecEnv(clips, N, variant="terrain", precision=args.precision, lanes_per_env=args.lanes, auto_reset=False,
                        device=local_rank, seed=1234 + rank, terrain=TerrainModel.builtin_playground(),
                        initial_pos_planar=terrain_planar(N, 77 + rank))
    else:
        clips = all_clips
        env = SRBVecEnv(clips, N, variant=args.variant, precision=args.precision, lanes_per_env=args.lanes, auto_reset=True,
                        device=local_rank, seed=1234 + rank)
    env.reset()
    act = torch.zeros(N, ACT_DIM, dtype=torch.float32, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    flush_sink = torch.zeros((), dtype=torch.int64, device=dev)
    acc = torch.zeros(1 + 2 * OBS_DIM + 3, dtype=torch.float64, device=dev)   # (count, sum x, sum x^2) + episode (sum return, sum length, count)
    K, W = args.steps, args.warmup

    n_resets = [0]

    def one_step(k, timed):
        if args.variant == "terrain" and k % TERRAIN_EPISODE == 0 and k > 0:
            n_resets[0] += int(timed)
            env.reset()                                                     # episode boundary of the terrain workload, untimed
        env.tracking_action(act, sigma=SIGMA, seed=99 + rank, step=k)      # the "policy": outside the timed events
        if args.flush == "write":
            flush.fill_(k & 0xFF)                                           # L2 flush between steps
        elif args.flush == "read":
            flush_sink.add_(flush.view(torch.int64).sum())                  # diagnostic: evict with clean lines
        ev = None
        if timed:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        env.step(act)
        if timed:
            ev[1].record()
        cev = None
        if (k + 1) % ROLLOUT == 0 or (timed and k == W + K - 1):           # the path's one collective: once per rollout
            if timed:                                                       # chunk and at the end of the timed loop
                cev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                cev[0].record()
            if fused:
                env.rollout_stats_allreduce(acc)                            # statistics + all-reduce over NVLink peer memory: ONE kernel
            else:
                acc.zero_()                                                 # one rollout chunk's sufficient statistics
                env.rollout_stats(acc)                                      # obs statistics kernel + episode counters (device copy)
                if dist is not None:
                    dist.all_reduce(acc)                                    # NCCL over NVLink
            if timed:
                cev[1].record()
        return ev, cev

    fused = args.collective == "fused"
    fused_check = None
    if fused:
        # set-up (CUDA IPC handles of the ranks' exchange blocks) + untimed cross-check of the fused collective against the statistics
        # kernel + NCCL on the same observations; any failure (no peer access on this box, ...) falls back to the NCCL form on ALL ranks
        ok = 1
        try:
            env.connect_ranks(dist, rank, world)
            ref = torch.zeros_like(acc)
            env.rollout_stats(ref)
            if dist is not None:
                dist.all_reduce(ref)
            env.rollout_stats_allreduce(acc)
            torch.cuda.synchronize()
            ok = int(bool(torch.allclose(acc, ref, rtol=1e-12, atol=1e-9)) and env.xchg_status() == 0)
        except Exception as e:                                              # noqa: BLE001
            print(f"[bench] rank {rank}: fused collective unavailable ({e}); using NCCL", file=sys.stderr)
            ok = 0
        okt = torch.tensor([ok], dtype=torch.int32, device=dev)
        if dist is not None:
            dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        fused = bool(int(okt.item()))
        fused_check = fused
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for k in range(W):
        one_step(k, False)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = env.launch_count()
    t_wall0 = time.perf_counter()
    evs, cevs = [], []
    for k in range(W, W + K):
        ev, cev = one_step(k, True)
        evs.append(ev)
        if cev is not None:
            cevs.append(cev)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    step_ms = np.array([a.elapsed_time(b) for a, b in evs])
    coll_ms = float(sum(a.elapsed_time(b) for a, b in cevs))
    total_ms = float(step_ms.sum()) + coll_ms
    n_tracking = K
    step_launches = env.launch_count() - launches0 - n_tracking - len(cevs) - n_resets[0]

    # ---- end to end through the host-buffer C-ABI call (srb_step_host): pinned host actions in, obs/rew/done out
    e2e_ms = None
    Ke = min(K, 200)
    if not args.no_e2e:
        pin = env.pinned()
        host_mode = int(os.environ.get("SRB_HOST_MODE", "1"))     # experiment switch: 0 staged copies, 1 zero copy (library default), 2 split
        if host_mode != 1:
            env.set_host_mode(host_mode)
        for k in range(3):
            env.step_host()
        tt = 0.0
        for k in range(Ke):
            env.tracking_action(act, sigma=SIGMA, seed=7 + rank, step=10_000 + k)
            pin["act"][...] = act.cpu().numpy()          # the policy's output lands in pinned host memory (untimed)
            flush.fill_(k & 0xFF)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            env.step_host()                               # H2D actions + kernel + D2H obs/rew/done/info + sync
            tt += time.perf_counter() - t0
        e2e_ms = 1e3 * tt
    clocks = sampler.stop() if rank == 0 else None     # sampled from warm-up through the end-to-end loop
    # ---- max over ranks
    tvec = torch.tensor([total_ms, e2e_ms if e2e_ms is not None else 0.0, float(np.median(step_ms))], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tvec, op=dist.ReduceOp.MAX)
    total_ms_max, e2e_ms_max, med_ms_max = (float(x) for x in tvec.cpu())
    ep = acc[301:304].cpu().numpy() if args.variant != "terrain" else np.zeros(3)   # all-reduced episode counters (whole job)
    secondary = [] if args.no_secondary else run_secondary(args, rank, local_rank, world, dist, all_clips, flush)

    if rank == 0:
        peaks = {}
        src = "measured"
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            peak = float(peaks["hbm_gbs"])
        except Exception:
            peak, src = 6650.0, "fallback"
        bytes_step = BYTES_PER_STEP[args.precision] * N
        kern_ms = float(step_ms.mean())
        achieved = bytes_step / (kern_ms * 1e-3) / 1e9
        traffic, traffic_detail = None, None
        for tp in ("r02_traffic.json", "r01_traffic.json"):                # ncu --set full capture of the same command (constants, not measured in this run)
            tp = os.path.join(ROOT, "profiles", tp)
            if traffic is None and os.path.exists(tp):
                try:
                    tj = json.load(open(tp))
                    tv = tj.get(f"fp{args.precision}_n{N}")
                    if isinstance(tv, dict):
                        traffic = float(tv["read"]) + float(tv["write"])
                        traffic_detail = dict(tv, source=f"ncu constant from profiles/{os.path.basename(tp)} (dram__bytes_read.sum / dram__bytes_write.sum per launch, cold cache)")
                    elif tv is not None:
                        traffic = float(tv)
                        traffic_detail = {"read": float(tv), "write": 0.0, "source": f"ncu constant from profiles/{os.path.basename(tp)}: reads only, write-backs leave the 126 MB L2 after the kernel"}
                except Exception:
                    traffic = None
        value = world * N * K / (total_ms_max * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.precision == 32 else "f64", "data": "synthetic",
            "config": {"workload": (f"SRBEnv5_mocap_list_terrain_window on the playground terrain (plane / pyramid / hill / rough height field), {N} envs per GPU x {world} GPU"
                                    if args.variant == "terrain" else
                                    f"SRBEnv5_mocap_list_window flat ground ({args.variant} variant), {N} envs per GPU x {world} GPU"),
                       "envs_per_gpu": N, "clips": ("aiming_show, lifted per environment in the kernel" if args.variant == "terrain" else
                                                    "stand1 + aiming_show (shipped clips; walk/run clips are absent from the reference mount)"),
                       "actions": (f"device-side reference-tracking controller + N(0,{SIGMA}^2); done ignored as walk_SRBTrack2025_terrain_singlethreaded.py does, "
                                   f"all environments re-seeded every {TERRAIN_EPISODE} steps outside the timed events" if args.variant == "terrain" else
                                   f"device-side reference-tracking controller + N(0,{SIGMA}^2), auto-reset on done"),
                       "l2": {"write": "flushed between steps (256 MiB write)", "read": "DIAGNOSTIC: evicted between steps by reading 256 MiB", "none": "DIAGNOSTIC: not flushed"}[args.flush] + "; per-step CUDA events on the launching stream",
                       "collective": (f"ONE kernel: obs-normaliser statistics + episode counters + their all-reduce over NVLink peer memory (304 doubles, srb_rollout_stats_allreduce; checked against the NCCL path before timing: {fused_check})" if fused else
                                      "obs-normaliser statistics kernel + episode counters, one NCCL all-reduce (304 doubles)") + f" every {ROLLOUT} steps and at the end of the timed loop: {len(cevs)} inside the timed total",
                       "lanes_per_env": env_lanes(args), "median_step_ms": med_ms_max, "wall_s_timed_loop": t_wall,
                       "episodes_finished": int(ep[2]), "mean_episode_len": float(ep[1] / max(ep[2], 1))},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": f"{src} (MEASURED_PEAKS.json hbm_gbs, burst copy)", "algorithmic_bytes_per_env_step": BYTES_PER_STEP[args.precision],
                         "traffic_detail": traffic_detail, "kernel": "srb_step_kernel", "avg_launch_ms": kern_ms},
            "collective_ms": coll_ms, "collectives_timed": len(cevs),
            "clocks": clocks,
            "gpu_launches": int(step_launches),
        }
        if e2e_ms is not None:
            line["e2e"] = {"value": world * N * Ke / (e2e_ms_max * 1e-3), "unit": "env-steps/s",
                           "h2d_bytes_per_step": N * ACT_DIM * 4, "d2h_bytes_per_step": N * (OBS_DIM * 4 + 4 + 1 + RINFO_DIM * 4),
                           "steps": Ke, "host_mode": host_mode,
                           "api": "srb_step_host, host buffers: the coalesced-I/O step kernel reads the actions from and writes obs/reward/done/info to mapped pinned host memory over PCIe inside the timed region (zero copy), then stream sync" if host_mode == 1 else
                                  ("srb_step_host, host buffers, staged copy-engine transfers (H2D actions, kernel, D2H obs/reward/done/info) inside the timed region" if host_mode == 0 else
                                   "srb_step_host, host buffers, split mode (look-ahead windows by copy engine during the step kernel, the rest zero copy) inside the timed region")}
        if secondary:
            line["secondary"] = secondary
        if not args.no_cpu_baseline:
            cores = usable_cores()
            v, total, tmax = cpu_oracle_throughput(cores, 1, 1, None, args.variant, budget_s=15.0)
            line["cpu_baseline"] = {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port",
                                    "sample": f"{cores} forked single-thread workers x 1 env, oracle port of SRBEnv.step for 15 s ({total} env-steps in {tmax:.1f} s)"}
        print(json.dumps(line), flush=True)
    env.close()
    if dist is not None:
        dist.destroy_process_group()


def _timed(fn, steps, warmup, flush, dist, per_step_setup=None):
    """CUDA-event time of `steps` calls of fn() after `warmup` calls, L2 flushed between calls; mean ms, max over ranks."""
    import torch
    for i in range(warmup):
        if per_step_setup:
            per_step_setup(i)
        fn()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    evs = []
    for i in range(steps):
        if per_step_setup:
            per_step_setup(warmup + i)
        flush.fill_(i & 0xFF)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def run_secondary(args, rank, local_rank, world, dist, all_clips, flush):
    """The other BASELINE configurations and the MMSTO solve, each one record: value (whole job), roofline fraction of its kernel
    (algorithmic bytes per unit from SURVEY 8d / DESIGN.md), kernel name.  Same timing rules as the headline loop."""
    import torch
    from ipcdnnwalk_b200 import SRBVecEnv, CDMVecEnv, TerrainModel, Skeleton, ShortTrajOpt, vrml_skeleton
    import helpers_cdm as hc
    dev = torch.device("cuda", local_rank)
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    S, Wm = 30, 5
    recs = []

    def rec(name, config, kernel, units_per_gpu, ms, bytes_per_unit, unit, scaling, dtype, extra=None):
        value = units_per_gpu * world / (ms * 1e-3)
        ach = bytes_per_unit * units_per_gpu / (ms * 1e-3) / 1e9
        r = {"name": name, "config": config, "kernel": kernel, "value": value, "unit": unit, "ms_per_step": ms, "dtype": dtype, "n_gpus": world, "scaling": scaling,
             "steps": S, "warmup": Wm, "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "algorithmic_bytes_per_unit": bytes_per_unit}}
        if extra:
            r.update(extra)
        recs.append(r)

    # ---- headline workload at the reference's precision (fp64)
    N = args.envs
    env = SRBVecEnv(all_clips, N, variant="test", precision=64, auto_reset=True, device=local_rank, seed=1234 + rank)
    env.reset()
    act = torch.zeros(N, 22, dtype=torch.float32, device=dev)
    ms = _timed(lambda: env.step(act), S, Wm, flush, dist, lambda k: env.tracking_action(act, sigma=SIGMA, seed=5 + rank, step=k))
    rec("flat_fp64", f"BASELINE configs[1] at reference precision: flat ground, {N} envs per GPU", "srb_step_kernel<double>", N, ms, BYTES_PER_STEP[64], "env-steps/s", "weak", "f64")
    env.close()
    # ---- configs[2]: terrain, 8192 envs per GPU (65536 on 8)
    N = 8192
    env = SRBVecEnv(all_clips[1:], N, variant="terrain", precision=32, auto_reset=False, device=local_rank, seed=99 + rank,
                    terrain=TerrainModel.builtin_playground(), initial_pos_planar=terrain_planar(N, 77 + rank))
    env.reset()
    act = torch.zeros(N, 22, dtype=torch.float32, device=dev)

    def terrain_setup(k):
        if k % TERRAIN_EPISODE == 0 and k > 0:
            env.reset()
        env.tracking_action(act, sigma=SIGMA, seed=6 + rank, step=k)
    ms = _timed(lambda: env.step(act), S, Wm, flush, dist, terrain_setup)
    rec("terrain_fp32", f"BASELINE configs[2] per-GPU share: SRBEnv5_mocap_list_terrain_window on the playground (plane / pyramid / hill / rough), {N} envs per GPU, done ignored, re-seeded every {TERRAIN_EPISODE} steps",
        "srb_step_kernel<float, TERRAIN>", N, ms, 1216, "env-steps/s", "weak", "f32")
    env.close()
    # ---- configs[3]: gym_cdm2 deepmimiccdm-v1 (walk2-v1) and its terrain variant (walkterrain-v1), 16384 envs, driven by the reference's
    #      trained policies (evaluated in torch, untimed)
    N = 16384
    tab, skel = hc.load_tables()
    params = {"noise_q": 0.1, "noise_pos": 0.01, "noise_dq": 0.1, "noise_vx": 0.01, "noise_vy": 0.1, "noise_vz": 0.01}

    def cdm_record(name, config, kernel, policy_file, bytes_per_step, **env_kw):
        cenv = CDMVecEnv(tab, skel, N, precision=32, auto_reset=True, seed=3 + rank, params=params, device=local_rank, **env_kw)
        pol = dict(np.load(os.path.join(ROOT, "tests", "golden", policy_file)))
        Wt = [torch.as_tensor(pol[f"W{i}"], dtype=torch.float32, device=dev) for i in range(int(pol["n_layers"]))]
        Bt = [torch.as_tensor(pol[f"b{i}"], dtype=torch.float32, device=dev) for i in range(int(pol["n_layers"]))]
        mean = torch.as_tensor(pol["obs_mean"], dtype=torch.float32, device=dev)
        istd = torch.as_tensor(1.0 / np.sqrt(pol["obs_var"] + 1e-8), dtype=torch.float32, device=dev)
        state = {"obs": cenv.reset(), "act": None}

        def policy(_k):
            h = torch.clamp((state["obs"] - mean) * istd, -10.0, 10.0)
            for w, b in zip(Wt, Bt):
                h = torch.tanh(h @ w.T + b)
            state["act"] = h

        def cdm_step():
            state["obs"] = cenv.step(state["act"])[0]
        for k in range(100):                                   # let the gait reach its steady state
            policy(k); cdm_step()
        ms = _timed(cdm_step, S, Wm, flush, dist, policy)
        rec(name, config, kernel, N, ms, bytes_per_step, "env-steps/s", "weak", "f32")
        cenv.close()
    cdm_record("cdm_fp32", f"BASELINE configs[3]: gym_cdm2 deepmimiccdm-v1 (walk2-v1 constants), {N} envs per GPU, actions from the reference's trained walk2-v1 policy",
               "cdm_step_kernel<float>", "cdm_policy_walk2.npz", 740)
    try:                                                       # the terrain variant is an extra record: never let it take the bench line down
        cdm_record("cdm_terrain_fp32", f"gym_cdm2 deepmimiccdmterrain-v1 (walkterrain-v1): the same step on the gym_cdm2 height field, {N} envs per GPU, actions from the reference's "
                   "trained walkterrain-v1 policy", "cdm_step_kernel<float, TERRAIN>", "cdm_policy_walkterrain.npz", 740 + 16, terrain=hc.load_terrain())
    except Exception as exc:                                   # noqa: BLE001
        print(f"[bench] walkterrain-v1 record skipped: {exc!r}", file=sys.stderr)
    # ---- configs[4]: FK + CoM + centroidal momentum of 1 M poses, split over the GPUs (strong scaling)
    sk = Skeleton(vrml_skeleton.builtin(), device=local_rank)
    n = (1 << 20) // world
    g = torch.Generator(device=dev); g.manual_seed(4 + rank)
    q = torch.randn(n, 4, device=dev, generator=g); q = q / q.norm(dim=1, keepdim=True)
    pose = torch.cat([torch.rand(n, 3, device=dev, generator=g) * 10 - 5, q, torch.rand(n, 53, device=dev, generator=g) * 2 - 1], 1).contiguous()
    dq = (torch.rand(n, 59, device=dev, generator=g) * 2 - 1).contiguous()
    out = sk.forward(pose, dq)
    for mode, d, by in (("fk+com", None, 836 - 24), ("fk+com+momentum", dq, 836)):
        ms = _timed(lambda: sk.forward(pose, d, out=out), 10, 3, flush, dist)
        rec("fk_fp32_" + ("momentum" if d is not None else "com"), f"BASELINE configs[4]: BoneForwardKinematics ({mode}) of 2^20 hanyang poses split over {world} GPU", "fk_hanyang_tma_*_kernel<float>",
            n, ms, by, "poses/s", "strong", "f32")
    sk.close()
    # ---- MMSTO: 4096 windows of 7 frames x 59 dofs per GPU, 30 L-BFGS iterations (online configuration)
    P = dict(np.load(os.path.join(ROOT, "tests", "golden", "mmsto_bench_problems.npz")))
    nw = 4096
    markers = [dict(bone=int(b), lpos=list(map(float, l)), weight=float(w)) for b, l, w in zip(P["marker_bone"], P["marker_lpos"], P["marker_weight"])]
    opt = ShortTrajOpt(vrml_skeleton.builtin(), 7, markers, device=local_rank)
    opt.set_dof_weights(P["dof_weights"]); opt.set("max_iterations", 30)
    tile = lambda a: torch.as_tensor(np.tile(a, (nw // a.shape[0],) + (1,) * (a.ndim - 1)), device=dev)
    keys = ("gpos_target", "x_original", "dx", "ddx", "com", "momentum")
    targs = [tile(P[k]) for k in keys]; em = tile(P["euler_mot"]); vc = tile(P["velcon"])
    info = {}

    def solve():
        info["i"] = opt.solveEndEffectors_euler(em, *targs, velcon=vc)[1]
    ms = _timed(solve, 3, 2, flush, dist)
    evals = float(info["i"][:, 3].sum().item())
    by = (7 * 59 * 2 + 7 * 12 + 6 * 59 + 5 * 59 + 7 * 3 + 6 * 6 + 7 * 59) * 8     # inputs + solution per window (the L-BFGS workspace stays on chip / in L2)
    rec("mmsto_f64", f"ShortTrajOpt.solveEndEffectors (online configuration), {nw} windows per GPU, 30 L-BFGS iterations ({evals / nw:.1f} evaluations per window)", "mmsto_kernel",
        nw, ms, by, "windows/s", "weak", "f64", {"evaluations_per_s": evals * world / (ms * 1e-3), "steps": 3, "warmup": 2})
    opt.close()
    return recs if rank == 0 else []


def env_lanes(args):
    if args.lanes:
        return args.lanes
    # the library's default (srb_create): by batch size
    return 8 if args.envs <= (3072 if args.precision == 64 else 4096) else 4


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and "RANK" not in os.environ:      # convenience: launch the ranks ourselves
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", str(29500 + os.getpid() % 400), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
