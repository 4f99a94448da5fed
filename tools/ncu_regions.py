"""Instruction / stall-sample share of the step kernel by source region (function level) from an ncu source-page csv.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv ; python tools/ncu_regions.py src.csv"""
import csv, sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(open(sys.argv[1])))
hdr = None; sec = 0; secs = {}
for r in rows:
    if len(r) > 6 and r[0] == "Line No":
        hdr = r; sec += 1; continue
    if hdr is None or r[0] == "": continue
    try: ln = int(r[0])
    except ValueError: continue
    d = dict(zip(hdr, r))
    try: v = (int(float(d["Instructions Executed"])), int(float(d["# Samples"])))
    except ValueError: v = (0, 0)
    secs.setdefault(sec, {})[ln] = v
src = open(os.path.join(R, "ipcdnnwalk_b200", "csrc", "srb_kernels.cuh")).read().split("\n")
def find(pat):
    for i, l in enumerate(src):
        if pat in l: return i + 1
    raise KeyError(pat)
marks = [("load/store_state", "void load_state"), ("terrain..reset", "struct TerrainHit"), ("write_obs", "// observation (SRBEnv.get_obs_window"), ("gsum/HReduce", "// group helpers"),
         ("ldl", "// 6x6 SPD solve"), ("QP.build", "void build(int lane"), ("QP.tvals/phi", "static constexpr bool SHARED_LIN"), ("QP.hessian", "// H (lower triangle, 21)"),
         ("QP.solve", "// returns number of Newton systems solved"), ("QP.solve_ls", "// Same Newton iteration, globalised"), ("sds", "// SDS swing-foot LQR filter"),
         ("reward", "// reward (SRBEnv._get_reward"), ("done", "// SRBEnv._get_done"), ("k:prologue", "// the step kernel"), ("k:feet", "// ---- feet (:320"),
         ("k:substeps-preQP", "// ---- 4 physics sub-steps"), ("k:QPcall+clamp", "QP<real, G> qp;"), ("k:integrate", "// mj_step2: acceleration"),
         ("k:epilogue", "// ---- history push"), ("k:reset+store", "// ---- fused auto-reset")]
pos = sorted((find(p), n) for n, p in marks)
half = max(secs) // 2 if max(secs) % 2 == 0 else max(secs)          # two launches are captured: use the first
S = secs[1]
ti = sum(sum(x[0] for x in v.values()) for k, v in secs.items() if k <= half); ts = sum(sum(x[1] for x in v.values()) for k, v in secs.items() if k <= half)
print(f"warp instructions per launch {ti}  ({ti / 1024:.0f} per warp at 1024 warps)")
for (a, n), (b, _) in zip(pos, pos[1:] + [(10 ** 7, "")]):
    i = sum(v[0] for l, v in S.items() if a <= l < b); s = sum(v[1] for l, v in S.items() if a <= l < b)
    print(f"{n:20s} {100 * i / ti:5.1f}% inst {100 * s / ts:5.1f}% smp")
if 2 in secs: print(f"{'srb_math.cuh':20s} {100 * sum(x[0] for x in secs[2].values()) / ti:5.1f}% inst {100 * sum(x[1] for x in secs[2].values()) / ts:5.1f}% smp")
