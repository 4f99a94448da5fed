"""tests/golden/mmsto_bench_problems.npz: 16 seeded MMSTO windows (inputs only: desired motion, marker / CoM / momentum targets, contact
flags, velocity half-spaces) in the shapes walk_SRBTrack2025.py:505-587 builds, so that bench.py can time the batched solve without
importing the oracle (tests/helpers_mmsto.make_problem builds the targets with oracle/mmsto.EulerTree).
    python tests/tools/make_golden_mmsto_bench.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers_mmsto as hm          # noqa: E402

KEYS = ("gpos_target", "x_original", "dx", "ddx", "com", "momentum", "euler_mot", "velcon", "con")
ps = [hm.make_problem(s) for s in range(16)]
out = {k: hm.stack(ps, k) for k in KEYS}
out["dof_weights"] = hm.ankle_dof_weights()
out["marker_bone"] = np.array([m["bone"] for m in hm.MARKERS]); out["marker_lpos"] = np.array([m["lpos"] for m in hm.MARKERS])
out["marker_weight"] = np.array([m["weight"] for m in hm.MARKERS])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "mmsto_bench_problems.npz"), **out)
print({k: np.asarray(v).shape for k, v in out.items()})
