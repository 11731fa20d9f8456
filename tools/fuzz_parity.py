"""Differential sweep: many seeds of the small synthetic configs, CUDA path vs CPU oracle (run on the GPU box).
Test infrastructure (imports the oracle); prints one line per mismatch and a summary."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.pyoracle import OracleSession  # noqa: E402
from volcano_b200 import engine  # noqa: E402
from volcano_b200.synth import make_snapshot  # noqa: E402

cfgs = sys.argv[1].split(",") if len(sys.argv) > 1 else ["tiny", "tiny_fut", "tiny_topo", "small_roles", "small_soft"]
n_seeds = int(sys.argv[2]) if len(sys.argv) > 2 else 20
engine.init(0)
bad = 0
total = 0
for cfg in cfgs:
    for seed in range(100, 100 + n_seeds):
        snap = make_snapshot(cfg, seed)
        r = engine.gpu_engine(snap)
        o = OracleSession(snap, threads=4)
        dec, vis, fe = o.allocate()
        o.close()
        ok = (len(dec) == len(r.decisions) and np.array_equal(dec["task"], r.decisions["task"]) and
              np.array_equal(dec["node"], r.decisions["node"]) and np.array_equal(dec["kind"], r.decisions["kind"]) and
              np.array_equal(dec["score"], r.decisions["score"]) and np.array_equal(vis, r.visits) and np.array_equal(fe, r.fit_errors))
        total += 1
        if not ok:
            bad += 1
            k = next((i for i in range(min(len(dec), len(r.decisions))) if dec[i] != r.decisions[i]), -1)
            print(f"MISMATCH {cfg} seed={seed}: oracle {len(dec)} decisions, gpu {len(r.decisions)}; first difference at {k}")
print(f"{total - bad}/{total} sessions identical (bit-equal scores)")
