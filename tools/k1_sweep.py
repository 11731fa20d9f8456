"""K1 (dense task x node mask + score pass) against the number of distinct (class, request) groups G: the r01 figure
(0.98 of the HBM copy peak) rests on G = 368 groups at cfg2 — tasks of one pod template share a row, so 10^9 pairs
collapse to G x N evaluations and the pass is a broadcast store. With heterogeneous pods G grows towards T and the
evaluation kernel (K1a, fp64) takes over. Prints one JSON line per G: group count, ms of the whole pass and of the
materialising kernel alone, achieved GB/s on the algorithmic bytes of SURVEY 8(d).
usage: python tools/k1_sweep.py        (run on the GPU box)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from volcano_b200 import engine  # noqa: E402
from volcano_b200.synth import CONFIGS, SynthConfig, make_snapshot  # noqa: E402

engine.init(0)
base = CONFIGS["cfg2"]
peak = 6578.7
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = float(json.load(open(p))["hbm_gbs"])
for hetero in (0, 4, 16, 64, 320):
    cfg = SynthConfig("cfg2_h%d" % hetero, base.n_nodes, base.n_tasks, base.n_queues, base.plugins, hetero=hetero)
    snap = make_snapshot(cfg)
    e = engine.Engine(snap)
    e.upload()
    dense_ms, expand_ms, nbytes = e.score_matrix_device(repeats=3)
    e.close()
    import numpy as np
    key = np.concatenate([snap.t_klass[None, :].astype(np.float64), snap.t_resreq], axis=0).T
    G = len(np.unique(key, axis=0))
    print(json.dumps({"hetero": hetero, "groups": int(G), "tasks": snap.T, "nodes": snap.N, "dense_pass_ms": dense_ms,
                      "expand_kernel_ms": expand_ms, "eval_and_rest_ms": dense_ms - expand_ms, "algorithmic_bytes": nbytes,
                      "whole_pass_gbs": nbytes / dense_ms / 1e6, "frac_of_copy_peak_whole_pass": nbytes / dense_ms / 1e6 / peak,
                      "frac_of_copy_peak_expand_only": nbytes / expand_ms / 1e6 / peak}), flush=True)
