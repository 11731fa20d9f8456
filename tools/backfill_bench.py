"""Measure the backfill action (vc_backfill_run) on a synthetic session next to the CPU oracle: pods/s of the device
kernel (CUDA events), end to end through the C ABI (host buffers, order + H2D + kernel + D2H), and the oracle on the
host cores.  Test infrastructure (imports the oracle as the checker / CPU baseline).
Usage: python tools/backfill_bench.py [config=cfg2_bf] [repeats=3] [oracle_threads=16]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.pyoracle import OracleSession  # noqa: E402
from volcano_b200 import engine  # noqa: E402
from volcano_b200.synth import make_snapshot  # noqa: E402


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2_bf"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    threads = int(sys.argv[3]) if len(sys.argv) > 3 else min(16, os.cpu_count() or 1)
    snap = make_snapshot(cfg)
    engine.init(0)
    e = engine.Engine(snap)
    e.upload()
    kms, e2e = [], []
    res = None
    for _ in range(reps + 1):  # first round warms up
        e.allocate()
        t0 = time.perf_counter()
        res = e.backfill()
        e2e.append((time.perf_counter() - t0) * 1e3)
        kms.append(res.stats["commit_ms"])
    e.close()
    kms, e2e = kms[1:], e2e[1:]
    placed = len(res.decisions)
    o = OracleSession(snap, threads=threads)
    o.allocate()
    t0 = time.perf_counter()
    dec, vis, fe = o.backfill()
    cpu_ms = (time.perf_counter() - t0) * 1e3
    o.close()
    same = bool(np.array_equal(dec, res.decisions) and np.array_equal(vis, res.visits) and np.array_equal(fe, res.fit_errors))
    out = {
        "workload": f"{cfg}: {snap.N} nodes, {snap.T} regular + {snap.B} BestEffort pending pods, allocate then backfill",
        "backfill_tasks": snap.B, "placed": placed, "fit_errors": int(len(res.fit_errors)), "visits": int(len(res.visits)),
        "kernel_ms_median": float(np.median(kms)), "e2e_ms_median": float(np.median(e2e)),
        "pods_per_s_kernel": snap.B / (float(np.median(kms)) / 1e3), "pods_per_s_e2e": snap.B / (float(np.median(e2e)) / 1e3),
        "cpu_oracle": {"ms": cpu_ms, "pods_per_s": snap.B / (cpu_ms / 1e3), "threads": threads, "kind": "port"},
        "identical_to_oracle": same,
    }
    print(json.dumps(out))


if __name__ == "__main__":
    main()
