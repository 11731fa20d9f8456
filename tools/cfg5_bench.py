"""BASELINE.json configs[4] (32 queues, 10k nodes near capacity, ~10k pending tasks per cycle, actions allocate + preempt +
reclaim): steady-state cycles on one GPU, each cycle a fresh synthetic session of the same distribution (make_cfg5 with
its own seed) through the C ABI — upload (incl. the node.Tasks table), vc_allocate_run, vc_preempt_run, vc_reclaim_run —
and, for the first cycle, the same actions on the CPU oracle with every statement compared.
usage: python tools/cfg5_bench.py [cycles] [n_nodes] [n_pending]   (run on the GPU box)"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from volcano_b200 import engine  # noqa: E402
from volcano_b200.synth import make_cfg5  # noqa: E402

cycles = int(sys.argv[1]) if len(sys.argv) > 1 else 6
n_nodes = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
n_pending = int(sys.argv[3]) if len(sys.argv) > 3 else 10_000
engine.init(0)
rows = []
parity = None
for c in range(cycles):
    t0 = time.perf_counter()
    snap = make_cfg5(seed=1000 + c, n_nodes=n_nodes, n_pending=n_pending)
    t_gen = time.perf_counter() - t0
    e = engine.Engine(snap)
    t0 = time.perf_counter()
    e.upload()
    t1 = time.perf_counter()
    ra = e.allocate()
    t2 = time.perf_counter()
    rp = e.preempt()
    t3 = time.perf_counter()
    rr = e.reclaim()
    t4 = time.perf_counter()
    e.close()
    placed = len(ra.decisions) + int((rp.decisions["kind"] == 1).sum()) + int((rr.decisions["kind"] == 1).sum())
    rows.append({"cycle": c, "pending": snap.T, "running": snap.RT, "jobs": snap.J, "allocated": len(ra.decisions),
                 "preempt_pipelined": int((rp.decisions["kind"] == 1).sum()), "preempt_evicted": int((rp.decisions["kind"] == 2).sum()),
                 "reclaim_pipelined": int((rr.decisions["kind"] == 1).sum()), "reclaim_evicted": int((rr.decisions["kind"] == 2).sum()),
                 "upload_ms": 1e3 * (t1 - t0), "allocate_ms": 1e3 * (t2 - t1), "preempt_ms": 1e3 * (t3 - t2),
                 "reclaim_ms": 1e3 * (t4 - t3), "cycle_ms": 1e3 * (t4 - t0), "placed": placed,
                 "preempt_launches": rp.stats["kernel_launches"], "reclaim_launches": rr.stats["kernel_launches"],
                 "preempt_host_ms": {"in_launch_calls": rp.stats["prof_cycles"][0] / 1e3, "waiting_for_handout": rp.stats["prof_cycles"][1] / 1e3,
                                     "ranked_ahead": rp.stats["prof_cycles"][2], "ranked_ahead_used": rp.stats["prof_cycles"][3]},
                 "reclaim_host_ms": {"in_launch_calls": rr.stats["prof_cycles"][0] / 1e3, "waiting_for_handout": rr.stats["prof_cycles"][1] / 1e3,
                                     "ranked_ahead": rr.stats["prof_cycles"][2], "ranked_ahead_used": rr.stats["prof_cycles"][3]},
                 "generate_s": t_gen})
    if c == 0:
        from oracle.pyoracle import OracleSession
        threads = max(1, min(16, len(os.sched_getaffinity(0))))
        o = OracleSession(snap, threads=threads)
        u0 = time.perf_counter()
        oa = o.allocate()
        u1 = time.perf_counter()
        op = o.preempt()
        u2 = time.perf_counter()
        orr = o.reclaim()
        u3 = time.perf_counter()
        o.close()
        same = (np.array_equal(ra.decisions, oa[0]) and np.array_equal(ra.visits, oa[1]) and
                all(np.array_equal(a.decisions[f], b[0][f]) for a, b in ((rp, op), (rr, orr)) for f in ("task", "node", "kind", "visit")) and
                np.array_equal(rp.visits, op[1]) and np.array_equal(rr.visits, orr[1]))
        parity = {"statements_identical": bool(same), "cpu_threads": threads, "cpu_allocate_ms": 1e3 * (u1 - u0),
                  "cpu_preempt_ms": 1e3 * (u2 - u1), "cpu_reclaim_ms": 1e3 * (u3 - u2), "cpu_cycle_ms": 1e3 * (u3 - u0)}
steady = rows[1:] if len(rows) > 1 else rows
tot_ms = sum(r["cycle_ms"] for r in steady)
out = {"workload": f"cfg5: {n_nodes} nodes, ~{n_pending} pending tasks / cycle, 32 queues, allocate + preempt + reclaim",
       "cycles_timed": len(steady), "pods_per_s": 1e3 * sum(r["placed"] for r in steady) / tot_ms,
       "cycle_ms_p50": float(np.median([r["cycle_ms"] for r in steady])), "parity_cycle0": parity, "cycles": rows}
print(json.dumps(out))
