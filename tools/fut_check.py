"""One synthetic config through the CUDA path and the oracle; prints the kernel that served it, times and the first
difference. `python tools/fut_check.py mid_fut [threads]` (GPU box)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from volcano_b200 import engine  # noqa: E402
from volcano_b200.synth import make_snapshot  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "mid_fut"
    threads = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    engine.init(0)
    snap = make_snapshot(name)
    if os.environ.get("SAMP"):  # the reference's default adaptive feasible-node sampling, single-worker reading
        snap.conf.percentage_nodes_to_find = int(os.environ["SAMP"]) if os.environ["SAMP"].isdigit() else 0
        threads = 1
    t0 = time.time()
    res = engine.gpu_engine(snap)
    print(f"{name}: gpu {res.stats['commit_ms']:.1f} ms (kernel {res.stats['commit_kernel']}), {len(res.decisions)} ops, "
          f"{int((res.decisions['kind'] == 1).sum())} pipelined, wall {time.time() - t0:.1f} s", flush=True)
    from oracle.pyoracle import OracleSession
    o = OracleSession(snap, threads=threads)
    t0 = time.time()
    rdec, rvis = o.allocate()[:2]
    o.close()
    print(f"oracle {time.time() - t0:.1f} s, {len(rdec)} ops")
    n = min(len(res.decisions), len(rdec))
    same = all(np.array_equal(res.decisions[k][:n], rdec[k][:n]) for k in ("task", "node", "kind", "visit", "score"))
    if same and len(res.decisions) == len(rdec) and np.array_equal(res.visits, rvis):
        print("identical")
        return 0
    for i in range(n):
        if res.decisions[i] != rdec[i]:
            print("first difference at", i, res.decisions[i], rdec[i])
            break
    return 1


if __name__ == "__main__":
    sys.exit(main())
