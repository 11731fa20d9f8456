"""One upload + one allocate on a synthetic config (target for ncu captures)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from volcano_b200 import engine  # noqa: E402
from volcano_b200.synth import make_snapshot  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
mode = sys.argv[2] if len(sys.argv) > 2 else "allocate"
snap = make_snapshot(cfg)
if os.environ.get("SAMP"):  # feasible-node sampling (0 = the reference's adaptive default)
    snap.conf.percentage_nodes_to_find = int(os.environ["SAMP"]) if os.environ["SAMP"].isdigit() else 0
e = engine.Engine(snap)
e.upload()
if mode == "dense":
    print(e.score_matrix_device(repeats=3))
else:
    r = e.allocate()
    print(len(r.decisions), r.stats["commit_ms"])
e.close()
