"""Phase timers of k_backfill (VC_PROF=1): cycles of CTA 0 per task in record+sweep / CTA fold / mailbox / apply."""
import os
import sys
os.environ["VC_PROF"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from volcano_b200 import engine  # noqa: E402
from volcano_b200.synth import make_snapshot  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2_bf"
snap = make_snapshot(cfg)
engine.init(0)
e = engine.Engine(snap)
e.upload()
for _ in range(2):
    e.allocate()
    r = e.backfill()
pc = r.stats["prof_cycles"][:4]
n = max(1, r.stats["n_steps"])
print(f"{cfg}: {r.stats['commit_ms']:.2f} ms, {n} tasks; cycles/task: sweep {pc[0]/n:.0f}, fold barrier {pc[1]/n:.0f}, "
      f"mailbox {pc[2]/n:.0f}, apply+top barrier {pc[3]/n:.0f}")
e.close()
