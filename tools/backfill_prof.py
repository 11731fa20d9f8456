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
    r = e.backfill()  # libvcalloc prints two more sums on stderr: divide by the bystander / republisher step counts below
pc = r.stats["prof_cycles"][:8]
n = max(1, r.stats["n_steps"])
n_full, n_own = max(1, pc[1]), max(1, pc[3])
n_other = max(1, n - pc[1] - pc[3])
print(f"{cfg}: {r.stats['commit_ms']:.2f} ms, {n} tasks ({pc[1]} gather steps, CTA 0 republished in {pc[3]}); cycles per step of its kind: "
      f"gather sweep {pc[0]/n_full:.0f}, all-gather {pc[4]/n_full:.0f}; republisher sweep {pc[2]/n_own:.0f}, fold+publish {pc[6]/n_own:.0f}; "
      f"bystander poll {pc[5]/n_other:.0f} ({n_other} steps); apply+barriers per task {pc[7]/n:.0f}")
e.close()
