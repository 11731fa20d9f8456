import importlib.util, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
spec = importlib.util.spec_from_file_location("fz", os.path.join(ROOT, "tools", "fuzz_api.py")); fz = importlib.util.module_from_spec(spec); spec.loader.exec_module(fz)
from oracle.pyoracle import OracleSession
from volcano_b200 import engine
seed = int(sys.argv[1])
tc, tiers, actions = fz.make_case(seed)
snap = tc.RegisterSession(tiers, actions=actions, **tc.conf_kw)
print("N", snap.N, "T", snap.T, "J", snap.J, "Q", snap.Q, "actions", actions)
print("tiers", [[(p.name, hex(p.enabled)) for p in t] for t in tiers])
print("soft", None if snap.hn_job_soft is None else snap.hn_job_soft.tolist(), "alloc0", None if snap.hn_job_allocated is None else snap.hn_job_allocated.tolist())
o = OracleSession(snap); dec, vis, fe = o.allocate(); o.close()
print("ORACLE visits", vis.tolist()); print("ORACLE dec", [(int(d['task']), int(d['node']), int(d['kind']), float(d['score'])) for d in dec]); print("fit", fe.tolist())
if len(sys.argv) > 2 and sys.argv[2] == "cpu": sys.exit(0)
engine.init(0)
r = engine.gpu_engine(snap)
print("GPU visits", r.visits.tolist()); print("GPU dec", [(int(d['task']), int(d['node']), int(d['kind']), float(d['score'])) for d in r.decisions]); print("fit", r.fit_errors.tolist())
print("stats", {k: r.stats[k] for k in ("n_steps",)})
