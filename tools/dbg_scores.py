import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from volcano_b200 import engine
from volcano_b200.synth import make_snapshot
from oracle.pyoracle import OracleSession
snap = make_snapshot(sys.argv[1], int(sys.argv[2]))
r = engine.gpu_engine(snap)
o = OracleSession(snap); dec, vis, fe = o.allocate()
bad = np.nonzero(np.abs(r.decisions["score"] - dec["score"]) > 1e-9)[0]
print("n", len(dec), "bad", len(bad))
for i in bad[:12]:
    t = dec["task"][i]
    print(i, "task", t, "node", dec["node"][i], "gpu", r.decisions["score"][i], "oracle", dec["score"][i], "diff", r.decisions["score"][i] - dec["score"][i],
          "req", snap.t_resreq[:, t][[0, 1, 5]], "class", snap.t_klass[t])
