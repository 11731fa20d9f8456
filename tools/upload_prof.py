import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from volcano_b200 import engine
from volcano_b200.synth import make_snapshot
snap = make_snapshot(sys.argv[1] if len(sys.argv) > 1 else "cfg2")
e = engine.Engine(snap); e.upload(); e.upload()
os.environ["VC_PROF_UPLOAD"] = "1"
t = time.perf_counter(); e.upload(); print("upload wall ms", 1e3 * (time.perf_counter() - t))
e.close()
