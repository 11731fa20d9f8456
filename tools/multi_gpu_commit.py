"""torchrun --nproc-per-node W tools/multi_gpu_commit.py [cfg] [cycles]: ONE allocate session across W GPUs (node axis cut
over the CTAs of all ranks, per-step exchange through peer-mapped memory) against the same session on one GPU: identical
decisions, ms per cycle (max over ranks, CUDA events around each rank's kernel), us per step."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from volcano_b200 import engine  # noqa: E402
from volcano_b200.parallel_commit import MultiGpuSession  # noqa: E402
from volcano_b200.synth import make_snapshot  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
cycles = int(sys.argv[2]) if len(sys.argv) > 2 else 3
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
snap = make_snapshot(cfg)
single = None
if rank == 0:  # the same session on one GPU
    e1 = engine.Engine(snap, local)
    e1.upload()
    e1.allocate()
    single = e1.allocate()
    e1.close()
dist.barrier()
ms = MultiGpuSession(snap, local)
times = []
res = None
for c in range(cycles + 1):
    res = ms.allocate()
    t = torch.tensor([res.stats["commit_ms"]], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if c > 0:
        times.append(float(t.item()))
if rank == 0:
    same = (np.array_equal(res.decisions, single.decisions) and np.array_equal(res.visits, single.visits) and
            np.array_equal(res.fit_errors, single.fit_errors))
    out = {"cfg": cfg, "world": world, "identical_to_one_gpu": bool(same), "placed": int(len(res.decisions)),
           "ms_per_cycle": float(np.median(times)), "one_gpu_ms": single.stats["commit_ms"],
           "steps": int(single.stats["n_steps"]), "us_per_step": 1e3 * float(np.median(times)) / max(1, single.stats["n_steps"]),
           "pods_per_s": len(res.decisions) / (float(np.median(times)) * 1e-3)}
    print(json.dumps(out), flush=True)
    if not same:
        raise SystemExit("decisions differ from the one-GPU session")
ms.close()
dist.destroy_process_group()
