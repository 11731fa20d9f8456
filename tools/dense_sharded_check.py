"""torchrun --nproc-per-node N tools/dense_sharded_check.py [cfg]: node-sharded dense pass over NCCL vs the
single-shard pass — per-task best (score, node) must be identical. Prints one JSON line from rank 0."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from volcano_b200 import engine  # noqa: E402
from volcano_b200.parallel import shard_bounds, sharded_dense_best  # noqa: E402
from volcano_b200.synth import make_snapshot  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
snap = make_snapshot(cfg)
eng = engine.Engine(snap, device=local)
eng.upload()
dev = torch.device("cuda", local)
# reference: full node axis on every rank
eng.set_shard(0, snap.N)
_, _, bs_full, bn_full = eng.score_matrix(want_mask=False, want_score=False)
# sharded
times = []
for it in range(5):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    s, n = sharded_dense_best(eng, world, rank, dev, materialize=False)
    torch.cuda.synchronize(); dist.barrier(); times.append(time.perf_counter() - t0)
ok = bool(np.array_equal(n.cpu().numpy(), bn_full) and np.array_equal(s.cpu().numpy(), bs_full))
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"cfg": cfg, "world": world, "shard": shard_bounds(snap.N, world, rank), "identical": bool(flag.item()),
                      "sharded_best_ms": 1e3 * min(times), "tasks": snap.T, "nodes": snap.N}))
eng.close()
dist.destroy_process_group()
