"""ONE allocate session across the GPUs of a node, one process per GPU over torch.distributed: the plumbing
around vc_comm_create / vc_comm_attach / vc_comm_prepare (include/vcalloc.h). torch.distributed only moves the 64-byte CUDA
IPC handles and provides the barriers; the per-step exchange of the commit loop is device-initiated (peer-mapped mailbox
and publication ring written with NVLink stores from inside the persistent kernel)."""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import engine
from .snapshot import Snapshot


class MultiGpuSession:
    def __init__(self, snap: Snapshot, device: int):
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.eng = engine.Engine(snap, device)
        mine = self.eng.comm_create(self.world, self.rank)
        t = torch.tensor(list(mine), dtype=torch.uint8, device=f"cuda:{device}")
        out = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(out, t)
        self.eng.comm_attach(b"".join(bytes(x.cpu().tolist()) for x in out))
        self.eng.upload()

    def upload(self):
        self.eng.upload()

    def allocate(self):
        """Every rank calls this; rank 0's result carries the decisions."""
        self.eng.comm_prepare()
        torch.cuda.synchronize()
        dist.barrier()  # no rank writes into a slab that is still being cleared
        res = self.eng.allocate()
        dist.barrier()  # every rank is out of the kernel before the next prepare
        return res

    def close(self):
        self.eng.close()
