"""world_size-2 gloo tests (CPU) of the node-sharded dense pass plumbing: shard bounds and the cross-shard
fold of per-task best (score, node) pairs reproduce the single-shard arg-max of the oracle."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, score, mask, want_score, want_node, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from volcano_b200.parallel import fold_best, shard_bounds
    T, N = score.shape
    b, e = shard_bounds(N, world, rank)
    bs = np.zeros(T)
    bn = np.full(T, -1, np.int32)
    for t in range(T):
        cand = [n for n in range(b, e) if (mask[t, n // 64] >> np.uint64(n % 64)) & np.uint64(1) and cand_ok[t, n]]
        if cand:
            sc = score[t, cand]
            k = int(np.argmax(sc))  # first max = lowest index
            bs[t], bn[t] = sc[k], cand[k]
    s, n = fold_best(torch.from_numpy(bs), torch.from_numpy(bn))
    ok = bool(np.array_equal(n.numpy(), want_node) and np.array_equal(s.numpy(), want_score))
    q.put((rank, ok))
    dist.destroy_process_group()


cand_ok = None


def test_shard_bounds():
    from volcano_b200.parallel import shard_bounds
    for n, w in [(10000, 8), (700, 2), (64, 4), (128, 3), (50000, 8)]:
        blocks = [shard_bounds(n, w, r) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        for (b0, e0), (b1, e1) in zip(blocks, blocks[1:]):
            assert e0 == b1 and b1 % 64 == 0 or b1 == n


def test_fold_best_two_ranks():
    global cand_ok
    from oracle.pyoracle import OracleSession
    from volcano_b200.synth import make_snapshot
    snap = make_snapshot("tiny", 4)
    o = OracleSession(snap)
    mask, score, bs, bn = o.score_matrix()
    o.close()
    # candidates = feasible nodes of the chosen gradient; without Releasing resources that is every feasible node
    cand_ok = np.ones_like(score, dtype=bool)
    world = 2
    ctx = mp.get_context("fork")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, score, mask, bs, bn, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
