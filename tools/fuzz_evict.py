"""Random API-level clusters for the preempt and reclaim actions: nodes filled with running pods of several jobs and
queues (preemptable / not, task and job priorities, critical pods, Bound pods), starving jobs with pending pods, queue
capabilities / priorities / closed and non-reclaimable queues, random subsets and tier splits of the plugins that vote
on victims (conformance, gang, priority, drf, proportion) plus the scoring plugins, and a random action list out of
allocate / preempt / reclaim.  make_case(seed) -> (TestCommonStruct, tiers, actions); used by tests/test_gpu_fuzz.py and
`python tools/fuzz_evict.py <first> <last>` on a GPU box (CUDA path vs CPU oracle, statement by statement)."""
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from volcano_b200.api import (BuildNode, BuildPod, BuildPodGroup, BuildQueue, BuildResourceList)  # noqa: E402
from volcano_b200.snapshot import PluginOption  # noqa: E402
from volcano_b200.uthelper import TestCommonStruct  # noqa: E402


def make_case(seed):
    rnd = random.Random(seed * 104729 + 7)
    n_nodes = rnd.randint(1, 6)
    nodes = []
    for i in range(n_nodes):
        labels = {"zone": rnd.choice(["a", "b"])} if rnd.random() < 0.4 else {}
        nodes.append(BuildNode(f"n{i}", BuildResourceList(str(rnd.choice([2, 3, 4])), rnd.choice(["4G", "8G", "16G"]),
                                                          ("pods", str(rnd.choice([3, 6, 20])))), labels))
    n_queues = rnd.randint(1, 3)
    queues = []
    for q in range(n_queues):
        cap = BuildResourceList(str(rnd.choice([2, 4, 8, 16])), rnd.choice(["4G", "8G", "32G"])) if rnd.random() < 0.4 else None
        qu = BuildQueue(f"q{q}", rnd.randint(1, 4), cap)
        qu.priority = rnd.choice([0, 0, 1, 5])
        if rnd.random() < 0.1:
            qu.state = "Closed"
        if rnd.random() < 0.15:
            qu.reclaimable = False
        queues.append(qu)
    pgs, pods = [], []
    n_jobs = rnd.randint(2, 7)
    for j in range(n_jobs):
        q = rnd.randrange(n_queues)
        n_run, n_pend = rnd.randint(0, 5), rnd.randint(0, 4)
        size = n_run + n_pend
        if size == 0:
            n_pend = size = 1
        min_member = rnd.choice([0, 1, max(1, size // 2), size, size + 1])
        roles = rnd.random() < 0.3
        tmm = None
        if roles and rnd.random() < 0.5:
            tmm = {"w": rnd.randint(0, max(1, size // 2))}
        phase = rnd.choice(["Inqueue", "Running", "Running", "Pending"])
        pg = BuildPodGroup(f"pg{j}", "ns", f"q{q}", min_member, tmm, phase)
        pg.priority = rnd.choice([0, 10, 10, 1000])
        pgs.append(pg)
        req = BuildResourceList(rnd.choice(["500m", "1", "2", "3"]), rnd.choice(["1G", "2G", "4G"]))
        sel = {"zone": rnd.choice(["a", "b"])} if rnd.random() < 0.15 else {}
        for k in range(size):
            running = k < n_run
            role = ("m" if k == 0 else "w") if roles else "w"
            labels = {"volcano.sh/task-spec": role}
            r = rnd.random()
            if r < 0.35:
                labels["volcano.sh/preemptable"] = "true"
            elif r < 0.55:
                labels["volcano.sh/preemptable"] = "false"
            p = BuildPod("ns" if rnd.random() > 0.05 else "kube-system", f"j{j}-{role}-{k}", rnd.choice(nodes).name if running else "",
                         "Running" if running and rnd.random() > 0.1 else "Pending", req, f"pg{j}", labels, {} if running else sel)
            if p.namespace != "ns":
                p.group_name = f"pg{j}"
            if rnd.random() < 0.3:
                p.priority = rnd.choice([1, 5, 100])
            if not running and rnd.random() < 0.08:
                p.preemption_policy = "Never"
            if running and rnd.random() < 0.05:
                p.priority_class_name = "system-node-critical"
            p.creation_ts = rnd.randint(0, 3)
            pods.append(p)
    # pods of a kube-system namespace belong to a job id of their own namespace: keep them in "ns" groups only
    for p in pods:
        if p.namespace != "ns":
            p.namespace = "ns"
            p.priority_class_name = "system-cluster-critical"
            p.uid = ""
            p.__post_init__()
    voters = [n for n in ("conformance", "gang", "priority", "drf", "proportion") if rnd.random() < 0.75]
    scorers = [n for n in ("predicates", "nodeorder", "binpack") if rnd.random() < 0.5]
    chosen = voters + scorers
    rnd.shuffle(chosen)
    if not chosen:
        chosen = ["gang"]
    split = rnd.randint(0, len(chosen))
    args = {"binpack": {"binpack.weight": rnd.choice([1, 5]), "binpack.cpu": rnd.choice([1, 3])}}
    tiers = [t for t in ([PluginOption.defaults(n, args.get(n)) for n in chosen[:split]],
                         [PluginOption.defaults(n, args.get(n)) for n in chosen[split:]]) if t]
    actions = tuple(a for a in ("allocate", "preempt", "reclaim") if rnd.random() < 0.7) or ("preempt",)
    if rnd.random() < 0.3:
        actions = ("enqueue",) + actions
    if rnd.random() < 0.2 and "preempt" in actions and "reclaim" in actions:
        actions = tuple(a for a in actions if a not in ("preempt", "reclaim")) + ("reclaim", "preempt")
    tc = TestCommonStruct(Name=f"evict-fuzz{seed}", Nodes=nodes, Pods=pods, PodGroups=pgs, Queues=queues)
    return tc, tiers, actions


def compare(res, ref, seed):
    import numpy as np
    for name in ("preempt", "reclaim"):
        a, b = getattr(res, name), getattr(ref, name)
        assert (a is None) == (b is None), (seed, name)
        if a is None:
            continue
        assert np.array_equal(a.visits, b.visits), (seed, name, a.visits, b.visits)
        for f in ("task", "node", "kind", "visit"):
            assert np.array_equal(a.decisions[f], b.decisions[f]), (seed, name, f, a.decisions, b.decisions)
    assert np.array_equal(res.decisions, ref.decisions), seed
    assert np.array_equal(res.visits, ref.visits), seed
    return sum(len(getattr(ref, n).decisions) for n in ("preempt", "reclaim") if getattr(ref, n) is not None)


if __name__ == "__main__":
    import numpy as np
    from oracle.pyoracle import OracleSession
    from volcano_b200 import engine
    from volcano_b200.uthelper import AllocateResult
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
    first, last = int(sys.argv[1]), int(sys.argv[2])
    engine.init(0)
    import conftest  # the oracle engine of the test suite
    oracle = conftest.oracle_engine.__wrapped__() if hasattr(conftest.oracle_engine, "__wrapped__") else None
    ops = evictions = 0
    for seed in range(first, last):
        tc, tiers, actions = make_case(seed)
        snap = tc.RegisterSession(tiers, actions=actions)
        if snap.T == 0 or snap.N == 0 or snap.B > 0:
            continue
        ref = oracle(snap)
        res = engine.gpu_engine(snap)
        ops += compare(res, ref, seed)
    print(f"seeds [{first},{last}): {ops} evict/pipeline operations identical")
