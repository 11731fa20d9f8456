"""Edge cases through the C ABI on the GPU: empty and degenerate sessions, against the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run_both(snap, oracle_engine):
    from volcano_b200 import engine
    res = engine.gpu_engine(snap)
    ref = oracle_engine(snap)
    assert np.array_equal(res.decisions[["task", "node", "kind", "visit"]], ref.decisions[["task", "node", "kind", "visit"]])
    assert np.array_equal(res.visits, ref.visits)
    assert np.array_equal(res.fit_errors, ref.fit_errors)
    return res


def _cluster(n_nodes, pods_per_job, jobs, cpu="4", mem="8Gi", req=("1", "1Gi"), min_member=None, node_labels=None, selector=None):
    from volcano_b200.api import BuildNode, BuildPod, BuildPodGroup, BuildQueue, BuildResourceList
    from tests.golden.reference_cases import allocate_tiers
    from volcano_b200.uthelper import TestCommonStruct
    nodes = [BuildNode(f"n{i}", BuildResourceList(cpu, mem, ("pods", "10")), node_labels or {}) for i in range(n_nodes)]
    pgs, pods = [], []
    for j in range(jobs):
        pgs.append(BuildPodGroup(f"pg{j}", "ns", "q", pods_per_job if min_member is None else min_member, None, "Inqueue"))
        pods += [BuildPod("ns", f"pg{j}-w-{k}", "", "Pending", BuildResourceList(*req), f"pg{j}", {}, selector or {})
                 for k in range(pods_per_job)]
    tc = TestCommonStruct(Name="edge", Pods=pods, Nodes=nodes, PodGroups=pgs, Queues=[BuildQueue("q", 1, None)])
    return tc.RegisterSession(allocate_tiers())


def test_no_tasks(oracle_engine):
    res = _run_both(_cluster(3, 0, 2), oracle_engine)
    assert len(res.decisions) == 0 and len(res.visits) == 0


def test_no_nodes(oracle_engine):
    res = _run_both(_cluster(0, 2, 2), oracle_engine)
    assert len(res.decisions) == 0


def test_no_jobs(oracle_engine):
    res = _run_both(_cluster(2, 0, 0), oracle_engine)
    assert len(res.decisions) == 0


def test_single_node_fills_up(oracle_engine):
    res = _run_both(_cluster(1, 6, 1, min_member=1), oracle_engine)  # 4 cpus: 4 of 6 pods fit
    assert len(res.decisions) == 4


def test_gang_cannot_be_satisfied_is_discarded(oracle_engine):
    res = _run_both(_cluster(1, 6, 1), oracle_engine)  # minMember 6 on a 4-cpu node
    assert len(res.decisions) == 0 and (res.visits["outcome"] == 2).all()


def test_selector_matches_nothing(oracle_engine):
    res = _run_both(_cluster(2, 2, 2, node_labels={"a": "b"}, selector={"a": "c"}), oracle_engine)
    assert len(res.decisions) == 0 and len(res.fit_errors) > 0


def test_many_nodes_one_task(oracle_engine):
    res = _run_both(_cluster(300, 1, 1), oracle_engine)
    assert len(res.decisions) == 1


def test_repeated_sessions_on_one_engine(oracle_engine):
    """vc_snapshot_upload / vc_allocate_run are re-entrant on one snapshot handle: same shape, new contents."""
    from volcano_b200 import engine
    from volcano_b200.synth import make_snapshot
    a, b = make_snapshot("tiny", 11), make_snapshot("tiny", 12)
    if (a.J, a.NR) != (b.J, b.NR):
        pytest.skip("shapes differ")
    e = engine.Engine(a)
    try:
        for snap in (a, b, a):
            e.upload(snap)
            res = e.allocate()
            ref = oracle_engine(snap)
            assert np.array_equal(res.decisions["node"], ref.decisions["node"])
    finally:
        e.close()


def test_unsupported_features_fail_loudly():
    from volcano_b200 import abi, engine
    from volcano_b200.synth import make_snapshot
    snap = make_snapshot("tiny", 1)
    snap.j_flags[0] |= abi.VC_JOB_UNSUPPORTED
    with pytest.raises(engine.VcError) as ei:
        engine.gpu_engine(snap)
    assert ei.value.code == abi.VC_EUNSUPPORTED


def test_incremental_node_upload(oracle_engine):
    """vc_snapshot_update_nodes: the accounting rows of the dirty nodes alone are uploaded; the next cycle must be the one a
    full upload of the moved cluster gives (and the oracle's)."""
    from volcano_b200 import engine
    from volcano_b200.synth import make_snapshot
    snap = make_snapshot("small", 11)
    e = engine.Engine(snap)
    e.upload()
    first = e.allocate()
    # the cluster moved: pods finished on 5 % of the nodes (Used -> Idle), pods started on another 5 %
    rng = np.random.default_rng(5)
    dirty = rng.choice(snap.N, size=max(2, snap.N // 10), replace=False).astype(np.int32)
    half = len(dirty) // 2
    unit = np.ones(snap.R)  # quantities stay whole: extended resources move in units of 1000 milli (one device)
    unit[snap.dim_names.index("nvidia.com/gpu")] = 1000.0
    for n in dirty[:half]:
        freed = np.floor(snap.n_used[:, n] / 2 / unit) * unit
        snap.n_used[:, n] -= freed
        snap.n_idle[:, n] += freed
    for n in dirty[half:]:
        taken = np.floor(snap.n_idle[:, n] / 3 / unit) * unit
        snap.n_used[:, n] += taken
        snap.n_idle[:, n] -= taken
    snap.n_pod_count[dirty] = snap.n_used[snap.pods_dim, dirty].astype(np.int32)
    kd = [snap.dim_names.index(k) for k in ("cpu", "memory", "nvidia.com/gpu")]
    for k, d in enumerate(kd):
        scale = 1000.0 if k == 2 else 1.0
        snap.n_k8s_requested[k, dirty] = snap.n_used[d, dirty] / scale
        snap.n_k8s_nonzero_requested[k, dirty] = snap.n_used[d, dirty] / scale
    e.update_nodes(dirty, snap.n_idle[:, dirty], snap.n_used[:, dirty], snap.n_releasing[:, dirty], snap.n_pipelined[:, dirty],
                   snap.n_k8s_requested[:, dirty], snap.n_k8s_nonzero_requested[:, dirty], snap.n_pod_count[dirty])
    inc = e.allocate()
    e.close()
    full = engine.gpu_engine(snap)
    ref = oracle_engine(snap)
    assert np.array_equal(full.decisions, ref.decisions) and np.array_equal(full.visits, ref.visits), "full upload vs oracle"
    assert np.array_equal(inc.decisions, full.decisions) and np.array_equal(inc.visits, full.visits), "incremental vs full upload"
    assert np.array_equal(inc.fit_errors, ref.fit_errors) and np.array_equal(full.fit_errors, ref.fit_errors)
    assert not np.array_equal(first.decisions["node"][:200], inc.decisions["node"][:200]) or len(first.decisions) != len(inc.decisions)
    assert inc.stats["h2d_bytes"] < full.stats["h2d_bytes"] / 5
