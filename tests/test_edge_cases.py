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
