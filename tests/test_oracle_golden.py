"""Pin the CPU oracle against the reference's own golden vectors (tests/golden/reference_cases.py)."""
import ctypes as C

import numpy as np
import pytest

from oracle import pyoracle
from oracle.pyoracle import OracleSession
from tests.golden import reference_cases as G
from volcano_b200.snapshot import PluginOption, SchedulerConf, build_conf, encode_cluster
from volcano_b200.uthelper import TestCommonStruct


@pytest.mark.parametrize("case", G.allocate_cases(), ids=lambda c: c.Name[:40])
def test_allocate_cases(case, oracle_engine):
    case.RegisterSession(G.allocate_tiers())
    case.Run(oracle_engine)
    assert case.CheckAll() is None, case.CheckAll()


@pytest.mark.parametrize("case", G.fareshare_cases(), ids=lambda c: c.Name[:40])
def test_fareshare_cases(case, oracle_engine):
    case.RegisterSession(G.fareshare_tiers())
    case.Run(oracle_engine)
    assert case.CheckAll() is None, case.CheckAll()


@pytest.mark.parametrize("case,weights", G.nodeorder_cases(), ids=lambda c: getattr(c, "Name", "w")[:40])
def test_nodeorder_cases(case, weights, oracle_engine):
    case.RegisterSession(G.nodeorder_tiers(**weights))
    case.Run(oracle_engine)
    assert case.CheckAll() is None, case.CheckAll()


@pytest.mark.parametrize("args,expected", G.BINPACK_CASES)
def test_binpack_scores(args, expected):
    tc = TestCommonStruct(Name="binpack", **G.binpack_cluster())
    snap = tc.RegisterSession([[PluginOption.make("binpack", arguments=args, EnabledNodeOrder=True)]])
    s = OracleSession(snap)
    L = pyoracle.lib()
    for t, key in enumerate(snap.task_keys):
        for n, node in enumerate(snap.node_names):
            got = L.vco_binpack_score(s.h, t, n)
            assert abs(got - expected[key][node]) <= G.BINPACK_EPS, (key, node, got)
    s.close()


def test_binpack_arguments():
    args, want = G.BINPACK_ARGUMENTS
    dims = ["cpu", "memory", "example.com/foo", "nvidia.com/gpu", "pods"]
    c = build_conf(SchedulerConf(tiers=[[PluginOption.make("binpack", arguments=args, EnabledNodeOrder=True)]]),
                   dims, ("cpu", "memory", "nvidia.com/gpu"))
    assert c.binpack_weight == want["weight"]
    got = {d: c.binpack_dim_weight[i] for i, d in enumerate(dims)}
    assert got == {"cpu": 5, "memory": 2, "example.com/foo": 1, "nvidia.com/gpu": 7, "pods": -1}


@pytest.mark.parametrize("pct,n,want", G.NUM_FEASIBLE)
def test_num_feasible_nodes(pct, n, want):
    assert pyoracle.lib().vco_num_feasible_nodes(n, pct, 100, 5) == want


def _res(r, names):
    cpu, mem, sc = r
    v = np.zeros(2 + len(names))
    v[0], v[1] = cpu, mem
    has = 0
    for i, nm in enumerate(names):
        if sc and nm in sc:
            v[2 + i] = sc[nm]
            has |= 1 << (2 + i)
    return v, has


@pytest.mark.parametrize("table,infinity", [(G.LESS_EQUAL_ZERO, 0), (G.LESS_EQUAL_INFINITY, 1)])
def test_less_equal(table, infinity):
    names = ["hugepages-test", "scalar.test/scalar1"]
    dp = C.POINTER(C.c_double)
    for l, r, want in table:
        lv, lh = _res(l, names)
        rv, rh = _res(r, names)
        got = pyoracle.lib().vco_less_equal(lv.ctypes.data_as(dp), lh, rv.ctypes.data_as(dp), rh, 4, infinity)
        assert bool(got) == want, (l, r, infinity)


_RNAMES = ["hugepages-test", "scalar.test/scalar1"]


def _res_out(v, has):
    return (v[0], v[1], {nm: v[2 + i] for i, nm in enumerate(_RNAMES) if has & (1 << (2 + i))})


def _expect(r):
    return (float(r[0]), float(r[1]), {k: float(x) for k, x in (r[2] or {}).items()})


def resource_ops(diff_fn, min_fn):
    """Shared by the oracle test below and tests/test_host_logic.py (same goldens against the product's host code)."""
    dp, u32 = C.POINTER(C.c_double), C.c_uint32
    for l, r, inc_w, dec_w in G.DIFF_ZERO:
        lv, lh = _res(l, _RNAMES)
        rv, rh = _res(r, _RNAMES)
        inc, dec = np.zeros(4), np.zeros(4)
        ih, dh = u32(0), u32(0)
        diff_fn(lv.ctypes.data_as(dp), lh, rv.ctypes.data_as(dp), rh, 4, inc.ctypes.data_as(dp), C.byref(ih), dec.ctypes.data_as(dp),
                C.byref(dh))
        assert _res_out(inc, ih.value) == _expect(inc_w) and _res_out(dec, dh.value) == _expect(dec_w), (l, r)
    for table, infinity in ((G.MIN_DIMENSION_ZERO, 0), (G.MIN_DIMENSION_INFINITY, 1)):
        for l, r, want in table:
            lv, lh = _res(l, _RNAMES)
            rv, rh = _res(r, _RNAMES)
            out = np.zeros(4)
            oh = u32(0)
            min_fn(lv.ctypes.data_as(dp), lh, rv.ctypes.data_as(dp), rh, 4, infinity, out.ctypes.data_as(dp), C.byref(oh))
            assert _res_out(out, oh.value) == _expect(want), (l, r, infinity)


def test_less_equal_with_dimension():
    """api/resource_info_test.go:773-883 (non-nil req): the comparison inside proportion's queueAllocatable."""
    dp = C.POINTER(C.c_double)
    for l, r, req, want in G.LESS_EQUAL_WITH_DIMENSION:
        (lv, lh), (rv, rh), (qv, qh) = (_res(x, G.LE_DIM_NAMES) for x in (l, r, req))
        got = pyoracle.lib().vco_less_equal_with_dimension(lv.ctypes.data_as(dp), lh, rv.ctypes.data_as(dp), rh,
                                                           qv.ctypes.data_as(dp), qh, 2 + len(G.LE_DIM_NAMES), -1)
        assert bool(got) == want, (l, r, req)


def test_resource_diff_and_min_dimension():
    """api/resource_info_test.go:315-421 TestDiff (Zero) and :1562-1693 TestMinDimensionResourceZero / Infinity: the two
    Resource operations proportion's water-filling is built from (proportion.go:180-250)."""
    resource_ops(pyoracle.lib().vco_diff_zero, pyoracle.lib().vco_min_dimension)


def test_upstream_score_sanity():
    """recalled kube-scheduler formulas vs the reference's placement goldens."""
    L = pyoracle.lib()
    gi = 1 << 30
    # leastAllocated n1(2c,4Gi) vs n2(4c,8Gi), pod 1c/1G: (50*50+76*50)/100=63 vs (75*50+88*50)/100=81
    assert (L.vco_least_requested_score(1000, 2000) * 50 + L.vco_least_requested_score(10**9, 4 * gi) * 50) // 100 == 63
    assert (L.vco_least_requested_score(1000, 4000) * 50 + L.vco_least_requested_score(10**9, 8 * gi) * 50) // 100 == 81
    assert (L.vco_most_requested_score(1000, 2000) + L.vco_most_requested_score(10**9, 4 * gi)) // 2 == 36
    assert (L.vco_most_requested_score(1000, 4000) + L.vco_most_requested_score(10**9, 8 * gi)) // 2 == 18


@pytest.mark.parametrize("case", G.proportion_cases(), ids=lambda c: c.Name[:40])
def test_proportion_queue_priority(case, oracle_engine):
    case.RegisterSession(G.proportion_tiers())
    case.Run(oracle_engine)
    assert case.CheckAll() is None, case.CheckAll()


def _check_tdm(snap, mask, score, want_nodes, want_scores):
    got = {snap.node_names[n] for n in range(snap.N) if (mask[0, n // 64] >> np.uint64(n % 64)) & np.uint64(1)}
    assert got == want_nodes
    key = snap.task_keys[0]
    for n in range(snap.N):
        if snap.node_names[n] in got:
            assert abs(score[0, n] - want_scores[key][snap.node_names[n]]) <= G.TDM_EPS


@pytest.mark.parametrize("case,nodes,scores", G.tdm_cases(), ids=lambda c: getattr(c, "Name", "x")[:40])
def test_tdm_predicate_and_score(case, nodes, scores):
    snap = case.RegisterSession(G.tdm_tiers())
    o = OracleSession(snap)
    mask, score, _, _ = o.score_matrix()
    o.close()
    _check_tdm(snap, mask, score, nodes, scores)


@pytest.mark.parametrize("raw,delta,err", G.PARSE_REVOCABLE_ZONE)
def test_parse_revocable_zone(raw, delta, err):
    from volcano_b200.snapshot import parse_revocable_zone
    if err:
        with pytest.raises(ValueError):
            parse_revocable_zone(raw)
    else:
        start, end = parse_revocable_zone(raw)
        assert int((end - start).total_seconds()) == delta


def test_tdm_windows_from_arguments():
    """rz1 "0:00-0:00" covers the whole day, rz2 "0:00-0:01" only its first minute (tdm_test.go:218-221)."""
    import datetime as dt
    from volcano_b200.snapshot import tdm_zones_active
    assert tdm_zones_active(G.TDM_ARGUMENTS, dt.datetime(2024, 5, 1, 12, 0)) == {"rz1": True, "rz2": False}
    assert tdm_zones_active(G.TDM_ARGUMENTS, dt.datetime(2024, 5, 1, 0, 0, 30)) == {"rz1": True, "rz2": True}


@pytest.mark.parametrize("case,args,expected", G.nta_cases(), ids=lambda c: getattr(c, "Name", "x")[:40])
def test_nta_batch_node_order_normal_pods(case, args, expected):
    """network_topology_aware_test.go:2011-2766 (pods without a network topology)."""
    snap = case.RegisterSession(G.nta_tiers(args))
    o = OracleSession(snap)
    L = pyoracle.lib()
    t = snap.task_keys.index("ns1/p1")
    for n, name in enumerate(snap.node_names):
        got = L.vco_nta_node_score(o.h, t, n)
        assert abs(got - expected.get(name, 0.0)) <= G.NTA_EPS, (name, got)
    # the same numbers through util.PrioritizeNodes for the feasible nodes
    mask, score, _, _ = o.score_matrix()
    for n, name in enumerate(snap.node_names):
        if (mask[t, n // 64] >> np.uint64(n % 64)) & np.uint64(1):
            assert abs(score[t, n] - expected.get(name, 0.0)) <= G.NTA_EPS
    o.close()


def test_nta_hypernode_resource_cache():
    """network_topology_aware_test.go:3124-3247 Test_initHyperNodeResourceCache."""
    from volcano_b200.api import BuildHyperNode, BuildNode, BuildPod, BuildPodGroup, BuildQueue, BuildResourceList
    N, numbers = G.NTA_CACHE_NODES, G.NTA_CACHE_NUMBERS
    nodes = [BuildNode(f"node-{i}", BuildResourceList("100", "8000", ("example.com/foo", "10"))) for i in range(N)]
    hns = [BuildHyperNode(f"hypernode-tier-{t}-index-{k}", t, [(f"node-{i}", "Node") for i in range(k, N, cnt)])
           for t, cnt in enumerate(numbers, 1) for k in range(cnt)]
    tc = TestCommonStruct(Name="cache", Nodes=nodes, HyperNodes=hns, Queues=[BuildQueue("q1", 1, None)],
                          PodGroups=[BuildPodGroup("pg1", "ns1", "q1", 1, None, "Inqueue")],
                          Pods=[BuildPod("ns1", "p1", "", "Pending", BuildResourceList("1", "1"), "pg1")])
    snap = tc.RegisterSession(G.nta_tiers({}))
    # the reference test fills api.Resource directly (raw 10 for the scalar, not a parsed quantity)
    snap.n_allocatable[snap.dim_names.index("example.com/foo")] = G.NTA_CACHE_PER_NODE["alloc"][2]
    snap.n_used[:] = snap.n_allocatable / 2
    snap.n_used[snap.dim_names.index("pods")] = 0
    o = OracleSession(snap)
    L = pyoracle.lib()
    dims = [snap.dim_names.index(d) for d in ("cpu", "memory", "example.com/foo")]
    a, u = np.zeros(snap.R), np.zeros(snap.R)
    for t, cnt in enumerate(numbers, 1):
        for k in (0, cnt - 1):
            h = snap.hn_names.index(f"hypernode-tier-{t}-index-{k}")
            L.vco_hypernode_status(o.h, h, a.ctypes.data_as(pyoracle._dp), u.ctypes.data_as(pyoracle._dp))
            assert tuple(a[dims]) == tuple(N // cnt * v for v in G.NTA_CACHE_PER_NODE["alloc"])
            assert tuple(u[dims]) == tuple(N // cnt * v for v in G.NTA_CACHE_PER_NODE["used"])
    o.close()


def test_go_pow_small_integer_exponents():
    """Tier weights use Go's math.Pow (network_topology_aware.go:470-476): y == 0 -> 1 (also 0**0), y == 1 -> x,
    otherwise binary exponentiation; fading 0.8 over four tiers gives the 2.952 total behind the 25.8 golden."""
    L = pyoracle.lib()
    assert L.vco_go_pow_uint(0.0, 0) == 1.0 and L.vco_go_pow_uint(0.0, 3) == 0.0
    assert L.vco_go_pow_uint(0.8, 1) == 0.8
    assert L.vco_go_pow_uint(0.8, 2) == 0.8 * 0.8
    assert L.vco_go_pow_uint(0.8, 3) == 0.8 * (0.8 * 0.8)
    assert abs(sum(L.vco_go_pow_uint(0.8, k) for k in range(4)) - 2.952) < 1e-12


@pytest.mark.parametrize("name,kind,allocated,weight,running,score_nodes,expected", G.NTA_SOFT_CASES,
                         ids=[c[0][:40] for c in G.NTA_SOFT_CASES])
def test_nta_soft_topology_node_scores(name, kind, allocated, weight, running, score_nodes, expected):
    """network_topology_aware_test.go:1072-1992 TestNetworkTopologyAwareNodeScore_Soft."""
    tc = G.nta_soft_case(kind, allocated, running)
    snap = tc.RegisterSession(G.nta_tiers({"weight": weight}))
    assert snap.hn_job_soft[0] == 1
    o = OracleSession(snap)
    t = snap.task_keys.index("c1/pending")
    nodes = np.array([snap.node_names.index(n) for n in score_nodes], np.int32)
    out = np.zeros(len(nodes))
    a = snap.hn_names.index(allocated) if allocated else -1
    assert snap.hn_job_allocated[0] == a
    pyoracle.lib().vco_nta_topo_scores(o.h, t, a, nodes.ctypes.data_as(pyoracle._i32p), len(nodes),
                                       out.ctypes.data_as(pyoracle._dp))
    o.close()
    for n, got in zip(score_nodes, out):
        assert abs(got - expected[n]) <= G.NTA_EPS, (n, got)


def test_nta_soft_topology_allocate(oracle_engine):
    tc = G.nta_soft_allocate_case()
    tc.RegisterSession(G.nta_soft_allocate_tiers())
    tc.Run(oracle_engine)
    leaf = {k: v.split("-")[0] for k, v in tc.binds.items()}
    mine = [leaf[f"c1/p{i}"] for i in range(1, 7)]
    assert mine.count("s3") == 3 and mine.count("s4") == 3, mine  # s3 has room for 3 more, then the sibling leaf


@pytest.mark.parametrize("case", G.NTA_SOFT_ALLOCATE_CASES, ids=[c[0][24:70] for c in G.NTA_SOFT_ALLOCATE_CASES])
def test_allocate_with_network_topologies_soft(case, oracle_engine):
    """allocate_test.go:359-747: the soft-mode rows of TestAllocateWithNetWorkTopologies (bind count; bind map where the
    reference pins it).  Rescheduled rows also keep the new pod next to the job's allocated hypernode when it has room."""
    tc = G.nta_soft_allocate_golden(case)
    tc.RegisterSession(G.nta_soft_allocate_golden_tiers())
    tc.Run(oracle_engine)
    assert len(tc.binds) == tc.ExpectBindsNum, tc.binds
    if tc.ExpectBindMap:
        assert tc.CheckBind() is None, tc.CheckBind()


def test_predicates_node_num_allocate_then_backfill(oracle_engine):
    """predicates_test.go:194-259 TestNodeNum: the pod-count cap seen by the backfill action."""
    tc = G.predicates_node_num_case()
    snap = tc.RegisterSession(G.predicates_node_num_tiers(), actions=("allocate", "backfill"))
    assert snap.T == 0 and snap.B == 3
    tc.Run(oracle_engine)
    assert tc.CheckBind() is None, tc.CheckBind()
    assert [snap.backfill_task_keys[t] for t in tc.result.backfill.fit_errors] == ["ns1/worker-3"]


def test_backfill_pick_up_pending_tasks():
    """backfill_test.go:39-154 TestPickUpPendingTasks."""
    tc = G.backfill_pick_case()
    snap = tc.RegisterSession(G.backfill_pick_tiers(), actions=("allocate", "backfill"))
    assert snap.B == 8 and snap.T == 8
    o = OracleSession(snap)
    order = o.backfill_pick_order()
    o.close()
    assert [snap.backfill_task_keys[t] for t in order] == G.BACKFILL_PICK_EXPECTED


@pytest.mark.parametrize("name,alloc,pods,idle,used", G.NODE_INFO_ADD_POD, ids=[c[0][:30] for c in G.NODE_INFO_ADD_POD])
def test_node_accounting_at_session_open(name, alloc, pods, idle, used):
    """api/node_info_test.go:36-131: Idle / Used of a node given the pods the cache holds on it (incl. an Unknown-phase
    pod, which still occupies the node, and an Idle that goes negative)."""
    from volcano_b200.api import BuildNode, BuildPod, BuildPodGroup, BuildQueue, BuildResourceList
    node = BuildNode("n1", BuildResourceList(alloc[0], alloc[1], ("pods", alloc[2])))
    ps = [BuildPod("c1", n, "n1", phase, BuildResourceList(*req), "pg1") for n, phase, req in pods]
    ps.append(BuildPod("c1", "pending", "", "Pending", BuildResourceList("1", "1G"), "pg1"))
    snap = encode_cluster([node], ps, [BuildPodGroup("pg1", "c1", "q1", 1)], [BuildQueue("q1", 1)], SchedulerConf.default())
    dims = [snap.dim_names.index(d) for d in ("cpu", "memory", "pods")]
    assert tuple(snap.n_idle[dims, 0]) == idle
    assert tuple(snap.n_used[dims, 0]) == used


@pytest.mark.parametrize("min_available,is_ready,is_pipelined", G.JOB_INFO_CASES)
def test_job_readiness_counters(min_available, is_ready, is_pipelined):
    """api/job_info_test.go:400-487 TestJobInfo (IsReady / IsPipelined from the status counters)."""
    from volcano_b200.synth import make_snapshot
    snap = make_snapshot("tiny", 1)
    snap.j_ready_num[0] = G.JOB_INFO_COUNTS["ready"]
    snap.j_waiting_num[0] = G.JOB_INFO_COUNTS["waiting"]
    snap.j_pending_besteffort[0] = G.JOB_INFO_COUNTS["pending_besteffort"]
    snap.j_min_available[0] = min_available
    o = OracleSession(snap)
    L = pyoracle.lib()
    assert bool(L.vco_job_is_ready(o.h, 0)) == is_ready
    assert bool(L.vco_job_is_pipelined(o.h, 0)) == is_pipelined
    o.close()


def test_best_effort_and_status_counters_from_pods():
    """The counters themselves as the host mirror derives them from pods: a best-effort pending pod counts in
    PendingBestEffortTaskNum and stays out of the allocate task list (allocate.go:255-271); Running pods are Ready."""
    from volcano_b200.api import BuildNode, BuildPod, BuildPodGroup, BuildQueue, BuildResourceList
    cpu = BuildResourceList("100m", "0")
    pods = [BuildPod("c1", "pending-besteffort", "", "Pending", None, "pg1"),
            BuildPod("c1", "running-besteffort", "n1", "Running", None, "pg1"),
            BuildPod("c1", "pending", "", "Pending", cpu, "pg1"),
            BuildPod("c1", "running", "n1", "Running", cpu, "pg1")]
    snap = encode_cluster([BuildNode("n1", BuildResourceList("4", "8Gi", ("pods", "10")))], pods,
                          [BuildPodGroup("pg1", "c1", "q1", 3)], [BuildQueue("q1", 1)], SchedulerConf.default())
    assert snap.T == 1 and snap.task_keys == ["c1/pending"]
    assert snap.j_pending_besteffort[0] == 1 and snap.j_ready_num[0] == 2 and snap.j_n_tasks_total[0] == 4


@pytest.mark.parametrize("score_map,expected_nodes,expected_score", G.SELECT_BEST_NODE)
def test_select_best_node(score_map, expected_nodes, expected_score):
    """util/scheduler_helper_test.go:34-90: the canonical tie-break (lowest index) is one of the reference's choices."""
    pairs = [(sc, n) for sc, ns in score_map.items() for n in ns]
    scores = np.array([p[0] for p in pairs] or [0.0], np.float64)
    nodes = np.array([p[1] for p in pairs] or [0], np.int32)
    bs = C.c_double(0.0)
    got = pyoracle.lib().vco_select_best(scores.ctypes.data_as(pyoracle._dp), nodes.ctypes.data_as(pyoracle._i32p), len(pairs),
                                         C.byref(bs))
    assert got in expected_nodes
    assert bs.value == expected_score
    if len(expected_nodes) > 1:
        assert got == min(expected_nodes)


@pytest.mark.parametrize("case", G.preempt_cases(), ids=lambda c: c.Name[:50])
def test_preempt_goldens(case, oracle_engine):
    """actions/preempt/preempt_test.go:54-425 TestPreempt on the oracle's restatement of the preempt action."""
    case.RegisterSession(G.preempt_tiers(), actions=("preempt",))
    case.Run(oracle_engine)
    assert case.CheckAll() is None, case.CheckAll()


@pytest.mark.parametrize("plugins,case", G.reclaim_cases(), ids=lambda c: getattr(c, "Name", "p")[:50])
def test_reclaim_goldens(plugins, case, oracle_engine):
    """actions/reclaim/reclaim_test.go:47-388 TestReclaim on the oracle's restatement of the reclaim action."""
    case.RegisterSession(G.reclaim_tiers(plugins), actions=("reclaim",))
    case.Run(oracle_engine)
    assert case.CheckAll() is None, case.CheckAll()


def test_predicate_nodes_error_cache_goldens():
    """util/predicate_helper_test.go:31-220 TestPredicateNodes, the three cases without node sharding, on the oracle's
    ph.PredicateNodes: no nodes -> no result and no cache entry; every node failing the predicate -> no result and a
    cache entry per node under "job/role" (written whatever enableErrorCache says, read only when it is set); a passing
    node -> returned, nothing cached. The reference injects predicate closures; here a node selector nobody / somebody
    satisfies plays that part. (Cases 4-6 exercise --scheduler-sharding-mode, which is outside the path.)"""
    from oracle.pyoracle import OracleSession
    from volcano_b200.api import BuildNode, BuildPod, BuildPodGroup, BuildQueue, BuildResourceList
    from volcano_b200.snapshot import PluginOption, SchedulerConf, encode_cluster

    def session(nodes, selector):
        pods = [BuildPod("ns", "task", "", "Pending", BuildResourceList("1", "1G"), "job1", {"volcano.sh/task-spec": "worker"}, selector)]
        conf = SchedulerConf(tiers=[[PluginOption.make("predicates", EnabledPredicate=True)]])
        return encode_cluster(nodes, pods, [BuildPodGroup("job1", "ns", "q", 1)], [BuildQueue("q", 1)], conf)

    big = BuildResourceList("8", "8G", ("pods", "10"))
    # "empty nodes returns empty result"
    o = OracleSession(session([], {}))
    nodes, cache, exists = o.predicate_nodes(0)
    o.close()
    assert len(nodes) == 0 and not exists
    # "predicate errors returns empty result": expectedErrCache {"job1/worker": {node1, node2}}
    o = OracleSession(session([BuildNode("node1", big, {"zone": "a"}), BuildNode("node2", big, {"zone": "a"})], {"zone": "b"}))
    nodes, cache, exists = o.predicate_nodes(0)
    o.close()
    assert len(nodes) == 0 and exists and list(cache) == [1, 1]
    # "predicate success returns node": expectedErrCache {}
    o = OracleSession(session([BuildNode("node1", big, {"zone": "b"})], {"zone": "b"}))
    nodes, cache, exists = o.predicate_nodes(0)
    o.close()
    assert list(nodes) == [0] and not exists and list(cache) == [0]


def _nominated_cluster(nominated, node0_free="4", mode_pct=None):
    """Two nodes: n0 nearly full (binpack prefers it), n1 empty. One pending pod of 1 cpu with a NominatedNodeName."""
    from volcano_b200.api import BuildNode, BuildPod, BuildPodGroup, BuildQueue, BuildResourceList
    from volcano_b200.snapshot import PluginOption
    from volcano_b200.uthelper import TestCommonStruct
    nodes = [BuildNode("n0", BuildResourceList("8", "16Gi", ("pods", "10"))), BuildNode("n1", BuildResourceList("8", "16Gi", ("pods", "10"))),
             BuildNode("n2", BuildResourceList("8", "16Gi", ("pods", "10")))]
    pods = [BuildPod("c1", "r0", "n0", "Running", BuildResourceList(node0_free, "1Gi"), "pgr"),
            BuildPod("c1", "p0", "", "Pending", BuildResourceList("1", "1Gi"), "pg1"),
            BuildPod("c1", "p1", "", "Pending", BuildResourceList("1", "1Gi"), "pg1")]
    pods[1].nominated_node_name = nominated
    tc = TestCommonStruct(Name="nominated", Nodes=nodes, Pods=pods, Queues=[BuildQueue("q1", 1, None)],
                          PodGroups=[BuildPodGroup("pgr", "c1", "q1", 1), BuildPodGroup("pg1", "c1", "q1", 1)])
    tiers = [[PluginOption.make("priority", EnabledJobOrder=True, EnabledTaskOrder=True), PluginOption.make("gang", EnabledJobReady=True)],
             [PluginOption.make("predicates", EnabledPredicate=True), PluginOption.make("binpack", EnabledNodeOrder=True)]]
    kw = {} if mode_pct is None else dict(percentage_nodes_to_find=mode_pct, min_nodes_to_find=1, last_processed_node_index=2)
    tc.RegisterSession(tiers, **kw)
    return tc


def test_nominated_node_is_tried_first(oracle_engine):
    """actions/allocate/allocate.go:624-634: a pending pod with Status.NominatedNodeName goes to that node when its request fits
    the node's FutureIdle and the predicates pass there - whatever the other nodes would score; otherwise the search over all
    nodes decides. (The reference has no unit test of this branch; the expectations restate the code.)"""
    # binpack alone would send both pods to n0 (4 of 8 cpu used); p0 is nominated to the empty n2
    tc = _nominated_cluster("n2")
    tc.Run(oracle_engine)
    assert tc.binds["c1/p0"] == "n2" and tc.binds["c1/p1"] == "n0"
    d = tc.result.decisions
    assert d["score"][list(d["task"]).index(tc.snap.task_keys.index("c1/p0"))] == 0.0  # a single candidate is not scored
    # nominated node without room (7 of 8 cpu used, the pod asks 1 cpu + ... fits; make it 8): falls back to the search
    tc = _nominated_cluster("n0", node0_free="8")
    tc.Run(oracle_engine)
    assert tc.binds["c1/p0"] in ("n1", "n2")
    # a nominated node that is not in the session counts as none
    tc = _nominated_cluster("gone")
    tc.Run(oracle_engine)
    assert tc.binds["c1/p0"] == "n0"
    assert (tc.snap.t_nominated == -1).all()


def test_nominated_node_resets_the_rotating_index(oracle_engine):
    """ph.PredicateNodes on the one-node list stores (start + processed) % 1 = 0 into util.lastProcessedNodeIndex
    (util/predicate_helper.go:135-136): with feasible-node sampling the next task's scan starts at node 0."""
    from oracle.pyoracle import OracleSession
    tc = _nominated_cluster("n2", mode_pct=34)  # 34 % of 3 nodes -> one feasible node per task, scan starts at index 2
    o = OracleSession(tc.snap, threads=1)
    dec, vis, fe = o.allocate()
    from oracle import pyoracle
    last = pyoracle.lib().vco_last_processed_node_index(o.h)
    o.close()
    names = {tc.snap.task_keys[t]: tc.snap.node_names[n] for t, n in zip(dec["task"], dec["node"])}
    assert names["c1/p0"] == "n2"       # the nominated node
    assert names["c1/p1"] == "n0"       # scanned from index 0 after the reset, first feasible node
    assert last == 1


@pytest.mark.parametrize("enable,case", G.priority_preempt_cases(), ids=lambda c: getattr(c, "Name", str(c))[:50])
def test_priority_plugin_preempt_goldens(enable, case, oracle_engine):
    """plugins/priority/priority_test.go:62-165 TestPreempt: allocate then preempt with the priority plugin alone - the
    victim on the shared node is the pod of lower task priority; with EnabledPreemptable false nothing is evicted."""
    case.RegisterSession(G.priority_preempt_tiers(enable), actions=("allocate", "preempt"))
    case.Run(oracle_engine)
    assert case.CheckAll() is None, case.CheckAll()


def test_conformance_plugin_golden():
    """plugins/conformance/conformance_test.go:32-92 TestConformancePlugin: of the preemptees kube-system/test-pod and
    test-namespace/test-pod only the second is a possible victim (conformance.go:48-63: critical priority classes and the
    kube-system namespace are never evicted). The session encoder carries that as VC_RT_CRITICAL per running task, which is
    all the conformance vote reads (device: ev_may_be_victim, host: EvictSession, oracle: tier_victims)."""
    from volcano_b200 import abi
    from volcano_b200.api import BuildNode, BuildPod, BuildPodGroup, BuildQueue, BuildResourceList
    from volcano_b200.snapshot import PluginOption
    from volcano_b200.uthelper import TestCommonStruct
    R1 = BuildResourceList("1", "1Gi")
    pods = [BuildPod("kube-system", "test-pod", "test-node", "Running", R1, "pg1"),
            BuildPod("test-namespace", "test-pod", "test-node", "Running", R1, "pg2"),
            BuildPod("test-namespace", "crit", "test-node", "Running", R1, "pg2")]
    pods[2].priority_class_name = "system-node-critical"
    tc = TestCommonStruct(Name="conformance plugin", Nodes=[BuildNode("test-node", BuildResourceList("8", "8Gi", ("pods", "10")))],
                          Pods=pods, Queues=[BuildQueue("q1", 1, None)],
                          PodGroups=[BuildPodGroup("pg1", "kube-system", "q1", 1), BuildPodGroup("pg2", "test-namespace", "q1", 1)])
    snap = tc.RegisterSession([[PluginOption.make("conformance", EnabledPreemptable=True)]], actions=("preempt",))
    crit = {snap.running_task_keys[r]: bool(snap.rt_flags[r] & abi.VC_RT_CRITICAL) for r in range(snap.RT)}
    assert crit == {"kube-system/test-pod": True, "test-namespace/test-pod": False, "test-namespace/crit": True}
