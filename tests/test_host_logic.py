"""The product's host-side session-open logic (volcano_b200/csrc/vc_host.hpp: plain C++17, compiled into libvcalloc.so) run
on the CPU through a small test shim (tests/hostshim/host_shim.cpp, g++): reference goldens for the primitives, and the
proportion water-filling / Go-heap task order cross-checked against the independent CPU oracle and a Python container/heap.
No CUDA involved; nothing here is a compute fallback."""
import ctypes as C
import importlib.util
import os
import subprocess

import numpy as np
import pytest

from oracle.pyoracle import OracleSession
from tests.golden import reference_cases as G
from volcano_b200 import abi
from volcano_b200.synth import make_snapshot

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostshim", "host_shim.cpp")
OUT = os.path.join(HERE, "hostshim", "_build", "host_shim.so")
_dp, _i32p = C.POINTER(C.c_double), C.POINTER(C.c_int32)


@pytest.fixture(scope="module")
def shim():
    deps = [SRC, os.path.join(HERE, "..", "volcano_b200", "csrc", "vc_host.hpp"), os.path.join(HERE, "..", "include", "vcalloc.h")]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", SRC, "-o", OUT], check=True)
    L = C.CDLL(OUT)
    L.vh_go_pow_uint.restype = C.c_double
    L.vh_go_pow_uint.argtypes = [C.c_double, C.c_uint]
    L.vh_num_feasible_nodes_to_find.restype = C.c_int32
    L.vh_num_feasible_nodes_to_find.argtypes = [C.c_int32] * 4
    L.vh_job_valid.argtypes = [C.POINTER(abi.vc_conf), C.POINTER(abi.vc_jobs), C.c_int]
    L.vh_proportion_open.argtypes = [C.POINTER(abi.vc_dims), C.POINTER(abi.vc_nodes), C.POINTER(abi.vc_jobs),
                                     C.POINTER(abi.vc_queues), _dp, _dp]
    L.vh_task_heap_order.argtypes = [C.POINTER(abi.vc_tasks), C.c_int, _i32p, C.c_int]
    L.vh_less_equal_zero.argtypes = [_dp, C.c_uint32, _dp, C.c_uint32, C.c_int]
    _u32p = C.POINTER(C.c_uint32)
    L.vh_diff_zero.argtypes = [_dp, C.c_uint32, _dp, C.c_uint32, C.c_int, _dp, _u32p, _dp, _u32p]
    L.vh_min_dimension.argtypes = [_dp, C.c_uint32, _dp, C.c_uint32, C.c_int, C.c_int, _dp, _u32p]
    L.vh_backfill_pick.argtypes = [C.POINTER(abi.vc_dims), C.POINTER(abi.vc_conf), C.POINTER(abi.vc_nodes), C.POINTER(abi.vc_tasks),
                                   C.POINTER(abi.vc_jobs), C.POINTER(abi.vc_queues), C.POINTER(abi.vc_tasks), C.c_int,
                                   C.POINTER(abi.vc_decision), C.c_int, C.c_int, _i32p]
    return L


@pytest.mark.parametrize("pct,n,want", G.NUM_FEASIBLE)
def test_num_feasible_nodes_golden(shim, pct, n, want):
    """util/scheduler_helper_test.go:126-179 against the product's CalculateNumOfFeasibleNodesToFind."""
    assert shim.vh_num_feasible_nodes_to_find(n, pct, 100, 5) == want


def test_less_equal_zero_golden(shim):
    """api/resource_info_test.go:609-771 (Zero flavour) against the product's HRes::less_equal_zero."""
    names = ["hugepages-test", "scalar.test/scalar1"]

    def res(r):
        cpu, mem, sc = r
        v = np.zeros(4)
        v[0], v[1] = cpu, mem
        has = 0
        for i, nm in enumerate(names):
            if sc and nm in sc:
                v[2 + i] = sc[nm]
                has |= 1 << (2 + i)
        return v, has

    for l, r, want in G.LESS_EQUAL_ZERO:
        lv, lh = res(l)
        rv, rh = res(r)
        assert bool(shim.vh_less_equal_zero(lv.ctypes.data_as(_dp), lh, rv.ctypes.data_as(_dp), rh, 4)) == want, (l, r)


def test_resource_diff_and_min_dimension_golden(shim):
    """api/resource_info_test.go:315-421, :1562-1693 against the product's HRes (hdiff, min_dim)."""
    from tests.test_oracle_golden import resource_ops
    resource_ops(shim.vh_diff_zero, shim.vh_min_dimension)


def test_go_pow_matches_the_oracle_and_exact_cases(shim):
    from oracle import pyoracle
    for x in (0.0, 0.5, 0.8, 1.0, 0.3, 2.0, 0.999):
        for n in range(0, 9):
            assert shim.vh_go_pow_uint(x, n) == pyoracle.lib().vco_go_pow_uint(x, n), (x, n)
    assert shim.vh_go_pow_uint(0.5, 3) == 0.125 and shim.vh_go_pow_uint(2.0, 5) == 32.0 and shim.vh_go_pow_uint(0.8, 0) == 1.0
    assert abs(shim.vh_go_pow_uint(0.8, 3) - 0.8 ** 3) < 1e-15


@pytest.mark.parametrize("cfg,seed", [("tiny", None), ("tiny", 4), ("small", None), ("small_fut_soft", 2), ("tiny_bf", None)])
def test_proportion_water_filling_matches_the_oracle(shim, cfg, seed):
    """proportion OnSessionOpen (proportion.go:90-264): deserved per queue and its share, bit-equal between the product's
    host code and the oracle's independent restatement (queues with weights 1..4, two of them capped)."""
    snap = make_snapshot(cfg, seed)
    d, n, j, q = snap.dims(), snap.nodes(), snap.jobs(), snap.queues()
    des = np.zeros((snap.R, snap.Q))
    share = np.zeros(snap.Q)
    shim.vh_proportion_open(C.byref(d), C.byref(n), C.byref(j), C.byref(q), des.ctypes.data_as(_dp), share.ctypes.data_as(_dp))
    o = OracleSession(snap)
    odes, oshare = o.queue_deserved()
    o.close()
    assert np.array_equal(des, odes) and np.array_equal(share, oshare)
    assert des.sum() > 0


def test_proportion_on_api_level_clusters(shim):
    """The same on random API-level clusters (capabilities, guarantees, closed queues, scalar resources)."""
    spec = importlib.util.spec_from_file_location("fuzz_api", os.path.join(HERE, "..", "tools", "fuzz_api.py"))
    # the generator module imports the CUDA engine binding lazily at run time only; importing it needs no GPU
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    checked = 0
    for seed in range(300, 380):
        tc, tiers, actions = mod.make_case(seed)
        if not tiers or not any(po.name == "proportion" for t in tiers for po in t):
            continue
        snap = tc.RegisterSession(tiers, actions=actions, **tc.conf_kw)
        if snap.N == 0 or snap.J == 0:
            continue
        d, n, j, q = snap.dims(), snap.nodes(), snap.jobs(), snap.queues()
        des = np.zeros((snap.R, snap.Q))
        share = np.zeros(snap.Q)
        shim.vh_proportion_open(C.byref(d), C.byref(n), C.byref(j), C.byref(q), des.ctypes.data_as(_dp), share.ctypes.data_as(_dp))
        o = OracleSession(snap)
        odes, oshare = o.queue_deserved()
        o.close()
        assert np.array_equal(des, odes) and np.array_equal(share, oshare), seed
        checked += 1
    assert checked >= 30


def _go_heap_pop_order(items, less):
    """container/heap Push for every item in order, then Pop until empty (util/priority_queue.go:30-111)."""
    h = []

    def up(j):
        while True:
            i = (j - 1) // 2
            if i == j or j == 0 or not less(h[j], h[i]):
                break
            h[i], h[j] = h[j], h[i]
            j = i

    def down(i, n):
        while True:
            j1 = 2 * i + 1
            if j1 >= n or j1 < 0:
                break
            j = j1
            if j1 + 1 < n and less(h[j1 + 1], h[j1]):
                j = j1 + 1
            if not less(h[j], h[i]):
                break
            h[i], h[j] = h[j], h[i]
            i = j

    for x in items:
        h.append(x)
        up(len(h) - 1)
    out = []
    while h:
        n = len(h) - 1
        h[0], h[n] = h[n], h[0]
        down(0, n)
        out.append(h.pop())
    return out


def test_task_order_is_the_go_heap_pop_order(shim):
    """ssn.TaskOrderFn (priority, then pod-name index / creation time / UID, helpers.go:54-69) through a Go binary heap:
    the comparator is not a strict weak order when pod indices are missing, so the pop order - not a sort - is the contract."""
    rng = np.random.default_rng(7)
    snap = make_snapshot("tiny")
    for trial in range(40):
        n = int(rng.integers(2, 40))
        snap.t_priority[:n] = rng.integers(0, 3, n)
        snap.t_pod_index[:n] = np.where(rng.random(n) < 0.3, -1, rng.integers(0, 6, n))
        snap.t_creation_ts[:n] = rng.integers(0, 4, n)
        snap.t_uid_rank[:n] = rng.permutation(n)
        by_prio = trial % 2

        def less(l, r):
            if by_prio and snap.t_priority[l] != snap.t_priority[r]:
                return snap.t_priority[l] > snap.t_priority[r]
            li, ri = snap.t_pod_index[l], snap.t_pod_index[r]
            if li < 0 or ri < 0 or li == ri:
                if snap.t_creation_ts[l] == snap.t_creation_ts[r]:
                    return snap.t_uid_rank[l] < snap.t_uid_rank[r]
                return snap.t_creation_ts[l] < snap.t_creation_ts[r]
            return not (li > ri)

        items = np.arange(n, dtype=np.int32)
        tk = snap.tasks()
        shim.vh_task_heap_order(C.byref(tk), by_prio, items.ctypes.data_as(_i32p), n)
        assert list(items) == _go_heap_pop_order(list(range(n)), less), trial


def test_job_valid_gang(shim):
    """gang validJobFn (gang.go:58-93): valid tasks against MinAvailable and the per-role minima."""
    from volcano_b200.api import BuildNode, BuildPod, BuildPodGroup, BuildQueue, BuildResourceList
    from volcano_b200.snapshot import SchedulerConf, encode_cluster
    req = BuildResourceList("1", "1G")
    pods = [BuildPod("c1", f"a{i}", "", "Pending", req, "pg1", {"volcano.sh/task-spec": "w"}) for i in range(2)]
    pods += [BuildPod("c1", f"b{i}", "", "Pending", req, "pg2", {"volcano.sh/task-spec": "w"}) for i in range(3)]
    pods += [BuildPod("c1", f"c{i}", "", "Pending", req, "pg3", {"volcano.sh/task-spec": "m"}) for i in range(2)]
    pods += [BuildPod("c1", "d0", "", "Pending", req, "pg4", {"volcano.sh/task-spec": "m"})]
    pgs = [BuildPodGroup("pg1", "c1", "q1", 3), BuildPodGroup("pg2", "c1", "q1", 3),
           BuildPodGroup("pg3", "c1", "q1", 2, {"m": 1, "w": 1}), BuildPodGroup("pg4", "c1", "q1", 1, {"m": 1, "w": 1})]
    snap = encode_cluster([BuildNode("n1", BuildResourceList("8", "8G", ("pods", "10")))], pods, pgs, [BuildQueue("q1", 1)],
                          SchedulerConf.default())
    j = snap.jobs()
    got = [shim.vh_job_valid(C.byref(snap.conf), C.byref(j), k) for k in range(4)]
    # 2 < minAvailable 3; 3 >= 3; enough tasks but role "w" has none against its minimum 1 (CheckTaskValid,
    # job_info.go:993-1019); MinAvailable 1 < sum of the role minima 2: the role check is skipped
    assert got == [0, 1, 0, 1]


def _pick_vs_oracle(shim, snap):
    o = OracleSession(snap)
    dec, vis, fe = o.allocate()
    want = o.backfill_pick_order()  # on the state allocate left
    o.close()
    d, n, t, j, q, bt = snap.dims(), snap.nodes(), snap.tasks(), snap.jobs(), snap.queues(), snap.backfill_tasks()
    ops = np.ascontiguousarray(dec)
    out = np.zeros(max(snap.B, 1), np.int32)
    got_n = shim.vh_backfill_pick(C.byref(d), C.byref(snap.conf), C.byref(n), C.byref(t), C.byref(j), C.byref(q), C.byref(bt),
                                  snap.B, ops.ctypes.data_as(C.POINTER(abi.vc_decision)), len(ops), 1, out.ctypes.data_as(_i32p))
    assert list(out[:got_n]) == list(want)
    return got_n


@pytest.mark.parametrize("cfg,seed", [("tiny_bf", None), ("tiny_bf", 5), ("small_bf", None), ("small_soft_bf", 3)])
def test_backfill_pick_order_matches_the_oracle(shim, cfg, seed):
    """pickUpPendingTasks (backfill.go:118-199) as vc_backfill_run computes it on the host - session state rebuilt from
    allocate's operation list, queues / jobs / tasks in Go-heap pop order - against the oracle, which walks its own live
    session after its own allocate pass."""
    snap = make_snapshot(cfg, seed)
    assert _pick_vs_oracle(shim, snap) == snap.B


def test_backfill_pick_order_on_api_level_clusters(shim):
    spec = importlib.util.spec_from_file_location("fuzz_api", os.path.join(HERE, "..", "tools", "fuzz_api.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    checked = picked = 0
    for seed in range(500, 700):
        tc, tiers, actions = mod.make_case(seed)
        if not tiers or "backfill" not in actions:
            continue
        snap = tc.RegisterSession(tiers, actions=actions, **tc.conf_kw)
        if snap.B == 0 or snap.N == 0 or snap.T == 0:
            continue
        picked += _pick_vs_oracle(shim, snap)
        checked += 1
    assert checked >= 40 and picked >= 60


@pytest.mark.parametrize("seed", range(6))
def test_proportion_deserved_properties(shim, seed):
    """Size-independent properties of the water-filling (proportion.go:180-250) on the product's host code: a queue never
    deserves more than it requests or than the cluster holds, the queues together never more than the cluster, and a queue
    with pending requests and no cap gets a positive share of cpu."""
    snap = make_snapshot("small", seed)
    d, n, j, q = snap.dims(), snap.nodes(), snap.jobs(), snap.queues()
    des = np.zeros((snap.R, snap.Q))
    share = np.zeros(snap.Q)
    shim.vh_proportion_open(C.byref(d), C.byref(n), C.byref(j), C.byref(q), des.ctypes.data_as(_dp), share.ctypes.data_as(_dp))
    total = snap.n_allocatable.sum(axis=1)
    assert np.all(des >= 0)
    assert np.all(des <= snap.q_request + 1e-6)             # MinDimensionResource(request)
    assert np.all(des.sum(axis=1) <= total * (1 + 1e-12))    # remaining never goes negative
    for qi in range(snap.Q):
        if snap.q_request[0, qi] > 0 and not (snap.q_capability_has[qi] & abi.VC_RES_HAS_ANY):
            assert des[0, qi] > 0
        if snap.q_capability_has[qi] & abi.VC_RES_HAS_ANY:    # capped queues stay under their capability
            assert np.all(des[:2, qi] <= snap.q_capability[:2, qi] + 1e-6)
    assert np.all(share >= 0)


def test_run_length_exactness_preconditions(shim):
    """vch::runs_exact / future_rows_exact (vc_host.hpp): run-length batches need integer-valued rows and requests with
    |row| (+ |Releasing| + |Pipelined| per node for the FutureIdle expression) + 32 max|request| below 2^53."""
    import ctypes as C
    dp = C.POINTER(C.c_double)
    shim.vh_runs_exact.argtypes = [C.c_int] * 3 + [dp] * 3
    shim.vh_future_rows_exact.argtypes = [C.c_int] * 3 + [dp] * 4

    def arr(a):
        a = np.ascontiguousarray(a, np.float64)
        return a, a.ctypes.data_as(dp)

    D, N, T = 2, 5, 3
    idle, pi = arr(np.array([[4000, 8000, 0, 16000, 32000], [8e15, 7.9e15, 0, 1, 2]]))  # an ephemeral-storage-sized dimension
    used, pu = arr(np.zeros((D, N)))
    req, pr = arr(np.array([[100, 250, 4000], [0, 0, 0]]))
    assert shim.vh_runs_exact(D, N, T, pi, pu, pr) == 1
    frac, pf = arr(np.array([[100.5, 250, 4000], [0, 0, 0]]))
    assert shim.vh_runs_exact(D, N, T, pi, pu, pf) == 0           # a fractional request
    big, pb = arr(np.array([[100, 250, 4000], [0, 0, 5e13]]))
    assert shim.vh_runs_exact(D, N, T, pi, pu, pb) == 0           # 8e15 + 32 * 5e13 leaves the exact range
    rel, prl = arr(np.array([[0, 1000, 0, 0, 0], [0, 0, 0, 0, 0]]))
    pip, pp = arr(np.zeros((D, N)))
    assert shim.vh_future_rows_exact(D, N, T, pi, prl, pp, pr) == 1
    rel2, prl2 = arr(np.array([[0, 1000, 0, 0, 0], [2e15, 0, 0, 0, 0]]))
    assert shim.vh_future_rows_exact(D, N, T, pi, prl2, pp, pr) == 0   # Idle + Releasing of node 0 passes 2^53
    assert shim.vh_future_rows_exact(D, N, T, pi, None, None, pr) == 1  # no Releasing / Pipelined rows at all
