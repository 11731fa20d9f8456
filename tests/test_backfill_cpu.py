"""CPU-side checks of the backfill action's host logic and of its oracle (actions/backfill/backfill.go): what the encoder
hands to vc_snapshot_set_backfill, and the semantics the CUDA path is compared against on the GPU box."""
import numpy as np

from oracle.pyoracle import OracleSession
from volcano_b200 import abi
from volcano_b200.api import BuildNode, BuildPod, BuildPodGroup, BuildQueue, BuildResourceList
from volcano_b200.snapshot import SchedulerConf
from volcano_b200.synth import make_snapshot
from volcano_b200.uthelper import AllocateResult, TestCommonStruct


def _cluster():
    nodes = [BuildNode(f"n{i}", BuildResourceList("4", "8Gi", ("pods", "3"))) for i in range(3)]
    pods = [BuildPod("c1", f"w{i}", "", "Pending", BuildResourceList("1", "1Gi"), "pg1") for i in range(3)]
    pods += [BuildPod("c1", f"be{i}", "", "Pending", {}, "pg1") for i in range(4)]
    pods += [BuildPod("c1", f"xbe{i}", "", "Pending", {}, "pg2") for i in range(4)]
    pods += [BuildPod("c1", "xw0", "", "Pending", BuildResourceList("64", "1Gi"), "pg2")]
    return TestCommonStruct(Name="allocate+backfill", Nodes=nodes, Pods=pods, Queues=[BuildQueue("q1", 1, None)],
                            PodGroups=[BuildPodGroup("pg1", "c1", "q1", 2), BuildPodGroup("pg2", "c1", "q1", 5)])


def test_encoder_splits_besteffort_tasks():
    tc = _cluster()
    snap = tc.RegisterSession(SchedulerConf.default().tiers, actions=("allocate", "backfill"))
    assert snap.T == 4 and snap.B == 8
    assert snap.backfill_task_keys == [f"c1/be{i}" for i in range(4)] + [f"c1/xbe{i}" for i in range(4)]
    assert list(snap.j_pending_besteffort) == [4, 4]
    pods_dim = snap.dim_names.index("pods")
    assert np.all(snap.b_resreq[pods_dim] == 1.0) and snap.b_resreq.sum() == 8.0  # Resreq = pods:1 only
    assert np.all(snap.b_k8s_nonzero_req[0] == 100.0) and np.all(snap.b_k8s_nonzero_req[1] == 200.0 * 2**20)
    bt = snap.backfill_tasks()
    assert isinstance(bt, abi.vc_tasks)
    # a session without BestEffort pods hands nothing over
    assert make_snapshot("tiny").backfill_tasks() is None


def test_backfill_semantics_on_the_oracle(oracle_engine):
    """Order (gang's JobOrderFn puts the unready job first), no resource fit but the pod-count predicate, one-candidate
    shortcut, dispatch only when ssn.JobReady holds."""
    tc = _cluster()
    snap = tc.RegisterSession(SchedulerConf.default().tiers, actions=("allocate", "backfill"))
    tc.Run(oracle_engine)
    bf = tc.result.backfill
    assert [snap.job_names[j] for j in bf.visits["job"]] == ["c1/pg2", "c1/pg1"]
    assert list(bf.visits["outcome"]) == [abi.VC_VISIT_KEEP, abi.VC_VISIT_COMMIT]
    assert len(bf.decisions) == 6 and [snap.backfill_task_keys[t] for t in bf.fit_errors] == ["c1/be2", "c1/be3"]
    assert sorted(tc.binds) == ["c1/be0", "c1/be1", "c1/w0", "c1/w1", "c1/w2"]
    assert bf.decisions["score"][-1] == 0.0
    per_node = np.bincount(np.concatenate([tc.result.decisions["node"], bf.decisions["node"]]), minlength=3)
    assert list(per_node) == [3, 3, 3]


def test_backfill_alone_and_synthetic_configs():
    snap = make_snapshot("tiny_bf")
    assert snap.B == 150 and "backfill" in snap.actions
    o = OracleSession(snap)
    order = o.backfill_pick_order()
    assert sorted(order) == list(range(snap.B))  # every job is valid and not Pending: all tasks are picked up
    jobs = snap.b_job[order]
    assert len(set(zip(jobs[:-1], jobs[1:])) - {(a, a) for a in jobs}) == len(set(jobs)) - 1  # one contiguous run per job
    dec, vis, fe = o.backfill()  # no allocate before: the opening state
    o.close()
    assert len(dec) + len(fe) == snap.B and vis["n_ops"].sum() == len(dec)
    res = AllocateResult(dec, vis, fe)
    assert np.all(np.diff(res.decisions["visit"]) >= 0)


def _pending_phase_cluster():
    """Two PodGroups still in phase Pending (no enqueue action has promoted them), BestEffort pods only in pg2."""
    nodes = [BuildNode("n0", BuildResourceList("4", "8Gi", ("pods", "8")))]
    pods = [BuildPod("c1", "w0", "", "Pending", BuildResourceList("1", "1Gi"), "pg1")]
    pods += [BuildPod("c1", f"be{i}", "", "Pending", {}, "pg2") for i in range(2)]
    pods += [BuildPod("c1", "w1", "", "Pending", BuildResourceList("1", "1Gi"), "pg2")]
    return TestCommonStruct(Name="pending phase", Nodes=nodes, Pods=pods, Queues=[BuildQueue("q1", 1, None)],
                            PodGroups=[BuildPodGroup("pg1", "c1", "q1", 1, None, "Pending"),
                                       BuildPodGroup("pg2", "c1", "q1", 1, None, "Pending")])


def test_pending_podgroups_without_the_enqueue_action(oracle_engine):
    """allocate.go:154-164: with no enqueue action configured buildAllocateContext rewrites the phase of every Pending
    PodGroup to Inqueue, so the backfill action that follows (backfill.go:124 job.IsPending()) places their BestEffort
    pods; with enqueue configured both actions skip such jobs."""
    tc = _pending_phase_cluster()
    snap = tc.RegisterSession(SchedulerConf.default().tiers, actions=("allocate", "backfill"))
    assert snap.conf.enqueue_action_enabled == 0 and (snap.j_flags & abi.VC_JOB_PENDING_PHASE).all()
    tc.Run(oracle_engine)
    assert sorted(tc.binds) == ["c1/be0", "c1/be1", "c1/w0", "c1/w1"]
    tc = _pending_phase_cluster()
    tc.RegisterSession(SchedulerConf.default().tiers, actions=("enqueue", "allocate", "backfill"))
    tc.Run(oracle_engine)
    assert tc.binds == {}
    # backfill alone (allocate did not run, nothing rewrote the phase): Pending jobs are skipped
    tc = _pending_phase_cluster()
    snap = tc.RegisterSession(SchedulerConf.default().tiers, actions=("backfill",))
    o = OracleSession(snap)
    assert len(o.backfill_pick_order()) == 0
    o.close()
