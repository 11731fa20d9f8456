"""ctypes loader for the CPU oracle (oracle/oracle.cpp).

TEST INFRASTRUCTURE ONLY: import from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product (volcano_b200/) never imports this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from volcano_b200 import abi
from volcano_b200.snapshot import Snapshot

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "oracle.cpp")
    hdr = os.path.join(_HERE, "..", "include", "vcalloc.h")
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(p) > os.path.getmtime(_LIB_PATH) for p in (src, hdr))
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "_build/liboracle.so"], check=True, capture_output=True)
    return _LIB_PATH


_vp = C.c_void_p
_dp, _i32p, _u64p = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_uint64)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.vco_session_create.restype = _vp
        L.vco_session_create.argtypes = [C.POINTER(abi.vc_dims), C.POINTER(abi.vc_nodes), C.POINTER(abi.vc_tasks),
                                         C.POINTER(abi.vc_classes), C.POINTER(abi.vc_jobs), C.POINTER(abi.vc_queues),
                                         C.POINTER(abi.vc_conf), C.c_int]
        L.vco_session_destroy.argtypes = [_vp]
        L.vco_session_set_topology.argtypes = [_vp, C.POINTER(abi.vc_hypernodes)]
        L.vco_hypernode_status.argtypes = [_vp, C.c_int, _dp, _dp]
        L.vco_nta_node_score.restype = C.c_double
        L.vco_nta_node_score.argtypes = [_vp, C.c_int, C.c_int]
        L.vco_nta_topo_scores.argtypes = [_vp, C.c_int, C.c_int, _i32p, C.c_int, _dp]
        L.vco_job_allocated_hypernode.restype = C.c_int
        L.vco_job_allocated_hypernode.argtypes = [_vp, C.c_int]
        L.vco_last_processed_node_index.restype = C.c_int64
        L.vco_last_processed_node_index.argtypes = [_vp]
        L.vco_go_pow_uint.restype = C.c_double
        L.vco_go_pow_uint.argtypes = [C.c_double, C.c_uint]
        L.vco_allocate_run.argtypes = [_vp]
        L.vco_predicate_nodes.restype = C.c_int
        L.vco_predicate_nodes.argtypes = [_vp, C.c_int, _i32p, C.POINTER(C.c_uint8), C.POINTER(C.c_int)]
        L.vco_replay_check.restype = C.c_int64
        L.vco_replay_check.argtypes = [_vp, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int64, C.c_int64,
                                       C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.vco_session_set_backfill.argtypes = [_vp, C.c_int32, C.POINTER(abi.vc_tasks)]
        L.vco_backfill.argtypes = [_vp]
        L.vco_session_set_running.argtypes = [_vp, C.POINTER(abi.vc_running_tasks), C.POINTER(C.c_uint32)]
        L.vco_preempt.argtypes = [_vp]
        L.vco_reclaim.argtypes = [_vp]
        L.vco_backfill_pick_order.argtypes = [_vp, _i32p]
        for n in ("vco_num_decisions", "vco_num_visits", "vco_num_fit_errors"):
            getattr(L, n).restype = C.c_size_t
            getattr(L, n).argtypes = [_vp]
        L.vco_decisions.restype = C.POINTER(abi.vc_decision)
        L.vco_decisions.argtypes = [_vp]
        L.vco_visits.restype = C.POINTER(abi.vc_visit)
        L.vco_visits.argtypes = [_vp]
        L.vco_fit_errors.restype = _i32p
        L.vco_fit_errors.argtypes = [_vp]
        L.vco_num_sweeps.restype = C.c_int64
        L.vco_num_sweeps.argtypes = [_vp]
        L.vco_score_matrix.argtypes = [_vp, _u64p, _dp, _dp, _i32p]
        L.vco_queue_deserved.argtypes = [_vp, _dp, _dp]
        L.vco_node_state.argtypes = [_vp, _dp, _dp, _dp]
        L.vco_less_equal.restype = C.c_int
        L.vco_less_equal.argtypes = [_dp, C.c_uint32, _dp, C.c_uint32, C.c_int, C.c_int]
        _u32p = C.POINTER(C.c_uint32)
        L.vco_diff_zero.argtypes = [_dp, C.c_uint32, _dp, C.c_uint32, C.c_int, _dp, _u32p, _dp, _u32p]
        L.vco_less_equal_with_dimension.argtypes = [_dp, C.c_uint32, _dp, C.c_uint32, _dp, C.c_uint32, C.c_int, C.c_int]
        L.vco_min_dimension.argtypes = [_dp, C.c_uint32, _dp, C.c_uint32, C.c_int, C.c_int, _dp, _u32p]
        L.vco_num_feasible_nodes.restype = C.c_int32
        L.vco_num_feasible_nodes.argtypes = [C.c_int32] * 4
        for n in ("vco_least_requested_score", "vco_most_requested_score"):
            getattr(L, n).restype = C.c_int64
            getattr(L, n).argtypes = [C.c_int64, C.c_int64]
        for n in ("vco_binpack_score", "vco_nodeorder_score"):
            getattr(L, n).restype = C.c_double
            getattr(L, n).argtypes = [_vp, C.c_int, C.c_int]
        for n in ("vco_least_allocated", "vco_most_allocated", "vco_balanced_allocation"):
            getattr(L, n).restype = C.c_int64
            getattr(L, n).argtypes = [_vp, C.c_int, C.c_int]
        L.vco_predicate.restype = C.c_int
        L.vco_predicate.argtypes = [_vp, C.c_int, C.c_int]
        L.vco_job_share.restype = C.c_double
        L.vco_job_share.argtypes = [_vp, C.c_int]
        L.vco_job_ready.restype = C.c_int
        L.vco_job_ready.argtypes = [_vp, C.c_int]
        L.vco_select_best.restype = C.c_int
        L.vco_select_best.argtypes = [_dp, _i32p, C.c_int, _dp]
        for n in ("vco_job_is_ready", "vco_job_is_pipelined"):
            getattr(L, n).restype = C.c_int
            getattr(L, n).argtypes = [_vp, C.c_int]
        _lib = L
    return _lib


DECISION_DTYPE = np.dtype([("task", "<i4"), ("node", "<i4"), ("kind", "<i4"), ("visit", "<i4"), ("score", "<f8")])
VISIT_DTYPE = np.dtype([("job", "<i4"), ("outcome", "<i4"), ("first_op", "<i4"), ("n_ops", "<i4")])


class OracleSession:
    def __init__(self, snap: Snapshot, threads: int = 1):
        self.snap = snap
        L = lib()
        d, n, t, c, j, q = snap.dims(), snap.nodes(), snap.tasks(), snap.classes(), snap.jobs(), snap.queues()
        self.h = L.vco_session_create(C.byref(d), C.byref(n), C.byref(t), C.byref(c), C.byref(j), C.byref(q),
                                      C.byref(snap.conf), threads)
        if not self.h:
            raise RuntimeError("oracle session_create failed")
        topo = snap.topology()
        if topo is not None:
            L.vco_session_set_topology(self.h, C.byref(topo))
        bt = snap.backfill_tasks()
        if bt is not None and L.vco_session_set_backfill(self.h, snap.B, C.byref(bt)) != 0:
            raise RuntimeError("oracle set_backfill failed")
        nom = getattr(snap, "t_nominated", None)
        if nom is not None and snap.T and (nom >= 0).any():
            L.vco_session_set_nominated.argtypes = [_vp, C.POINTER(C.c_int32)]
            L.vco_session_set_nominated(self.h, np.ascontiguousarray(nom, np.int32).ctypes.data_as(C.POINTER(C.c_int32)))
        rt = snap.running_tasks()
        if rt is not None or snap.t_flags.any():
            tf = snap.t_flags.ctypes.data_as(C.POINTER(C.c_uint32)) if snap.T else None
            if L.vco_session_set_running(self.h, C.byref(rt) if rt is not None else None, tf) != 0:
                raise RuntimeError("oracle set_running failed")

    def close(self):
        if self.h:
            lib().vco_session_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def allocate(self):
        rc = lib().vco_allocate_run(self.h)
        if rc != 0:
            raise RuntimeError(f"oracle allocate rc={rc}")
        return self._results()

    def replay_check(self, decisions, visits, stride: int, offset: int = 0):
        """Apply `decisions` (any implementation's) to this FRESH session in order and re-derive every stride-th one on
        the state reached so far (feasible nodes + prioritizeNodes). -> (mismatches, first bad index or -1, checked)."""
        dec = np.ascontiguousarray(decisions, DECISION_DTYPE)
        vis = np.ascontiguousarray(visits, VISIT_DTYPE)
        first, checked = C.c_int64(-1), C.c_int64(0)
        bad = lib().vco_replay_check(self.h, dec.ctypes.data, len(dec), vis.ctypes.data, len(vis), stride, offset,
                                     C.byref(first), C.byref(checked))
        return int(bad), int(first.value), int(checked.value)

    def backfill(self):
        """The backfill action on the state the session is in (after allocate(), or the opening state)."""
        rc = lib().vco_backfill(self.h)
        if rc != 0:
            raise RuntimeError(f"oracle backfill rc={rc}")
        return self._results()

    def preempt(self):
        """The preempt action on the state the session is in (decision.task of VC_OP_EVICT indexes running_task_keys)."""
        rc = lib().vco_preempt(self.h)
        if rc != 0:
            raise RuntimeError(f"oracle preempt rc={rc}")
        return self._results()

    def reclaim(self):
        rc = lib().vco_reclaim(self.h)
        if rc != 0:
            raise RuntimeError(f"oracle reclaim rc={rc}")
        return self._results()

    def predicate_nodes(self, t: int):
        """ph.PredicateNodes for task t with a fresh helper -> (feasible nodes, error-cache marks per node, group exists)."""
        N = self.snap.N
        nodes = np.zeros(max(N, 1), np.int32)
        cache = np.zeros(max(N, 1), np.uint8)
        ex = C.c_int(0)
        n = lib().vco_predicate_nodes(self.h, t, nodes.ctypes.data_as(_i32p), cache.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(ex))
        return nodes[:n].copy(), cache[:N].copy(), bool(ex.value)

    def backfill_pick_order(self):
        out = np.zeros(max(self.snap.B, 1), np.int32)
        n = lib().vco_backfill_pick_order(self.h, out.ctypes.data_as(_i32p))
        return out[:n].copy()

    def _results(self):
        L = lib()
        nd = L.vco_num_decisions(self.h)
        nv = L.vco_num_visits(self.h)
        nf = L.vco_num_fit_errors(self.h)
        dec = np.ctypeslib.as_array(C.cast(L.vco_decisions(self.h), C.POINTER(C.c_uint8)),
                                    (nd * DECISION_DTYPE.itemsize,)).view(DECISION_DTYPE).copy() if nd else np.zeros(0, DECISION_DTYPE)
        vis = np.ctypeslib.as_array(C.cast(L.vco_visits(self.h), C.POINTER(C.c_uint8)),
                                    (nv * VISIT_DTYPE.itemsize,)).view(VISIT_DTYPE).copy() if nv else np.zeros(0, VISIT_DTYPE)
        fe = np.ctypeslib.as_array(L.vco_fit_errors(self.h), (nf,)).copy() if nf else np.zeros(0, np.int32)
        return dec, vis, fe

    def score_matrix(self):
        s = self.snap
        mw = (s.N + 63) // 64
        mask = np.zeros((s.T, mw), np.uint64)
        score = np.zeros((s.T, s.N), np.float64)
        bs = np.zeros(s.T, np.float64)
        bn = np.zeros(s.T, np.int32)
        lib().vco_score_matrix(self.h, mask.ctypes.data_as(_u64p), score.ctypes.data_as(_dp), bs.ctypes.data_as(_dp),
                               bn.ctypes.data_as(_i32p))
        return mask, score, bs, bn

    def queue_deserved(self):
        s = self.snap
        des = np.zeros((s.R, s.Q))
        share = np.zeros(s.Q)
        lib().vco_queue_deserved(self.h, des.ctypes.data_as(_dp), share.ctypes.data_as(_dp))
        return des, share

    def node_state(self):
        s = self.snap
        idle, used, pip = (np.zeros((s.R, s.N)) for _ in range(3))
        lib().vco_node_state(self.h, idle.ctypes.data_as(_dp), used.ctypes.data_as(_dp), pip.ctypes.data_as(_dp))
        return idle, used, pip

    def num_sweeps(self) -> int:
        return int(lib().vco_num_sweeps(self.h))
