"""framework.Action mirrors of the preempt and reclaim actions (actions/preempt/preempt.go:60-100,
actions/reclaim/reclaim.go:34-54).

    preempt.New() / reclaim.New()  ->  volcano_b200.preempt.New("preempt") / New("reclaim")
    action.Execute(ssn)            ->  vc_preempt_run / vc_reclaim_run on the device-side session of the cycle, then the
                                       committed statements replayed: Statement.Evict -> evictor, Statement.Pipeline

Listed after `volcano_b200.action.New()` in `TestCommonStruct.Run([...])` like the configured action order
"allocate, preempt, reclaim"; alone, they act on the opening session (the reference's own unit tests do that)."""
from __future__ import annotations

import numpy as np

from . import abi, engine
from .action import Session
from .uthelper import AllocateResult


class Action:
    def __init__(self, name: str, device: int = 0):
        assert name in ("preempt", "reclaim")
        self.name, self.device = name, device

    def Name(self) -> str:
        return self.name

    def Initialize(self) -> None:
        engine.init(self.device)

    def Execute(self, ssn: Session) -> None:
        dev = ssn.device_session(self.device)
        res = dev.preempt() if self.name == "preempt" else dev.reclaim()
        if ssn.result is None:
            ssn.result = AllocateResult(np.zeros(0, engine.DECISION_DTYPE), np.zeros(0, engine.VISIT_DTYPE), np.zeros(0, np.int32))
        setattr(ssn.result, self.name, res)
        snap = ssn.snapshot
        for v in res.visits:
            if v["outcome"] != abi.VC_VISIT_COMMIT:
                continue
            for op in res.decisions[v["first_op"]: v["first_op"] + v["n_ops"]]:
                if op["kind"] == abi.VC_OP_PIPELINE:
                    ssn.pipelined.setdefault(snap.job_names[v["job"]], []).append(snap.node_names[op["node"]])

    def UnInitialize(self) -> None:
        pass


def New(name: str = "preempt", device: int = 0) -> Action:
    return Action(name, device)
