"""framework.Action mirror for the allocate path (framework/interface.go:30-51, actions/allocate/allocate.go:105-140).

    action = allocate.New()          ->  Action()
    action.Name()                    ->  "allocate"
    action.Initialize()              ->  binds the process to its CUDA device (vc_init)
    action.Execute(ssn)              ->  marshal the session, run on the GPU, replay the operations
    action.UnInitialize()

`Session` here is the host mirror of the opened framework.Session: the encoded snapshot plus what Statement.Commit /
Pipeline would have left behind after the replay (binds, pipelined tasks, fit errors, allocated hypernodes).  It is what
`uthelper.TestCommonStruct.RegisterSession` returns a snapshot for; `TestCommonStruct.Run([Action()])` accepts it like
the reference's `test.Run(actions)`.  There is no CPU path: Execute raises engine.VcError(VC_ENODEV) without a GPU."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

from . import abi, engine
from .snapshot import Snapshot
from .uthelper import AllocateResult


@dataclass
class Session:
    snapshot: Snapshot
    result: Optional[AllocateResult] = None
    binds: Dict[str, str] = field(default_factory=dict)            # task key -> node (FakeBinder channel)
    pipelined: Dict[str, List[str]] = field(default_factory=dict)  # job -> nodes of pipelined tasks
    fit_errors: List[str] = field(default_factory=list)            # tasks with job.NodesFitErrors entries
    _engine: Optional[engine.Engine] = None                        # device-side session shared by the actions of a cycle

    def device_session(self, device: int = 0) -> engine.Engine:
        if self._engine is None:
            self._engine = engine.Engine(self.snapshot, device)
            self._engine.upload()
        return self._engine

    def close(self) -> None:
        if self._engine is not None:
            self._engine.close()
            self._engine = None

    def replay(self, result: AllocateResult) -> None:
        """Statement.Allocate / Pipeline + Commit for every kept visit (framework/statement.go:146-412)."""
        self.result = result
        snap = self.snapshot
        for v in result.visits:
            for op in result.decisions[v["first_op"]: v["first_op"] + v["n_ops"]]:
                key, node = snap.task_keys[op["task"]], snap.node_names[op["node"]]
                if op["kind"] == abi.VC_OP_ALLOCATE and v["outcome"] == abi.VC_VISIT_COMMIT:
                    self.binds[key] = node
                elif op["kind"] == abi.VC_OP_PIPELINE:
                    self.pipelined.setdefault(snap.job_names[v["job"]], []).append(node)
        self.fit_errors = [snap.task_keys[t] for t in result.fit_errors]

    def replay_backfill(self, result: AllocateResult) -> None:
        """Session.Allocate for every backfilled task (framework/session.go:746-796): bound when ssn.JobReady held."""
        snap = self.snapshot
        if self.result is None:
            self.result = AllocateResult(result.decisions[:0], result.visits[:0], result.fit_errors[:0])
        self.result.backfill = result
        for v in result.visits:
            if v["outcome"] != abi.VC_VISIT_COMMIT:
                continue
            for op in result.decisions[v["first_op"]: v["first_op"] + v["n_ops"]]:
                self.binds[snap.backfill_task_keys[op["task"]]] = snap.node_names[op["node"]]
        self.fit_errors += [snap.backfill_task_keys[t] for t in result.fit_errors]


class Action:
    def __init__(self, device: int = 0):
        self.device = device

    def Name(self) -> str:
        return "allocate"

    def Initialize(self) -> None:
        engine.init(self.device)

    def Execute(self, ssn: Session) -> None:
        ssn.replay(ssn.device_session(self.device).allocate())

    def UnInitialize(self) -> None:
        pass


def New(device: int = 0) -> Action:
    return Action(device)
