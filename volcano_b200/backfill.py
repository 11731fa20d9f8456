"""framework.Action mirror of the backfill action (actions/backfill/backfill.go:40-116).

    action = backfill.New()   ->  volcano_b200.backfill.New()
    action.Execute(ssn)       ->  vc_backfill_run on the device-side session the cycle's allocate action left, then
                                  Session.Allocate replayed for every placed BestEffort task

Run after `volcano_b200.action.New()` in the same `TestCommonStruct.Run([...])` list, like the reference's configured
action order "allocate, backfill" (pkg/scheduler/util.go:38-51)."""
from __future__ import annotations

from . import engine
from .action import Session


class Action:
    def __init__(self, device: int = 0):
        self.device = device

    def Name(self) -> str:
        return "backfill"

    def Initialize(self) -> None:
        engine.init(self.device)

    def Execute(self, ssn: Session) -> None:
        if ssn.snapshot.B == 0:  # no BestEffort pending task: pickUpPendingTasks returns nothing
            return
        ssn.replay_backfill(ssn.device_session(self.device).backfill())

    def UnInitialize(self) -> None:
        pass


def New(device: int = 0) -> Action:
    return Action(device)
