"""Reference goldens transcribed late in the round, on the CUDA path (this file sorts last on purpose: the suite runs with -x)."""
import numpy as np
import pytest

from tests.golden import reference_cases as G

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("enable,case", G.priority_preempt_cases(), ids=lambda c: getattr(c, "Name", str(c))[:50])
def test_priority_plugin_preempt_goldens_gpu(enable, case, oracle_engine):
    """plugins/priority/priority_test.go:62-165 TestPreempt (allocate + preempt, the priority plugin alone): the reference's
    expectations and every statement of the oracle."""
    from volcano_b200 import engine
    engine.init(0)
    snap = case.RegisterSession(G.priority_preempt_tiers(enable), actions=("allocate", "preempt"))
    case.Run(engine.gpu_engine)
    assert case.CheckAll() is None, case.CheckAll()
    ref = oracle_engine(snap)
    res = case.result
    assert np.array_equal(res.decisions, ref.decisions) and np.array_equal(res.visits, ref.visits)
    for f in ("task", "node", "kind", "visit"):
        assert np.array_equal(res.preempt.decisions[f], ref.preempt.decisions[f]), f
    assert np.array_equal(res.preempt.visits, ref.preempt.visits)
