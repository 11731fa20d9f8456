import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_engine():
    """Snapshot -> AllocateResult through the CPU oracle (test infrastructure)."""
    from oracle.pyoracle import OracleSession
    from volcano_b200.uthelper import AllocateResult

    def run(snap, threads=1):
        s = OracleSession(snap, threads=threads)
        dec, vis, fe = s.allocate()
        res = AllocateResult(dec, vis, fe)
        if snap.B > 0 and "backfill" in snap.actions:
            res.backfill = AllocateResult(*s.backfill())
        if snap.hn_job_soft is not None and snap.hn_job_soft.any():
            import numpy as np
            from oracle import pyoracle
            res.job_allocated_hypernodes = np.array([pyoracle.lib().vco_job_allocated_hypernode(s.h, j) for j in range(snap.J)],
                                                    np.int32)
        s.close()
        return res

    return run
