"""A C program (tools/cdriver/vcalloc_driver.c) as the consumer of include/vcalloc.h: no Python, no ctypes between the
session dump and the library. Without a GPU it must stop at vc_init with VC_ENODEV (no CPU path); on the GPU box its
decisions must be the ones the Python binding gets for the same session."""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fnv(h, data: bytes):
    for b in data:
        h ^= b
        h = (h * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def _hash_result(res):
    h = 1469598103934665603
    dec = np.stack([res.decisions[f] for f in ("task", "node", "kind", "visit")], axis=1).astype("<i4")
    h = _fnv(h, dec.tobytes())
    vis = np.stack([res.visits[f] for f in ("job", "outcome", "first_op", "n_ops")], axis=1).astype("<i4")
    return _fnv(h, vis.tobytes())


@pytest.fixture(scope="module")
def driver():
    import __graft_entry__ as g
    from volcano_b200.build import build_lib
    build_lib()
    return g.build_cdriver()


def test_c_driver_builds_and_refuses_without_gpu(driver, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from volcano_b200.synth import make_snapshot
    path = str(tmp_path / "tiny.bin")
    make_snapshot("tiny").dump(path)
    r = subprocess.run([driver, path], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU path" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", ["tiny", "small", "small_roles"])
def test_c_driver_matches_python_binding(driver, cfg, tmp_path):
    from volcano_b200 import engine
    from volcano_b200.synth import make_snapshot
    snap = make_snapshot(cfg)
    path = str(tmp_path / f"{cfg}.bin")
    snap.dump(path)
    r = subprocess.run([driver, path], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = json.loads(r.stdout.strip().splitlines()[-1])
    res = engine.gpu_engine(snap)
    assert out["decisions"] == len(res.decisions) and out["visits"] == len(res.visits) and out["fit_errors"] == len(res.fit_errors)
    assert out["hash"] == "%016x" % _hash_result(res)
