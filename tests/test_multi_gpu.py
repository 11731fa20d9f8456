"""Two ranks on two GPUs of one box (skipped on a single-GPU box): ONE allocate session whose node axis is cut over both
GPUs (peer-mapped mailbox / ring, vc_comm_*) must return the decisions of the same session on one GPU, and the
node-sharded dense pass (NCCL MAX all-reduce + all-gather fold) must return the single-shard per-task best."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _torchrun(script, *args, nproc=2, timeout=240):
    import torch
    if torch.cuda.device_count() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", script), *args]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    return json.loads(line)


@pytest.mark.parametrize("cfg", ["tiny", "small", "small_roles", "tiny_fut", "small_soft", "small_fut_soft"])
def test_one_session_across_two_gpus(cfg):
    out = _torchrun("multi_gpu_commit.py", cfg, "2")
    assert out["world"] == 2 and out["identical_to_one_gpu"] and out["placed"] > 0


def test_node_sharded_dense_pass_two_gpus():
    out = _torchrun("dense_sharded_check.py", "small")
    assert out["world"] == 2 and out["identical"]
