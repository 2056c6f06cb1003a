"""Single-verdict mode over real NCCL ranks: one process per GPU under torchrun (two ranks when the box has two GPUs, one
otherwise -- NCCL refuses two ranks on one device), tools/nccl_single_verdict.py does the checking."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_single_verdict_mode_over_nccl_ranks():
    import torch
    world = min(2, torch.cuda.device_count())
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "nccl_single_verdict.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    rep = json.loads(line)
    assert rep["ok"] and rep["world"] == world and len(rep["cases"]) == 2
    for case in rep["cases"]:
        for row in case["per_rank"]:
            assert row[3] == 1 and row[4] == 0 and row[5] == 1      # accept, tampered rejects, host surface accepts
