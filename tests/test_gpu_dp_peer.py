"""Data-parallel gradient exchange over NVLink peer memory (csrc/dp_peer.cu), world size 2 on one node: the library's own
all-reduce gives NCCL's sums, a data-parallel iteration is one CUDA graph, and the replicas stay bit-identical.  Needs two GPUs
(skipped on a one-GPU box; the host-side logic of the exchange is covered by the gloo tests)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_peer_allreduce_matches_nccl_and_replicas_stay_identical():
    env = dict(os.environ, DP_CHECK_ITERS="5", MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "dp_peer_check.py")]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=280)
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert p.returncode == 0 and lines, p.stdout[-2000:] + p.stderr[-2000:]
    res = json.loads(lines[-1])
    assert res["ok"] and res["peer_exchange"] and res["peer"]["one_graph"] and res["peer"]["replicas_bit_identical"]
    assert res["peer"]["barrier_timeouts"] == 0
