"""-m gpu, needs >= 2 GPUs (skipped below): the N-replica step programs on hardware.  tools/replica_check.py under torchrun:
A. R identical replicas == one replica bit for bit, B. distinct replicas keep bit-identical policies (and Adam state) on
all ranks, C. one step against the oracle's R-replica emulation (sum of per-replica policy gradients / R before Adam,
SURVEY.md 8e) -- in the exact mode (0) and in the production mode (3 = 3xTF32)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("precision", [0, 3])
def test_replica_exchange_on_hardware(precision):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (have %d)" % n)
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    env = dict(os.environ, ILSW_CHECK_PRECISION=str(precision))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "replica_check.py")]
    p = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    assert "replica_check OK world=%d precision=%d" % (world, precision) in p.stdout, p.stdout[-2000:]
