"""GPU tier, two or more GPUs (skipped on a one-GPU box): the data-parallel critic update and the
data-parallel observation normaliser against the reference arithmetic on the CONCATENATED batch
(SURVEY 8e), through the fused exchange (symmetric memory, no NCCL on the path) and through ncclAllReduce.
The check itself is tools/dp_oracle_check.py, launched under torch.distributed.run."""
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


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("dp", ["fused", "nccl"])
@pytest.mark.timeout(600)
def test_n_rank_update_equals_oracle_on_concatenated_batch(dp):
    n = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "dp_oracle_check.py"), "--dp", dp]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=580, cwd=ROOT)
    print(out.stdout[-3000:])
    assert "DP ORACLE CHECK PASSED" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]
