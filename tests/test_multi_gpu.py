"""GPU, two or more devices on the box (skipped otherwise): the multi-rank paths under torchrun -- the peer-memory
all-reduce against NCCL and the rank-ordered sum, and row-sharded training (host-driven and device-resident) against the
single-GPU trajectory."""
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _torchrun(script, nproc, env=None, timeout=600):
    import os

    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", "29611", str(ROOT / "scripts" / script)]
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run(cmd, cwd=ROOT, env=e, capture_output=True, text=True, timeout=timeout)


@pytest.mark.timeout(900)
@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs on one box")
def test_two_rank_exchange_and_sharded_training():
    r = _torchrun("p2p_check.py", 2)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "exact_vs_rank_order_sum_and_close_to_nccl=True" in r.stdout, r.stdout[-2000:]
    r = _torchrun("multi_gpu_train_check.py", 2, {"MG_ROWS": "400000"})
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert '"ok": true' in r.stdout
