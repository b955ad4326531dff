"""N > 1 on real GPUs: the 2-rank data-parallel step == the single-rank step on the whole batch.  With two GPUs the
ranks exchange through NCCL; on a one-GPU box both ranks share cuda:0 and exchange through gloo (NCCL refuses two ranks
on one device), so the test runs everywhere."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_dp2_train_step_matches_single_rank():
    import torch
    assert torch.cuda.device_count() >= 1
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "helpers", "dp_worker.py")]
    res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=240)
    sys.stdout.write(res.stdout[-2000:])
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
