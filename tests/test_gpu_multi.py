"""N > 1 on real GPUs (skipped on boxes with one GPU): NCCL data-parallel step == single-rank step on the whole batch."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_dp2_train_step_matches_single_rank():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "helpers", "dp_worker.py")]
    res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=240)
    sys.stdout.write(res.stdout[-2000:])
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
