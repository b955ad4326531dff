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


def _run_workflow(tmp_path, cfg_path, ranks, tag, port):
    out = str(tmp_path / ("out_%s.npz" % tag))
    worker = os.path.join(ROOT, "tests", "helpers", "dp_run_task_worker.py")
    if ranks == 1:
        cmd = [sys.executable, worker, cfg_path, out]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(ranks),
               "--master-addr", "127.0.0.1", "--master-port", str(port), worker, cfg_path, out]
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=420, env=env)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    import numpy as np
    return np.load(out)


def _write_cfg(tmp_path, name, run_folder, mutate):
    import yaml
    with open(os.path.join(ROOT, "configs", name)) as f:
        cfg = yaml.safe_load(f)
    cfg["run"]["run_folder"] = str(tmp_path / run_folder)
    mutate(cfg["run"])
    p = tmp_path / (run_folder + "_" + name)
    with open(p, "w") as f:
        yaml.safe_dump(cfg, f)
    return str(p)


@pytest.mark.gpu
def test_run_task_data_parallel_validation_labels_bit_exact(tmp_path):
    """`torchrun --nproc-per-node 2 run_task.py cfg.yml` (validation): every rank forwards its share of each batch's
    videos, the clip logits are gathered in batch order and fused on the host -- video logits and pooled labels are
    bit-identical to the 1-rank run (SURVEY 8e parity rule; val.py:91-110)."""
    import numpy as np

    def small_val(run):
        d = run["data"]["synthetic-val"]
        d["num_items"] = 7
        d["clips_per_video"] = [2, 1, 3, 1, 2, 1, 2]
        d["num_frames_per_clip"] = 2
        run["val"]["batch_size"] = 3
    one = _run_workflow(tmp_path, _write_cfg(tmp_path, "config5_lrcn_val.yml", "r1", small_val), 1, "val1", 0)
    two = _run_workflow(tmp_path, _write_cfg(tmp_path, "config5_lrcn_val.yml", "r2", small_val), 2, "val2", 29541)
    assert one["logits"].shape == (7, 101)
    assert np.array_equal(one["logits"], two["logits"])  # batch independence of the forward pass: bit for bit
    assert np.array_equal(one["logits"].argmax(1), two["logits"].argmax(1))
    assert np.array_equal(one["labels"], two["labels"]) and float(one["acc"]) == float(two["acc"])


@pytest.mark.gpu
def test_run_task_data_parallel_training_matches_single_rank(tmp_path):
    """The same for training: 2 ranks on shards of every batch (whole videos per rank, unequal shards included) end
    with the variables of the 1-rank run up to the summation order of the gradient all-reduce."""
    import numpy as np

    def small_train(run):
        d = run["data"]["synthetic-train"]
        d["num_items"] = 6
        d["clips_per_video"] = [1, 2, 1, 1, 2, 1]
        d["num_frames_per_clip"] = 2
        run["train"]["batch_size"] = 3
        run["train"]["epochs"] = 1
        run["train"]["dropout_keep_prob"] = 0.0
        run["train"]["base_lr"] = 0.01
    one = _run_workflow(tmp_path, _write_cfg(tmp_path, "config2_lrcn_train.yml", "t1", small_train), 1, "tr1", 0)
    two = _run_workflow(tmp_path, _write_cfg(tmp_path, "config2_lrcn_train.yml", "t2", small_train), 2, "tr2", 29543)
    keys = [k for k in one.files if k.startswith("sd|") and k != "sd|global_step"]
    assert keys and int(one["sd|global_step"]) == int(two["sd|global_step"]) == 2
    import vlb200  # noqa: F401
    from vlb200 import engine as E
    init = E.init_variables(E.EngineConfig(workflow="lrcn", fpc=2, clip_norm=10), seed=1234)
    worst = 0.0
    for k in keys:
        name = k[3:].replace("|", "/")
        upd = np.abs(one[k] - init[name]).max() + 1e-12
        worst = max(worst, float(np.abs(one[k] - two[k]).max() / upd))
    print("dp2 run_task training: worst update rel err %.3e over %d variables" % (worst, len(keys)))
    assert worst < 5e-2
