"""The drop-in, proven on the reference's own code: `/root/reference/run_task.py:main` runs UNMODIFIED (Settings, Feeder,
Dataset, Validation, do_train / do_test, feeder.save / init_saveload / resume_snap) with `tensorflow` replaced by
vlb200.tfshim and the two graph builders (Model, Train) by vlb200.compat -- exactly the change INTEGRATION.md section B
describes.  Runs where /root/reference exists (this container, CPU: the session's engine is the stand-in of
tests/helpers/fake_engine.py, so what is exercised is the whole seam: placeholders, feed dicts, fetch lists, saver files,
the .snap / checkpoint index the reference reads back).  The same seam over the real CUDA engine:
tests/test_gpu_workflow.py::test_tfshim_session_drives_the_real_engine."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "tests", "helpers", "reference_dropin_driver.py")

pytestmark = pytest.mark.skipif(not os.path.exists("/root/reference/run_task.py"),
                                reason="the reference tree is only present in the build container")


def _run(work, phase):
    res = subprocess.run([sys.executable, DRIVER, str(work), phase], capture_output=True, text=True, timeout=600,
                         cwd=str(work))
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("DROPIN_RESULT ")]
    assert line, res.stdout[-2000:]
    return json.loads(line[-1][len("DROPIN_RESULT "):]), res.stdout + res.stderr


def test_reference_run_task_trains_resumes_and_validates_through_the_shim(tmp_path):
    work = tmp_path / "w"
    work.mkdir()
    # ---- train: 4 videos, batch 2 -> 2 batches, 1 epoch; lr_decay [exp, interval, 2, 0.5] ----
    # (ONE checkpoint before the resume: the reference's init_saveload takes get_run_checkpoints(...)[-1], whose sort key
    # `"_".join(x.split()[2:])` is empty for every path without blanks, i.e. "latest" is os.listdir order - utils_.py:223-230)
    out, log = _run(work, "train")
    assert out["reference_modules"] == ["dataset_", "defs_", "feeder", "run_task", "settings_", "utils_", "val"]
    steps = [c for c in out["calls"] if c[0] == "train"]
    assert [c[1] for c in steps] == [3, 3]                          # clips per batch: whole videos (cpv 2,1 | 2,1)
    assert [c[2] for c in steps] == [0.001, 0.001]                  # lr = table[global_step] (train.py:129-132)
    assert out["global_step"] == 2
    assert list(out["snaps"].values()) == [[2, 0, 2]]               # feeder.save: [batch_index, epoch_index, global_step]
    assert all(any(f.endswith(ext) for f in out["checkpoint_files"]) for ext in (".npz", ".meta", ".index", ".snap"))
    assert out["index_first_line"].startswith('model_checkpoint_path: "') and "gs_2.graph-2" in out["index_first_line"]
    assert "Learning rate 0.00100000, global step: 2" in log        # the reference's own log line (run_task.py:51)
    assert "Epoch [1] training run complete." in log
    # ---- resume `latest` with 3 epochs: resume_snap + init_saveload + saver.restore are the reference's ----
    out, log = _run(work, "resume")
    assert "Resumed epoch [1] is already complete." in log          # run_task.py:68
    steps = [c for c in out["calls"] if c[0] == "train"]
    assert [c[2] for c in steps] == [0.0005, 0.0005, 0.00025, 0.00025] and out["global_step"] == 6
    assert any("ep_3_btch_2_gs_6.graph-6.npz" in f for f in out["checkpoint_files"])
    # ---- validation of the resumed checkpoint: the reference's Validation fuses clips to videos ----
    out, log = _run(work, "val")
    assert [c[1] for c in out["calls"] if c[0] == "forward"] == [3, 4, 2]  # cpv 2,1 | 3,1 | 2
    assert 0.0 <= out["accuracy"] <= 1.0
    assert "Validation run complete" in log
