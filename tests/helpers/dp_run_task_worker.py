"""Worker of tests/test_gpu_multi.py::test_run_task_data_parallel_*: `torchrun ... dp_run_task_worker.py cfg.yml out.npz`
runs the product workflow (run_task.main) on every rank; rank 0 stores what the test compares with the 1-rank run."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import vlb200  # noqa: E402,F401
from vlb200 import run_task  # noqa: E402

cfg, out = sys.argv[1], sys.argv[2]
captured = {}
orig_train, orig_test = run_task.do_train, run_task.do_test


def do_train(settings, train, feeder, engine, dp=None):
    orig_train(settings, train, feeder, engine, dp)
    captured["sd"] = engine.state_dict()


def do_test(settings, val, feeder, engine, dp=None):
    acc = orig_test(settings, val, feeder, engine, dp)
    captured["logits"] = val.item_logits.copy()
    captured["labels"] = val.item_labels.copy()
    captured["acc"] = np.float64(acc)
    return acc


run_task.do_train, run_task.do_test = do_train, do_test
run_task.main(cfg)
if int(os.environ.get("RANK", "0")) == 0:
    flat = {}
    for k, v in captured.items():
        if isinstance(v, dict):
            for kk, vv in v.items():
                flat["sd|" + kk.replace("/", "|")] = vv
        else:
            flat[k] = v
    np.savez(out, **flat)
import torch.distributed as dist  # noqa: E402
if dist.is_initialized():
    dist.barrier()
    dist.destroy_process_group()
