"""Host-side integer / numpy logic against golden vectors produced by the REFERENCE's own functions
(tests/golden/make_golden.py ran serialize.py, utils_.py, train.py and val.py from /root/reference under a stub
tensorflow module).  Everything here is bit-exact."""
import json
import os
import random
import types

import numpy as np
import pytest

import vlb200  # noqa: F401
from vlb200 import clips as C
from vlb200.defs import defs
from vlb200.train import precompute_learning_rates
from vlb200.utils import labels_to_one_hot
from vlb200.val import Validation

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "reference_host_golden.json")) as fh:
    GOLD = json.load(fh)
VAL = np.load(os.path.join(HERE, "golden", "reference_val_golden.npz"))


class _Logger(object):
    def add_to_log_storage(self, *a):
        pass


@pytest.mark.parametrize("case", GOLD["clips"], ids=lambda c: "%s-s%s-n%d-f%d-k%d" % (
    c["gen"], c["seed"], c["n"], c["fpc"], c["num"]))
def test_clip_generators_bit_exact(case):
    fn = getattr(C, case["gen"])
    st = types.SimpleNamespace(num_frames_per_clip=case["fpc"], clip_offset_or_num=case["num"],
                               generation_error=defs.generation_error.compromise, logger=_Logger())
    random.seed(case["seed"])
    if isinstance(case["clips"], dict):
        with pytest.raises(Exception):
            fn(list(range(case["n"])), st, "/tmp/video_x")
        return
    assert fn(list(range(case["n"])), st, "/tmp/video_x") == case["clips"]


def test_clip_generator_edge_cases():
    st = types.SimpleNamespace(num_frames_per_clip=4, clip_offset_or_num=2,
                               generation_error=defs.generation_error.abort, logger=_Logger())
    with pytest.raises(Exception):
        C.get_random_clips([], st, "/tmp/empty")          # no frames
    with pytest.raises(Exception):
        C.get_random_clips([0, 1], st, "/tmp/short")      # abort strategy on a short video
    st.generation_error = defs.generation_error.report
    assert C.get_sequential_clips([0, 1], st, "/tmp/short") == []
    st.generation_error = defs.generation_error.compromise
    assert C.get_random_clips([0, 1], st, "/tmp/short") == [[0, 0, 0, 1], [0, 0, 0, 1]]  # left padded with frame 0
    with pytest.raises(Exception):
        C.get_random_frames([0, 1, 2], st, "/tmp/x")      # broken in the reference -> refused here


@pytest.mark.parametrize("case", GOLD["onehot"])
def test_labels_to_one_hot(case):
    got = labels_to_one_hot(case["labels"], case["num_classes"])
    assert got.dtype == np.int32 and got.tolist() == case["onehot"]


def test_labels_to_one_hot_rejects_out_of_range():
    with pytest.raises(Exception):
        labels_to_one_hot([[7]], 7)


@pytest.mark.parametrize("case", GOLD["lr"], ids=lambda c: str(c["decay"]))
def test_learning_rate_table_bit_exact(case):
    got = precompute_learning_rates(case["base_lr"], case["decay"], case["num_batches"], case["epochs"])
    assert got == case["lrs"]  # Python floats, same operation order -> identical doubles


@pytest.mark.parametrize("method", ["avg", "last"])
def test_clip_to_video_fusion_and_accuracy(method, tmp_path):
    st = types.SimpleNamespace(num_classes=101, run_id="t", run_folder=str(tmp_path), val=None)
    v = Validation(st, use_device=False)
    v.process_validation_logits(VAL["logits"], VAL["labels"], [int(c) for c in VAL["cpv"]], method)
    assert np.array_equal(v.item_logits, VAL["video_logits_" + method])   # np.mean(axis=0): bit exact
    assert np.array_equal(v.item_labels, VAL["video_labels_" + method])
    assert v.get_chunk_accuracy(v.item_logits, v.item_labels) == float(VAL["accuracy_" + method])
    assert v.get_accuracy() == float(VAL["get_accuracy_" + method])


def test_clip_fusion_rejects_other_methods(tmp_path):
    st = types.SimpleNamespace(num_classes=101, run_id="t", run_folder=str(tmp_path), val=None)
    v = Validation(st, use_device=False)
    with pytest.raises(Exception):
        v.process_validation_logits(VAL["logits"], VAL["labels"], [int(c) for c in VAL["cpv"]], "maximum")


def test_chunked_accuracy_is_unweighted_mean(tmp_path):
    """val.py:197: the final accuracy is the plain mean of per-chunk accuracies."""
    st = types.SimpleNamespace(num_classes=3, run_id="t", run_folder=str(tmp_path),
                               val=types.SimpleNamespace(logits_save_interval=2))
    v = Validation(st, use_device=False)
    logits = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1]], np.float32)
    labels = np.array([[1, 0, 0], [1, 0, 0], [0, 0, 1]], np.int32)
    v.process_validation_logits(logits[:2], labels[:2], [1, 1], "avg")
    v.save_validation_logits_chunk()           # chunk 0: accuracy 0.5
    v.process_validation_logits(logits[2:], labels[2:], [1], "avg")
    v.save_validation_logits_chunk(save_all=True)   # chunk 1: accuracy 1.0
    assert v.get_accuracy() == 0.75
