"""Generate golden vectors by running the REFERENCE's own host-side functions (pure Python / numpy) from
/root/reference under a stub `tensorflow` module (SURVEY 8c recipe).  Runs only in the build container
(the GPU box has no /root/reference); the outputs are committed next to this script:

    tests/golden/reference_host_golden.json    clip generators, one-hot, LR tables
    tests/golden/reference_val_golden.npz      clip->video fusion, accuracies

Usage: python tests/golden/make_golden.py
"""
import collections
import collections.abc
import json
import os
import random
import sys
import tempfile
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    tf = types.ModuleType("tensorflow")
    sys.modules["tensorflow"] = tf
    import scipy.misc
    for n in ("imread", "imresize", "imsave"):
        setattr(scipy.misc, n, lambda *a, **k: None)
    collections.Iterable = collections.abc.Iterable
    sys.path.insert(0, REF)
    import defs_  # noqa
    import serialize  # noqa
    import train  # noqa
    import utils_  # noqa
    import val  # noqa
    return defs_, serialize, train, utils_, val


class _Logger(object):
    def add_to_log_storage(self, *a):
        pass


def main():
    defs_, serialize, train, utils_, val = import_reference()
    defs = defs_.defs
    golden = {"clips": [], "onehot": [], "lr": []}

    # ---- clip generators (serialize.py:293-378) ----
    for gen_name in ("get_random_clips", "get_sequential_clips"):
        fn = getattr(serialize, gen_name)
        for seed in (0.5, 1234.0):
            for n in (15, 20, 40, 5):
                for fpc in (3, 16):
                    for num in (1, 3, 5):
                        st = types.SimpleNamespace(num_frames_per_clip=fpc, clip_offset_or_num=num,
                                                   generation_error=defs.generation_error.compromise, logger=_Logger())
                        if gen_name == "get_sequential_clips" and n < fpc:
                            continue  # the reference extends frame list with random dups, then yields nothing useful
                        random.seed(seed)
                        try:
                            clips = fn(list(range(n)), st, "/tmp/video_x")
                        except Exception as ex:  # noqa
                            clips = {"error": str(ex)}
                        golden["clips"].append(dict(gen=gen_name, seed=seed, n=n, fpc=fpc, num=num, clips=clips))

    # ---- labels_to_one_hot (utils_.py:160-169) ----
    for labels, c in (([[0], [5], [2]], 7), ([[1, 3], [0]], 5), ([[100]], 101)):
        golden["onehot"].append(dict(labels=labels, num_classes=c,
                                     onehot=utils_.labels_to_one_hot(labels, c).tolist()))

    # ---- LR tables (train.py:50-109) ----
    tr = train.Train.__new__(train.Train)
    tmp = tempfile.mkdtemp()
    cases = [
        (0.001, None, 7, 3),
        (0.001, (defs.decay.exp, defs.periodicity.interval, 1000, 0.96), 900, 3),
        (0.01, (defs.decay.staircase, defs.periodicity.interval, 4, 0.5), 10, 2),
        (0.01, (defs.decay.exp, defs.periodicity.drops, 4, 0.5), 10, 3),
        (0.1, (defs.decay.staircase, defs.periodicity.drops, 3, 0.1, 5), 11, 2),
    ]
    for base_lr, decay, nb, epochs in cases:
        st = types.SimpleNamespace(train=types.SimpleNamespace(base_lr=base_lr, lr_decay=decay, epochs=epochs),
                                   run_folder=tmp, run_id="golden")
        lrs = tr.precompute_learning_rates(st, nb)
        golden["lr"].append(dict(base_lr=base_lr, decay=list(decay) if decay else None, num_batches=nb, epochs=epochs,
                                 lrs=[float(x) for x in lrs]))

    with open(os.path.join(OUT, "reference_host_golden.json"), "w") as fh:
        json.dump(golden, fh)

    # ---- clip -> video fusion and accuracy (val.py:158-203) ----
    rng = np.random.default_rng(42)
    cpv = [3, 1, 7, 25, 2, 4]
    c = 101
    logits = (rng.standard_normal((sum(cpv), c)) * 20).astype(np.float32)
    labels = np.zeros((sum(cpv), c), np.int32)
    off = 0
    for i, n in enumerate(cpv):
        labels[off:off + n, rng.integers(0, c)] = 1
        off += n
    out = {"logits": logits, "labels": labels, "cpv": np.array(cpv, np.int32)}
    for method in (defs.fusion_method.avg, defs.fusion_method.last):
        v = val.Validation.__new__(val.Validation)
        v.item_logits = np.zeros([0, c], np.float32)
        v.item_labels = np.zeros([0, c], np.float32)
        lg, lb = logits, labels
        for n in cpv:
            v.apply_clip_fusion(lg, n, lb, method)
            lg, lb = lg[n:], lb[n:]
        out["video_logits_" + method] = v.item_logits.astype(np.float32)
        out["video_labels_" + method] = v.item_labels
        out["accuracy_" + method] = np.float64(v.get_chunk_accuracy(v.item_logits, v.item_labels))
        # chunked accuracy (val.py:174-198): unweighted mean of chunk accuracies
        v.validation_logits_save_counter = 0
        v.validation_logits_save_interval = None
        out["get_accuracy_" + method] = np.float64(v.get_accuracy())
    np.savez(os.path.join(OUT, "reference_val_golden.npz"), **out)
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
