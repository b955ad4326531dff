"""Runs the REFERENCE's own `run_task.main` (from /root/reference, unmodified) through vlb200's tensorflow stand-in.

    python reference_dropin_driver.py <workdir> <phase: train|resume|val> [--cuda]

What is replaced is exactly what INTEGRATION.md section B tells a maintainer to replace:
    import tensorflow            ->  vlb200.tfshim (installed in sys.modules before the reference is imported)
    models.model.Model, train.Train  ->  vlb200.compat.Model, vlb200.compat.Train
Everything else -- settings_.Settings, feeder.Feeder, dataset_.Dataset (TFRecord reading, crop / mean subtraction per
frame), val.Validation, do_train / do_test, feeder.save / init_saveload / resume_snap -- is the reference's code.

The other patches below are ENVIRONMENT shims for a 2017 code base under Python 3.12 / numpy 2 / PyYAML 6 / scipy 1.18
(SURVEY 8c recipe); they would be unnecessary in the reference's own environment.

Without --cuda the session's engine is the CPU stand-in of tests/helpers/fake_engine.py (control flow only); with
--cuda it is the real Engine.  Prints one JSON line with what the test asserts on."""
import collections
import collections.abc
import json
import os
import pickle
import sys

import numpy as np
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REFERENCE = "/root/reference"


def environment_shims():
    collections.Iterable = collections.abc.Iterable  # parse_opts.py:3
    import scipy.misc  # dataset_.py:9 imports imread / imresize / imsave (removed from scipy long ago; unused for tfrecords)
    for n in ("imread", "imresize", "imsave"):
        setattr(scipy.misc, n, None)
    fromstring = np.fromstring  # dataset_.py:131 np.fromstring(bytes): binary mode removed in numpy 2
    np.fromstring = lambda s, dtype=float, **k: (np.frombuffer(s, dtype=dtype).copy()
                                                 if isinstance(s, (bytes, bytearray)) else fromstring(s, dtype=dtype, **k))
    load = yaml.load  # settings_.py:388 yaml.load(f) without a Loader (PyYAML >= 6 requires one)
    yaml.load = lambda f, Loader=None: load(f, Loader=Loader or yaml.SafeLoader)


def write_dataset(work, name, cpv, fpc, shape, classes, seed):
    """A dataset in the reference's serialized layout (serialize.py:138-151,246-256,664): `<inp>` (the paths file),
    `<inp>.tfrecord`, `<inp>.tfrecord.size`."""
    from vlb200 import tfrecord
    rng = np.random.default_rng(seed)
    base = os.path.join(work, name)
    if os.path.exists(base + ".tfrecord"):
        return base
    with open(base, "w") as f:
        for v in range(len(cpv)):
            f.write("video_%d %d\n" % (v, v % classes))
    with open(base + ".tfrecord", "wb") as f:
        for v, c in enumerate(cpv):
            for _ in range(c * fpc):
                img = rng.integers(0, 256, size=shape, dtype=np.uint8)
                tfrecord.write_record(f, tfrecord.serialize_frame(img, [v % classes]))
    tfrecord.write_size_file(base + ".tfrecord.size", cpv, fpc)
    return base


def config(work, phase, shape, classes):
    run_folder = os.path.join(work, "runfolder")
    cpv, fpc = [2, 1, 2, 1], 2
    train_base = write_dataset(work, "train_data", cpv, fpc, shape, classes, 0)
    val_base = write_dataset(work, "val_data", [2, 1, 3, 1, 2], fpc, shape, classes, 1)
    run = {
        "resume_file": None if phase == "train" else "latest", "run_folder": run_folder, "run_id": "dropin",
        "phase": "defs.phase.val" if phase == "val" else "defs.phase.train",
        "logging": {"save_freq_per_epoch": 1, "level": "logging.INFO", "tensorboard_folder": "tb", "print_tensors": False,
                    "email_notify": None},
        "network": {"num_classes": classes, "pipelines": [{"lrcn": {
            "input": "defs.dataset_tag.main", "representation": "defs.representation.dcnn", "frame_encoding_layer": "fc7",
            "classifier": "defs.classifier.lstm", "lstm_params": [64, 1, "defs.fusion_method.avg"]}}]},
        "train": {"batch_size": 2, "epochs": 1 if phase == "train" else 3, "optimizer": "defs.optim.sgd", "base_lr": 0.001,
                  "lr_mult": "None", "lr_decay": ["defs.decay.exp", "defs.periodicity.interval", 2, 0.5], "clip_norm": 10,
                  "dropout_keep_prob": 0.5},
        "val": {"batch_size": 2, "logits_save_interval": 0, "clip_fusion": ["defs.fusion_type.late", "defs.fusion_method.avg"]},
        "data": {
            "train-data": {"data_path": train_base, "data_format": "defs.data_format.tfrecord", "phase": "defs.phase.train",
                           "tag": "defs.dataset_tag.main", "image_shape": str(tuple(shape)),
                           "mean_image": "[99.197148, 105.293620, 109.503945]"},
            "val-data": {"data_path": val_base, "data_format": "defs.data_format.tfrecord", "phase": "defs.phase.val",
                         "tag": "defs.dataset_tag.main", "image_shape": str(tuple(shape)),
                         "mean_image": "[99.197148, 105.293620, 109.503945]"}},
    }
    path = os.path.join(work, "cfg_%s.yml" % phase)
    with open(path, "w") as f:
        yaml.safe_dump({"run": run}, f)
    return path, run_folder


def main():
    work, phase = sys.argv[1], sys.argv[2]
    cuda = "--cuda" in sys.argv
    os.makedirs(work, exist_ok=True)
    environment_shims()
    import vlb200  # noqa: F401
    from vlb200 import compat, tfshim
    tfshim.install()                       # (1) `import tensorflow` of the reference's files resolves to the stand-in
    sys.path.insert(0, REFERENCE)
    import run_task as ref                 # the reference's run_task.py, as it lies under /root/reference
    ref.Model, ref.Train = compat.Model, compat.Train   # (2) = editing its two import lines
    engines = []
    if not cuda:
        from helpers.fake_engine import FakeEngine

        def factory(cfg, max_clips, device, rank, world):
            engines.append(FakeEngine(cfg, max_clips))
            return engines[-1]
        tfshim.Session.engine_factory = staticmethod(factory)
    shape = (227, 227, 3) if cuda else (67, 67, 3)
    cfg_path, run_folder = config(work, phase, shape, 5)
    ref.main(cfg_path)
    sess = tfshim.get_default_graph().session
    eng = sess.engine
    out = {"phase": phase, "global_step": int(eng.global_step), "engine": type(eng).__name__,
           "reference_modules": sorted(m for m in ("run_task", "settings_", "feeder", "dataset_", "val", "utils_", "defs_")
                                       if getattr(sys.modules.get(m), "__file__", "").startswith(REFERENCE))}
    if hasattr(eng, "calls"):
        out["calls"] = [list(c) for c in eng.calls]
    ck = os.path.join(run_folder, "checkpoints")
    if os.path.isdir(ck):
        out["checkpoint_files"] = sorted(os.listdir(ck))
        snaps = {}
        for f in out["checkpoint_files"]:
            if f.endswith(".snap"):
                with open(os.path.join(ck, f), "rb") as fh:
                    snaps[f] = pickle.load(fh)
        out["snaps"] = snaps
        with open(os.path.join(ck, "checkpoint")) as fh:
            out["index_first_line"] = fh.readline().strip()
    acc = [f for f in os.listdir(run_folder) if f.startswith("accuracy_")]
    if acc:
        with open(os.path.join(run_folder, acc[0])) as fh:
            out["accuracy"] = float(fh.read())
    print("DROPIN_RESULT " + json.dumps(out))


if __name__ == "__main__":
    main()
