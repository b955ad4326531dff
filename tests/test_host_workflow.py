"""Host workflow on the CPU (run_task loops, resume, weights_file, checkpoints, data-parallel sharding) with a
stand-in engine (tests/helpers/fake_engine.py): the arithmetic is covered by the GPU tests, the control flow here."""
import os
import pickle

import numpy as np
import pytest
import yaml

import vlb200  # noqa: F401
from vlb200 import checkpoint, run_task
from vlb200 import engine as E
from vlb200.settings import Settings
from vlb200.train import precompute_learning_rates
from vlb200.defs import defs

from helpers.fake_engine import FakeEngine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cfg(name, tmp_path, mutate):
    with open(os.path.join(ROOT, "configs", name)) as f:
        cfg = yaml.safe_load(f)
    cfg["run"]["run_folder"] = str(tmp_path / "run")
    mutate(cfg["run"])
    p = tmp_path / name
    with open(p, "w") as f:
        yaml.safe_dump(cfg, f)
    return str(p)


class _Factory(object):
    """engine_factory for run_task.main that keeps the engines it built; `crash_at` makes the engine of the first run
    die at that global step (a killed job)."""

    def __init__(self, crash_at=None):
        self.engines = []
        self.crash_at = crash_at

    def __call__(self, cfg, max_clips, device, rank, world):
        eng = FakeEngine(cfg, max_clips, "cpu", rank, world)
        if self.crash_at is not None:
            inner, crash_at = eng.train_step, self.crash_at

            def step(*a, **k):
                if eng.global_step >= crash_at:
                    raise KeyboardInterrupt("killed")
                return inner(*a, **k)
            eng.train_step = step
        self.engines.append(eng)
        return eng


def _small_train(epochs=3, lr_decay=None, optimizer="defs.optim.sgd", resume=None, weights=None):
    def mutate(run):
        d = run["data"]["synthetic-train"]
        d["num_items"] = 8
        d["num_frames_per_clip"] = 2
        d["image_shape"] = "(67, 67, 3)"
        run["train"]["batch_size"] = 4
        run["train"]["epochs"] = epochs
        run["train"]["optimizer"] = optimizer
        run["train"]["lr_decay"] = lr_decay or ["defs.decay.exp", "defs.periodicity.interval", 2, 0.5]
        run["resume_file"] = resume
        if weights:
            run["network"]["pipelines"][0]["lrcn"]["weights_file"] = weights
    return mutate


def test_resume_from_an_epoch_end_checkpoint_moves_on_to_the_next_epoch(tmp_path):
    """the periodic save at the end of an epoch stores batch_index == num_batches with the
    un-incremented epoch index.  Resuming from it must NOT train that epoch again: the reference restores the batch
    index, finds the loop exhausted, logs "Resumed epoch is already complete" and continues with the next epoch
    (run_task.py:66-69), so that the global step keeps indexing the learning-rate table correctly."""
    first = _Factory(crash_at=3)  # 2 batches per epoch: the job dies in epoch 2, after the epoch-1 checkpoint (gs 2)
    with pytest.raises(KeyboardInterrupt):
        run_task.main(_cfg("config2_lrcn_train.yml", tmp_path, _small_train()), engine_factory=first)
    folder = tmp_path / "run" / "checkpoints"
    names = checkpoint.read_index(str(folder))
    assert len(names) == 1 and "ep_1_btch_2_gs_2" in names[0]
    with open(names[0] + ".snap", "rb") as f:
        assert pickle.load(f) == [2, 0, 2]  # batch_index == num_batches, epoch index 0, global step 2
    second = _Factory()
    run_task.main(_cfg("config2_lrcn_train.yml", tmp_path, _small_train(resume="latest")), engine_factory=second)
    eng = second.engines[0]
    steps = [c for c in eng.calls if c[0] == "train"]
    table = precompute_learning_rates(0.001, ["exp", "interval", 2, 0.5], 2, 3)
    assert len(steps) == 4 and eng.global_step == 6 == 3 * 2  # epochs * num_batches in total, none repeated
    assert [s[2] for s in steps] == [float(x) for x in table[2:6]]  # LR of step g is table[g]


def test_resume_in_the_middle_of_an_epoch(tmp_path):
    def mutate(run):
        _small_train(epochs=2)(run)
        run["logging"]["save_freq_per_epoch"] = 2  # a save after every batch
    first = _Factory(crash_at=3)
    with pytest.raises(KeyboardInterrupt):
        run_task.main(_cfg("config2_lrcn_train.yml", tmp_path, mutate), engine_factory=first)

    def mutate2(run):
        mutate(run)
        run["resume_file"] = "latest"
    second = _Factory()
    run_task.main(_cfg("config2_lrcn_train.yml", tmp_path, mutate2), engine_factory=second)
    eng = second.engines[0]
    assert len([c for c in eng.calls if c[0] == "train"]) == 1 and eng.global_step == 4  # only batch 2 of epoch 2 was left


def test_weights_file_initialises_the_encoder(tmp_path):
    """alexnet.py:50-52,69-71: conv1..fc7 come from the pickled dict layer -> [W, b] of bvlc_alexnet.npy; fc8 has 1000
    classes there and keeps its initialisation (here the pipeline stops at fc7 anyway)."""
    cfg = E.EngineConfig(workflow="lrcn", fpc=2, height=67, width=67)
    shapes = dict(E.variable_shapes(cfg))
    rng = np.random.default_rng(0)
    blob = {}
    for layer in ("conv1", "conv2", "conv3", "conv4", "conv5", "fc6", "fc7"):
        blob[layer] = [rng.standard_normal(shapes["dcnn/%sW" % layer]).astype(np.float32),
                       rng.standard_normal(shapes["dcnn/%sb" % layer]).astype(np.float32)]
    blob["fc8"] = [np.zeros((4096, 1000), np.float32), np.zeros(1000, np.float32)]
    path = str(tmp_path / "bvlc_alexnet.npy")
    np.save(path, blob, allow_pickle=True)
    fac = _Factory()
    run_task.main(_cfg("config2_lrcn_train.yml", tmp_path, _small_train(epochs=1, weights=path)), engine_factory=fac)
    eng = fac.engines[0]
    for layer in ("conv1", "conv2", "conv3", "conv4", "conv5", "fc6", "fc7"):
        assert np.array_equal(eng.vars["dcnn/%sW" % layer], blob[layer][0]), layer
        assert np.array_equal(eng.vars["dcnn/%sb" % layer], blob[layer][1]), layer
    missing = str(tmp_path / "nope.npy")
    with pytest.raises(Exception, match="does not exist"):
        run_task.main(_cfg("config2_lrcn_train.yml", tmp_path, _small_train(epochs=1, weights=missing)),
                      engine_factory=_Factory())


def test_checkpoint_keeps_adam_slots_and_writes_atomically(tmp_path):
    cfg = E.EngineConfig(workflow="lrcn", fpc=2, height=67, width=67, optimizer="adam")
    eng = FakeEngine(cfg, 4)
    rng = np.random.default_rng(1)
    for name, shape in eng.var_shapes:
        eng.adam[name + "/Adam"] = rng.standard_normal(shape).astype(np.float32)
        eng.adam[name + "/Adam_1"] = rng.random(shape).astype(np.float32)
    eng.adam_t, eng.global_step = 7, 7
    prefix = checkpoint.save(eng, str(tmp_path), "ep_1_btch_7_gs_7", 7, 0, max_to_keep=2)
    blob = np.load(prefix + ".npz")
    keys = {k.replace("|", "/") for k in blob.files}
    assert "beta1_power" in keys and "dcnn/conv1W/Adam" in keys and "dcnn/conv1W/Adam_1" in keys
    assert abs(float(blob["beta1_power"]) - 0.9 ** 8) < 1e-7
    assert not [f for f in os.listdir(tmp_path / "checkpoints") if ".tmp" in f]  # temporaries were renamed away
    fresh = FakeEngine(cfg, 4)
    snap = checkpoint.restore(fresh, prefix)
    assert snap == [7, 0, 7] and fresh.adam_t == 7 and fresh.global_step == 7
    for name, _ in eng.var_shapes:
        assert np.array_equal(fresh.adam[name + "/Adam"], eng.adam[name + "/Adam"])
        assert np.array_equal(fresh.adam[name + "/Adam_1"], eng.adam[name + "/Adam_1"])
    # validation ignores global_step and the optimiser slots (feeder.py:226-227)
    val_eng = FakeEngine(cfg, 4)
    checkpoint.restore(val_eng, prefix, is_validation=True)
    assert val_eng.global_step == 0 and not val_eng.adam
    # max_to_keep prunes the files only after the index stopped naming them
    p2 = checkpoint.save(eng, str(tmp_path), "a", 1, 0, max_to_keep=2)
    p3 = checkpoint.save(eng, str(tmp_path), "b", 2, 0, max_to_keep=2)
    names = checkpoint.read_index(str(tmp_path / "checkpoints"))
    assert names == [p2, p3] and not os.path.exists(prefix + ".npz")
    assert checkpoint.resolve(str(tmp_path), "latest") == p3


def test_image_shape_reaches_the_engine_config(tmp_path):
    st = Settings()
    fd = st.initialize(_cfg("config2_lrcn_train.yml", tmp_path, _small_train(epochs=1)))
    cfg = st.engine_config(fd.main.fpc, fd.main.opts.image_shape)
    assert (cfg.height, cfg.width) == (67, 67)
    with pytest.raises(Exception, match="too small"):
        st.engine_config(fd.main.fpc, (32, 32, 3))
    with pytest.raises(Exception, match=r"\[h, w, 3\]"):
        st.engine_config(fd.main.fpc, (227, 227, 1))


def test_data_parallel_shards_whole_videos_and_covers_the_batch():
    fpc = 3
    cpvs = [2, 1, 3, 1, 2]
    clips = sum(cpvs)
    frames = np.arange(clips * fpc)
    onehot = np.arange(clips)
    crops = np.arange(clips * fpc * 3).reshape(-1, 3)
    for world in (1, 2, 3, 4, 7):
        got_f, got_l = [], []
        for rank in range(world):
            dp = run_task.DataParallel(rank, world)
            f, l, c = dp.shard_batch(frames, onehot, cpvs, crops, fpc)
            assert len(f) == len(l) * fpc and len(c) == len(f)
            # shard boundaries fall on video boundaries
            starts = set(np.cumsum([0] + cpvs).tolist())
            if len(l):
                assert int(l[0]) in starts and int(l[-1]) + 1 in starts
            got_f.append(f)
            got_l.append(l)
        assert np.array_equal(np.concatenate(got_f), frames) and np.array_equal(np.concatenate(got_l), onehot)


def test_validation_workflow_with_the_stand_in_engine(tmp_path):
    def small_val(run):
        d = run["data"]["synthetic-val"]
        d["num_items"] = 5
        d["clips_per_video"] = [2, 1, 3, 1, 2]
        d["num_frames_per_clip"] = 2
        d["image_shape"] = "(67, 67, 3)"
        run["val"]["batch_size"] = 2
    fac = _Factory()
    acc = run_task.main(_cfg("config5_lrcn_val.yml", tmp_path, small_val), engine_factory=fac)
    eng = fac.engines[0]
    assert [c[1] for c in eng.calls] == [3, 4, 2]  # clips per batch: whole videos, 2 videos per batch
    assert 0.0 <= acc <= 1.0
