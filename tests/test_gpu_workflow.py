"""GPU: the reference-facing workflow (run_task loops, validation fusion on the device) end to end."""
import os
import types

import numpy as np
import pytest
import yaml

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VAL = np.load(os.path.join(ROOT, "tests", "golden", "reference_val_golden.npz"))


@pytest.mark.parametrize("method", ["avg", "last"])
def test_device_clip_fusion_bit_exact_vs_reference_golden(method, tmp_path):
    """Pooled video logits and labels computed by the CUDA segmented reduction == the reference's own
    Validation.apply_clip_fusion output (golden vectors), bit for bit."""
    import vlb200  # noqa: F401
    from vlb200.val import Validation
    st = types.SimpleNamespace(num_classes=101, run_id="t", run_folder=str(tmp_path), val=None)
    v = Validation(st, use_device=True)
    v.process_validation_logits(VAL["logits"], VAL["labels"], [int(c) for c in VAL["cpv"]], method)
    assert np.array_equal(v.item_logits, VAL["video_logits_" + method])
    assert np.array_equal(v.item_logits.argmax(1), VAL["video_logits_" + method].argmax(1))
    assert v.get_accuracy() == float(VAL["get_accuracy_" + method])


def _cfg(name, tmp_path, mutate):
    with open(os.path.join(ROOT, "configs", name)) as f:
        cfg = yaml.safe_load(f)
    cfg["run"]["run_folder"] = str(tmp_path / "run")
    mutate(cfg["run"])
    p = tmp_path / name
    with open(p, "w") as f:
        yaml.safe_dump(cfg, f)
    return str(p)


def test_run_task_train_then_resume_then_validate(tmp_path):
    import vlb200  # noqa: F401
    from vlb200 import run_task

    def small_train(run):
        d = run["data"]["synthetic-train"]
        d["num_items"] = 6
        d["num_frames_per_clip"] = 2
        run["train"]["batch_size"] = 2
        run["train"]["epochs"] = 1
    run_task.main(_cfg("config2_lrcn_train.yml", tmp_path, small_train))
    folder = tmp_path / "run" / "checkpoints"
    from vlb200 import checkpoint
    names = checkpoint.read_index(str(folder))
    assert len(names) == 1 and names[0].endswith("ep_1_btch_3_gs_3.graph-3")
    assert os.path.exists(tmp_path / "run" / "config2_train_scratch_lr_decay_schedule.txt")

    def small_val(run):
        d = run["data"]["synthetic-val"]
        d["num_items"] = 5
        d["clips_per_video"] = [2, 1, 3, 1, 2]
        d["num_frames_per_clip"] = 2
        run["val"]["batch_size"] = 2
        run["resume_file"] = "latest"
        run["run_id"] = "config2"
    acc = run_task.main(_cfg("config5_lrcn_val.yml", tmp_path, small_val))
    assert 0.0 <= acc <= 1.0
    files = os.listdir(tmp_path / "run")
    assert any(f.startswith("validation_logits_config2_val_resume") and f.endswith(".total") for f in files)
    assert "accuracy_config2_val_resume" in files
    import pickle
    tot = [f for f in files if f.endswith(".total")][0]
    with open(tmp_path / "run" / tot, "rb") as f:
        logits = pickle.load(f)
    assert logits.shape == (5, 101) and logits.dtype == np.float32  # one fused row per video (val.py:124-137 format)


def test_run_task_on_a_serialized_tfrecord_dataset(tmp_path):
    """train on `<path>.tfrecord` + `<path>.size` as written by the reference's serialize.py, read without
    TensorFlow, cropped / mirrored / mean-subtracted on the device; and the engine's forward on those batches equals
    its forward on the numpy-preprocessed frames."""
    import random
    import torch
    import vlb200  # noqa: F401
    from vlb200 import run_task, tfrecord
    from vlb200.defs import defs
    from vlb200.engine import Engine, EngineConfig
    from vlb200.feeder import Dataset

    rng = np.random.default_rng(5)
    cpv, fpc, raw_shape = [1, 2, 1], 2, (240, 250, 3)
    base = str(tmp_path / "ucf_like")
    with open(base + ".tfrecord", "wb") as f:
        for v, c in enumerate(cpv):
            for _ in range(c * fpc):
                tfrecord.write_record(f, tfrecord.serialize_frame(rng.integers(0, 256, size=raw_shape, dtype=np.uint8), [v]))
    tfrecord.write_size_file(base + ".size", cpv, fpc)

    def mutate(run):
        run["data"] = {"ucf": {"data_format": "defs.data_format.tfrecord", "data_path": base,
                               "image_shape": "(227, 227, 3)", "raw_image_shape": "(240, 250, 3)",
                               "imgproc": ["defs.imgproc.rand_crop", "defs.imgproc.rand_mirror", "defs.imgproc.sub_mean"],
                               "mean_image": [99.197148, 105.293620, 109.503945], "num_frames_per_clip": fpc,
                               "phase": "defs.phase.train", "tag": "defs.dataset_tag.main"}}
        run["train"]["batch_size"] = 2
        run["train"]["epochs"] = 1
    run_task.main(_cfg("config2_lrcn_train.yml", tmp_path, mutate))
    from vlb200 import checkpoint
    names = checkpoint.read_index(str(tmp_path / "run" / "checkpoints"))
    assert names and names[-1].endswith("ep_1_btch_2_gs_2.graph-2")

    # device preprocessing == numpy preprocessing on the same draws
    opts = types.SimpleNamespace(name="t", data_format=defs.data_format.tfrecord, data_path=base, image_shape=(227, 227, 3),
                                 num_frames_per_clip=fpc, imgproc=[defs.imgproc.rand_crop, defs.imgproc.rand_mirror],
                                 raw_image_shape=raw_shape, verify_records="full", mean_image=None)
    ds = Dataset(opts, batch_size=3, num_classes=101, epochs=1, save_freq_per_epoch=1)
    random.seed(9)
    frames, onehot, _ = ds.next_batch()
    mean = (99.197148, 105.293620, 109.503945)
    cfg = EngineConfig(workflow="lrcn", fusion="avg", fpc=fpc, num_classes=101, lstm_hidden=256, mean=mean)
    eng = Engine(cfg, max_clips=4)
    got = eng.forward(frames, ds.last_crops)
    pre = np.stack([(frames[i, y:y + 227, x:x + 227][:, ::-1] if m else frames[i, y:y + 227, x:x + 227]).astype(np.float32)
                    - np.array(mean, np.float32) for i, (y, x, m) in enumerate(ds.last_crops)])
    ref = eng.forward(pre)  # fp32 feed of already-preprocessed frames (feeder.py:97-100 contract)
    assert np.array_equal(got, ref)


def test_raw_resize_on_the_device_equals_pil_then_crop(tmp_path):
    """imgproc raw_resize + center_crop (dataset_.py:481-491): frames serialized at 120x160 are resampled to
    raw_image_shape on the device with Pillow's bilinear (bit-exact: the oracle restatement is pinned on PIL's own
    outputs) and then cropped; the engine's logits equal its logits on frames preprocessed on the host."""
    import torch
    import vlb200  # noqa: F401
    from oracle import resize_pil as R
    from vlb200 import run_task, tfrecord
    from vlb200.defs import defs
    from vlb200.engine import Engine, EngineConfig
    from vlb200.feeder import Dataset

    rng = np.random.default_rng(6)
    cpv, fpc, stored, raw_shape = [1, 1, 2], 2, (120, 160, 3), (240, 250, 3)
    base = str(tmp_path / "small_frames")
    with open(base + ".tfrecord", "wb") as f:
        for v, c in enumerate(cpv):
            for _ in range(c * fpc):
                tfrecord.write_record(f, tfrecord.serialize_frame(rng.integers(0, 256, size=stored, dtype=np.uint8), [v]))
    tfrecord.write_size_file(base + ".size", cpv, fpc)

    opts = types.SimpleNamespace(name="t", data_format=defs.data_format.tfrecord, data_path=base, image_shape=(227, 227, 3),
                                 num_frames_per_clip=fpc, imgproc=[defs.imgproc.raw_resize, defs.imgproc.center_crop],
                                 raw_image_shape=raw_shape, verify_records="full", mean_image=None)
    ds = Dataset(opts, batch_size=4, num_classes=101, epochs=1, save_freq_per_epoch=1)
    assert ds.resize_to == (240, 250)
    frames, onehot, _ = ds.next_batch()
    assert frames.shape == (8, 120, 160, 3) and frames.dtype == np.uint8  # still at the serialized size on the host
    mean = (99.197148, 105.293620, 109.503945)
    cfg = EngineConfig(workflow="lrcn", fusion="avg", fpc=fpc, num_classes=101, lstm_hidden=256, mean=mean)
    eng = Engine(cfg, max_clips=4)
    eng.set_read_resize(ds.resize_to)
    got = eng.forward(frames, ds.last_crops)
    resized = R.imresize_bilinear(frames, 240, 250)
    y0, x0 = ds.last_crops[0, 0], ds.last_crops[0, 1]
    assert (y0, x0) == ((240 - 227) // 2, (250 - 227) // 2)
    pre = resized[:, y0:y0 + 227, x0:x0 + 227].astype(np.float32) - np.array(mean, np.float32)
    eng.set_read_resize(None)
    ref = eng.forward(pre)
    assert np.array_equal(got, ref)

    # and the whole workflow runs with imgproc resize (no crop): frames resampled straight to the network input
    def mutate(run):
        run["data"] = {"ucf": {"data_format": "defs.data_format.tfrecord", "data_path": base,
                               "image_shape": "(227, 227, 3)",
                               "imgproc": ["defs.imgproc.resize", "defs.imgproc.sub_mean"],
                               "mean_image": [99.197148, 105.293620, 109.503945], "num_frames_per_clip": fpc,
                               "phase": "defs.phase.train", "tag": "defs.dataset_tag.main"}}
        run["train"]["batch_size"] = 2
        run["train"]["epochs"] = 1
    run_task.main(_cfg("config2_lrcn_train.yml", tmp_path, mutate))
    from vlb200 import checkpoint
    names = checkpoint.read_index(str(tmp_path / "run" / "checkpoints"))
    assert names and names[-1].endswith("gs_2.graph-2")


def test_tfshim_session_drives_the_real_engine(tmp_path):
    """The reference-facing seam (tfshim.Session / compat.Model / compat.Train / tf.train.Saver) over the CUDA engine:
    `sess.run([summaries, loss, lr, global_step, optimizer], feed_dict)` and `sess.run(model.logits, feed_dict)` give
    what direct Engine calls give on the same inputs, and a Saver round trip restores the variables bit for bit.
    (tests/test_reference_dropin.py runs the reference's own run_task.main through the same objects.)"""
    import vlb200  # noqa: F401
    from vlb200 import compat, tfshim
    from vlb200 import engine as E
    fpc, clips, classes = 2, 3, 11
    ds = types.SimpleNamespace(num_frames_per_clip=fpc, clips_per_video=[1] * 6, num_items=6,
                               get_image_shape=lambda: (227, 227, 3))
    pipe = types.SimpleNamespace(input=["main"], input_shape=[None], input_fusion=None, representation="dcnn",
                                 frame_encoding_layer="fc7", classifier="lstm", lstm_params=[64, 1, "avg"],
                                 frame_fusion=None, weights_file=None, fc_output_dim=None)
    tr = types.SimpleNamespace(optimizer="sgd", clip_norm=10, lr_mult=None, base_lr=0.01,
                               lr_decay=["exp", "interval", 1, 0.5], epochs=2, dropout_keep_prob=0.0)
    settings = types.SimpleNamespace(
        pipeline_names=["lrcn"], pipelines={"lrcn": pipe}, num_classes=classes, train=tr, val=None, global_step=0,
        run_folder=str(tmp_path), run_id="shim", get_dropout=lambda: 0.0, get_batch_size=lambda: clips,
        feeder=types.SimpleNamespace(get_dataset_by_tag=lambda tag: [ds], get_num_batches=lambda: 2))
    model = compat.Model(settings)
    summaries = types.SimpleNamespace(train=[], val=[])
    train = compat.Train(settings, settings.feeder, model.get_output(), summaries)
    sess = tfshim.Session()
    sess.run(tfshim.global_variables_initializer())
    rng = np.random.default_rng(3)
    frames = [rng.uniform(-1, 1, size=(227, 227, 3)).astype(np.float32) for _ in range(clips * fpc)]  # feeder.py:97-100
    onehot = np.zeros((clips, classes), np.int32)
    onehot[np.arange(clips), [1, 4, 7]] = 1
    fdict = {model.required_input[0][0]: frames, train.required_input[0][0]: list(onehot)}
    logits0 = sess.run(model.logits, feed_dict=fdict)
    # the same engine configuration built directly
    ref = E.Engine(model.cfg, max_clips=clips, params=E.init_variables(model.cfg))
    assert np.array_equal(logits0, ref.forward(np.stack(frames)))
    merged = tfshim.summary.merge(summaries.train)
    for step in range(2):
        _, loss, lr, gstep, _ = sess.run([merged, train.loss, train.current_lr, train.global_step, train.optimizer],
                                         feed_dict=fdict)
        rloss, rlr, rstep, _, _ = ref.train_step(np.stack(frames), onehot, [0.01, 0.005][step])
        assert gstep == rstep == step + 1 and abs(lr - rlr) < 1e-9 and abs(loss - rloss) < 1e-6 * max(1.0, abs(rloss))
    assert sess.run(train.global_step) == 2
    saver = tfshim.train.Saver(max_to_keep=3)
    prefix = saver.save(sess, str(tmp_path / "checkpoints" / "x.graph"), global_step=2)
    assert prefix.endswith("x.graph-2") and all(os.path.exists(prefix + e) for e in (".npz", ".meta", ".index"))
    names = tfshim._checkpoint_tensor_names(prefix)
    assert {v.name[:-2] for v in tfshim.global_variables()} <= set(names)  # the name diff of feeder.py:229-249 is empty
    before = sess.engine.state_dict()
    sess.run([train.loss, train.optimizer], feed_dict=dict(fdict))  # a third step moves the variables
    assert not np.array_equal(before["output_fc_w"], sess.engine.state_dict()["output_fc_w"])
    saver.restore(sess, prefix)
    after = sess.engine.state_dict()
    assert all(np.array_equal(before[k], after[k]) for k in before)


def test_validation_workflow_in_fp32_accuracy_mode(tmp_path, monkeypatch):
    """VLB200_FP32=1 routes the validation workflow (run_task.do_test -> Engine.forward) through the fp32-accuracy mode:
    the saved video logits move by bf16-level differences only and the accuracy file is written as usual."""
    import pickle
    import vlb200  # noqa: F401
    from vlb200 import run_task

    def small_val(run):
        d = run["data"]["synthetic-val"]
        d["num_items"] = 4
        d["clips_per_video"] = [2, 1, 1, 2]
        d["num_frames_per_clip"] = 2
        run["val"]["batch_size"] = 2
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("VLB200_FP32", mode)
        sub = tmp_path / ("m" + mode)
        sub.mkdir()
        acc = run_task.main(_cfg("config5_lrcn_val.yml", sub, small_val))
        files = os.listdir(sub / "run")
        tot = [f for f in files if f.endswith(".total")][0]
        with open(sub / "run" / tot, "rb") as f:
            out[mode] = (acc, pickle.load(f))
    a, b = out["0"][1], out["1"][1]
    assert a.shape == b.shape == (4, 101) and b.dtype == np.float32
    rel = np.abs(a.astype(np.float64) - b).max() / np.abs(b).max()
    # the two modes differ (the route is taken); on the synthetic uint8 frames the LSTM gates are saturated and bf16 vs
    # fp32 evaluations differ by 5-8e-2 (the noise floor of that regime), so the bound is that of the regime
    assert 0 < rel < 1.5e-1
