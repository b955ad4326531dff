"""Configuration language, pipeline validation rules, feeder contract, checkpoints (CPU)."""
import os
import pickle
import types

import numpy as np
import pytest
import yaml

import vlb200  # noqa: F401
from vlb200 import checkpoint, feeder
from vlb200.defs import defs
from vlb200.settings import Settings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name, tmp_path, mutate=None):
    with open(os.path.join(ROOT, "configs", name)) as f:
        cfg = yaml.safe_load(f)
    cfg["run"]["run_folder"] = str(tmp_path / "run")
    if mutate:
        mutate(cfg["run"])
    p = tmp_path / name
    with open(p, "w") as f:
        yaml.safe_dump(cfg, f)
    return str(p)


def test_defs_check():
    assert defs.check("defs.optim.sgd", defs.optim) == "sgd"
    assert defs.check("defs.fusion_method.avg", (defs.fusion_type, defs.fusion_method)) == "avg"
    assert defs.check("defs.dataset_tag.main", defs.dataset_tag, do_boolean=True) == (True, "main")
    assert defs.check("frames", defs.dataset_tag, do_boolean=True) == (False, None)
    for bad in ("optim.sgd", "defs.optim.nope", "defs.nofamily.x"):
        with pytest.raises(Exception):
            defs.check(bad, defs.optim)
    with pytest.raises(Exception):
        defs.check("defs.decay.exp", defs.optim)  # valid def, wrong family


def test_config2_maps_to_lrcn_engine_config(tmp_path):
    st = Settings()
    fd = st.initialize(_load("config2_lrcn_train.yml", tmp_path))
    cfg = st.engine_config(fd.main.fpc)
    assert (cfg.workflow, cfg.frame_encoding_layer, cfg.lstm_hidden, cfg.lstm_layers, cfg.fusion) == \
        ("lrcn", "fc7", 256, 1, "avg")
    assert (cfg.optimizer, cfg.clip_norm, cfg.dropout_keep_prob, cfg.num_classes, cfg.fpc) == ("sgd", 10, 0.5, 101, 16)
    assert st.train.lr_decay == ["exp", "interval", 1000, 0.96] and st.train.lr_mult is None
    assert st.run_id == "config2_train_scratch"
    assert fd.get_num_batches() == 4 and fd.max_clips_per_batch() == 64


def test_config1_singleframe_and_val_dropout(tmp_path):
    st = Settings()
    fd = st.initialize(_load("config1_singleframe_train.yml", tmp_path))
    cfg = st.engine_config(fd.main.fpc)
    assert cfg.workflow == "singleframe" and cfg.fusion == "avg" and cfg.frame_encoding_layer == "fc8"
    st5 = Settings()
    fd5 = st5.initialize(_load("config5_lrcn_val.yml", tmp_path))
    assert st5.get_dropout() == 0.0  # validation: dropout skipped (settings_.py:107-110, lstm.py:52)
    assert st5.val.clip_fusion_method == "avg" and fd5.max_clips_per_batch() == 64


def test_pipeline_validation_errors(tmp_path):
    def unknown_field(run):
        run["network"]["pipelines"][0]["lrcn"]["load_weights"] = "x.npy"  # config.example.yml:43 is rejected too
    with pytest.raises(Exception, match="Undefined pipeline field"):
        Settings().initialize(_load("config2_lrcn_train.yml", tmp_path, unknown_field))

    def bad_input(run):
        run["network"]["pipelines"][0]["lrcn"]["input"] = "frames"
    with pytest.raises(Exception, match="not a dataset tag"):
        Settings().initialize(_load("config2_lrcn_train.yml", tmp_path, bad_input))

    def lr_mult(run):
        run["train"]["lr_mult"] = 0.1
    st = Settings()
    fd = st.initialize(_load("config2_lrcn_train.yml", tmp_path, lr_mult))
    from vlb200.train import Train
    with pytest.raises(Exception, match="lr_mult"):
        Train(st, fd, engine=None)

    def rmsprop(run):
        run["train"]["optimizer"] = "defs.optim.rmsprop"
    st = Settings()
    fd = st.initialize(_load("config2_lrcn_train.yml", tmp_path, rmsprop))
    with pytest.raises(Exception, match="Undefined optimizer"):
        Train(st, fd, engine=None)
    with pytest.raises(Exception, match="ini files deprecated"):
        Settings().initialize("config.ini")

    def fusion_with_lstm(run):
        run["network"]["pipelines"][0]["lrcn"]["frame_fusion"] = ["defs.fusion_type.early", "defs.fusion_method.avg"]
    st = Settings()
    fd = st.initialize(_load("config2_lrcn_train.yml", tmp_path, fusion_with_lstm))
    with pytest.raises(Exception, match=r"only with \[none\] fusion"):  # the reference's own text (model.py:125)
        st.engine_config(16)
    st = Settings()
    fd = st.initialize(_load("config2_lrcn_train.yml", tmp_path))
    with pytest.raises(Exception, match="requires an fpc greater than 1"):  # model.py:121
        st.engine_config(1)


def test_feeder_contract(tmp_path):
    """Frames ordered video -> clip -> frame, one label row per clip, whole videos per batch, padding 0."""
    def ragged(run):
        d = run["data"]["synthetic-val"]
        d["clips_per_video"] = [3, 1, 2, 4, 1]
        d["num_items"] = 5
        d["num_frames_per_clip"] = 2
        d["image_shape"] = "(8, 8, 3)"
        run["val"]["batch_size"] = 2
    st = Settings()
    fd = st.initialize(_load("config5_lrcn_val.yml", tmp_path, ragged))
    seen = []
    while fd.loop():
        frames, onehot, cpvs, nd, nl, pad = fd.get_feed_dict()
        assert pad == 0 and frames.dtype == np.uint8 and onehot.dtype == np.int32
        assert nd == sum(cpvs) * 2 and nl == sum(cpvs) and frames.shape == (nd, 8, 8, 3)
        off = 0
        for c in cpvs:  # every clip of a video carries the video's label
            assert (onehot[off:off + c] == onehot[off]).all() and onehot[off].sum() == 1
            off += c
        seen.append(cpvs)
    assert seen == [[3, 1], [2, 4], [1]]
    assert fd.max_clips_per_batch() == 6


class _FakeEngine(object):
    def __init__(self):
        self.var_shapes = [("dcnn/conv1W", (2, 2)), ("output_fc_b", (3,))]
        self.sd = {"dcnn/conv1W": np.arange(4, dtype=np.float32).reshape(2, 2), "output_fc_b": np.ones(3, np.float32)}
        self.global_step = 17

    def state_dict(self):
        return dict(self.sd, global_step=np.int32(self.global_step))

    def load_state_dict(self, sd):
        self.loaded = sd
        if "global_step" in sd:
            self.global_step = int(sd["global_step"])

    cfg = types.SimpleNamespace(optimizer="sgd")

    def optimizer_state_dict(self):
        return {}

    def load_optimizer_state_dict(self, sd):
        return 0


def test_checkpoint_roundtrip_and_snap_format(tmp_path):
    eng = _FakeEngine()
    prefix = checkpoint.save(eng, str(tmp_path), "ep_1_btch_4_gs_17", 4, 0, max_to_keep=2)
    assert os.path.basename(prefix).endswith("_ep_1_btch_4_gs_17.graph-17")
    with open(prefix + ".snap", "rb") as f:
        assert pickle.load(f) == [4, 0, 17]      # [batch_index, epoch_index, global_step] (feeder.py:283-286)
    assert checkpoint.resolve(str(tmp_path), "latest") == prefix
    eng2 = _FakeEngine()
    eng2.global_step = 0
    assert checkpoint.restore(eng2, prefix) == [4, 0, 17]
    assert np.array_equal(eng2.loaded["dcnn/conv1W"], eng.sd["dcnn/conv1W"]) and eng2.global_step == 17
    eng3 = _FakeEngine()
    eng3.global_step = 0
    checkpoint.restore(eng3, prefix, is_validation=True)
    assert eng3.global_step == 0                  # global_step ignorable in validation (feeder.py:226-227)
    eng3.var_shapes.append(("dcnn/fc6W", (1,)))
    with pytest.raises(Exception, match="missing"):
        checkpoint.restore(eng3, prefix)
    # retention: max_to_keep
    for i in range(3):
        eng.global_step = 18 + i
        checkpoint.save(eng, str(tmp_path), "ep_1_btch_%d_gs_%d" % (5 + i, 18 + i), 5 + i, 0, max_to_keep=2)
    kept = checkpoint.read_index(os.path.join(str(tmp_path), "checkpoints"))
    assert len(kept) == 2 and all(os.path.exists(k + ".npz") for k in kept)
    with open(os.path.join(str(tmp_path), "checkpoints", "checkpoint")) as f:
        first = f.readline().strip()  # TensorFlow's layout: the reference's resume_snap reads this line (feeder.py:148-156)
    assert first == 'model_checkpoint_path: "%s"' % kept[-1]


def test_fusion_variants_map_to_engine_configs(tmp_path):
    """Model.build_pipeline's variants beyond the BASELINE workflows (model.py:103-117,137-151; lstm.py:81-93) ->
    EngineConfig; `reshape` (tf_util.py:24-27) is rejected with the reference's text."""
    from vlb200 import engine as E
    from vlb200.settings import pipeline_engine_config

    def net(**kw):
        base = dict(input=["main"], input_fusion=None, representation="dcnn", frame_encoding_layer="fc7", classifier="fc",
                    lstm_params=None, frame_fusion=None, weights_file=None)
        base.update(kw)
        return types.SimpleNamespace(**base)
    cfg = pipeline_engine_config(net(classifier="lstm", lstm_params=[64, 2, "state"]), 11, 4, "sgd", 10, 0.5)
    assert (cfg.workflow, cfg.fusion) == ("lrcn", "state")
    names = [n for n, _ in E.variable_shapes(cfg)]
    assert "fc_convert_w" in names and "output_fc_w" not in names  # convert_dim_fc's default name (model.py:142-143)
    cfg = pipeline_engine_config(net(frame_fusion=["early", "avg"]), 11, 4, "sgd", 10, 0.0)
    assert (cfg.workflow, cfg.early_fusion, cfg.fusion, cfg.frame_encoding_layer) == ("fc", True, "avg", "fc7")
    assert dict(E.variable_shapes(cfg))["fc_convert_w"] == (4096, 11)
    cfg = pipeline_engine_config(net(frame_encoding_layer="fc6", frame_fusion=["late", "last"]), 11, 4, "sgd", 10, 0.0)
    assert (cfg.workflow, cfg.early_fusion, cfg.fusion, cfg.frame_encoding_layer) == ("fc", False, "last", "fc6")
    assert "dcnn/fc7W" not in dict(E.variable_shapes(cfg))
    cfg = pipeline_engine_config(net(frame_encoding_layer="fc8", frame_fusion=["early", "avg"]), 11, 4, "sgd", 10, 0.0)
    assert cfg.workflow == "singleframe"  # pooled fc8 logits: early and late fusion coincide (no op in between)
    with pytest.raises(Exception, match="Undefined frame fusion type : reshape"):
        pipeline_engine_config(net(classifier="lstm", lstm_params=[64, 1, "reshape"]), 11, 4, "sgd", 10, 0.0)
    with pytest.raises(Exception, match="needs frame_fusion"):
        pipeline_engine_config(net(), 11, 4, "sgd", 10, 0.0)
    with pytest.raises(Exception, match="late fusion with no classifier"):
        pipeline_engine_config(net(classifier=None, frame_fusion=["late", "avg"]), 11, 4, "sgd", 10, 0.0)
    with pytest.raises(Exception, match="Multi-input"):
        pipeline_engine_config(net(input=["main", "aux"], input_fusion="concat"), 11, 4, "sgd", 10, 0.0)
