"""Run configuration: the YAML `run:` section with the reference's keys and `defs.*` values.

Reference: settings_.py:167-208 (read_network), :210-366 (read_config), :373-444 (initialize).  The configuration
language, validation rules and error texts are kept; what changes is what gets built from it: an EngineConfig for
the device path instead of a TensorFlow graph.
"""
import ast
import logging
import os
import time

import yaml

from .defs import defs
from .engine import EngineConfig
from .utils import configure_logging, debug, error, info, warning


def parse_seq(arg):
    """Tuples/lists may be written as strings in the YAML (parse_opts.py:6-12)."""
    if isinstance(arg, (list, tuple)):
        return arg
    try:
        return ast.literal_eval(arg)
    except Exception:
        error("Unable to literal-eval expression [%s]" % arg)


class Network(object):
    """One entry of network.pipelines (settings_.py:171-197)."""
    FIELDS = ("input", "representation", "frame_encoding_layer", "fc_output_dim", "classifier", "lstm_params",
              "weights_file", "frame_fusion", "input_shape", "input_fusion")

    def __init__(self):
        self.input = None
        self.representation = None
        self.frame_encoding_layer = None
        self.fc_output_dim = None
        self.classifier = None
        self.lstm_params = None
        self.weights_file = None
        self.frame_fusion = None
        self.input_shape = None
        self.input_fusion = None


class TrainOpts(object):
    batch_size = epochs = optimizer = base_lr = lr_mult = lr_decay = clip_norm = dropout_keep_prob = None
    epoch_index = 0


class ValOpts(object):
    batch_size = logits_save_interval = clip_fusion_type = clip_fusion_method = None


class DataOpts(object):
    pass


class Settings(object):
    def __init__(self):
        self.pipelines = {}
        self.pipeline_names = []
        self.train = None
        self.val = None
        self.run_id = None
        self.global_step = 0
        self.data = []
        self.resume_file = None

    # ------------------------------------------------------------------------------------------------
    def should_resume(self):
        return self.resume_file is not None and self.resume_file != "None"

    def get_dropout(self):
        """keep_prob in training, 0.0 (dropout skipped, lstm.py:52) in validation (settings_.py:107-110)."""
        if self.phase == defs.phase.train and self.train:
            return self.train.dropout_keep_prob
        return 0.0

    def read_network(self, content):
        net = Network()
        unknown = [k for k in content if k not in Network.FIELDS]
        if unknown:
            error("Undefined pipeline field(s):" + str(unknown))
        inputs = content.get("input")
        inputs = list(inputs) if isinstance(inputs, (list, tuple)) else [inputs]
        if any(x is None for x in inputs):
            error("<None> or undefined <input> tag in pipeline: %s" % content)
        for i, inp in enumerate(inputs):
            if inp in self.pipelines:
                continue
            ok, tag = defs.check(inp, defs.dataset_tag, do_boolean=True)
            if not ok:
                error("Input identifier [%s] is not a dataset tag, but no such pipeline has been declared yet." % inp)
            inputs[i] = tag
        net.input = inputs
        if content.get("representation") is None:
            error("Undefined required field [representation]")
        net.representation = defs.check(content["representation"], defs.representation)
        if net.representation == defs.representation.dcnn:
            if content.get("frame_encoding_layer") is None:
                error("Undefined required field [frame_encoding_layer]")
            net.frame_encoding_layer = content["frame_encoding_layer"]
        if net.representation == defs.representation.fc:
            if content.get("fc_output_dim") is None:
                error("Undefined required field [fc_output_dim]")
            net.fc_output_dim = content["fc_output_dim"]
        if content.get("classifier") is not None:
            net.classifier = defs.check(content["classifier"], defs.classifier)
        if net.classifier == defs.classifier.lstm:
            params = parse_seq(content.get("lstm_params"))
            net.lstm_params = [int(params[0]), int(params[1]), defs.check(params[2], defs.fusion_method)]
        net.weights_file = content.get("weights_file")
        ff = content.get("frame_fusion")
        if ff is not None:
            ff = parse_seq(ff)
            net.frame_fusion = [defs.check(ff[0], defs.fusion_type), defs.check(ff[1], defs.fusion_method)]
        shp = content.get("input_shape")
        net.input_shape = None if shp in (None, "None") else tuple(parse_seq(shp))
        if content.get("input_fusion") is not None:
            net.input_fusion = defs.check(content["input_fusion"], defs.fusion_method)
        return net

    def read_config(self, config, init_file):
        self.resume_file = config.get("resume_file")
        self.run_folder = config["run_folder"]
        self.run_id = config.get("run_id")
        phases = config["phase"]
        phases = phases if isinstance(phases, list) else [phases]
        self.phases = [defs.check(p, defs.phase) for p in phases]
        self.phase = self.phases[0]
        tag = ("train" if defs.phase.train in self.phases else "") + ("val" if defs.phase.val in self.phases else "")
        tag += "_resume" if self.should_resume() else "_scratch"
        self.run_id = "_".join([self.run_id if self.run_id else os.path.basename(init_file), tag])
        if not os.path.exists(self.run_folder):
            warning("Non existent run folder %s - creating." % self.run_folder)
            os.makedirs(self.run_folder, exist_ok=True)  # (several data-parallel ranks may get here at once)

        log = config.get("logging", {}) or {}
        self.save_freq_per_epoch = log.get("save_freq_per_epoch", 1)
        self.logging_level = log.get("level", "logging.INFO")
        if self.logging_level not in ["logging." + x for x in ("INFO", "DEBUG", "WARN")]:
            error("Invalid logging level: %s" % self.logging_level)
        self.tensorboard_folder = os.path.join(self.run_folder, log.get("tensorboard_folder", "tensorboard"))
        logfile = os.path.join(self.run_folder, "log_%s_%s.log" % (self.run_id, time.strftime("%d%m%y_%H%M%S")))
        configure_logging(logfile, getattr(logging, self.logging_level.split(".")[1]))

        for pipeline in config["network"]["pipelines"]:
            pname, content = list(pipeline.items())[0]
            debug("Reading network [%s]" % pname)
            self.pipelines[pname] = self.read_network(content)
            self.pipeline_names.append(pname)
        self.num_classes = int(config["network"]["num_classes"])

        for phase in self.phases:
            obj = config[phase]
            if phase == defs.phase.train:
                t = TrainOpts()
                t.batch_size = int(obj["batch_size"])
                t.epochs = int(obj["epochs"])
                t.optimizer = defs.check(obj["optimizer"], defs.optim)
                t.base_lr = float(obj["base_lr"])
                t.lr_mult = float(obj["lr_mult"]) if obj.get("lr_mult", "None") not in ("None", None) else None
                if obj.get("lr_decay") in (None, "None"):
                    t.lr_decay = None
                else:
                    d = parse_seq(obj["lr_decay"])
                    t.lr_decay = [defs.check(d[0], defs.decay), defs.check(d[1], defs.periodicity), int(d[2]),
                                  float(d[3])]
                t.clip_norm = int(obj["clip_norm"])  # read as int like settings_.py:288 (10.5 -> 10)
                t.dropout_keep_prob = float(obj["dropout_keep_prob"])
                self.train = t
            else:
                v = ValOpts()
                v.batch_size = int(obj["batch_size"])
                v.logits_save_interval = int(obj["logits_save_interval"])
                cf = parse_seq(obj["clip_fusion"])
                v.clip_fusion_type = defs.check(cf[0], defs.fusion_type)
                v.clip_fusion_method = defs.check(cf[1], defs.fusion_method)
                self.val = v

        for name, d in (config.get("data") or {}).items():
            o = DataOpts()
            o.name = name
            o.phase = defs.check(d["phase"], defs.phase)
            o.tag = defs.check(d.get("tag", "defs.dataset_tag.main"), defs.dataset_tag)
            o.data_format = defs.check(d.get("data_format", "defs.data_format.synthetic"), defs.data_format)
            o.data_path = d.get("data_path")
            o.image_shape = tuple(parse_seq(d.get("image_shape", (227, 227, 3))))
            o.mean_image = d.get("mean_image")
            o.imgproc = [defs.check(x, defs.imgproc) for x in parse_seq(d.get("imgproc") or [])]
            raw = d.get("raw_image_shape")
            o.raw_image_shape = tuple(parse_seq(raw)) if raw else None
            o.verify_records = d.get("verify_records", "length")
            o.num_frames_per_clip = int(d.get("num_frames_per_clip", 16))
            o.clips_per_video = d.get("clips_per_video", 1)
            o.num_items = int(d.get("num_items", 64))
            o.seed = int(d.get("seed", 0))
            self.data.append(o)

    # ------------------------------------------------------------------------------------------------
    def engine_config(self, fpc, image_shape=None):
        """Validity rules of Model.build_pipeline (models/model.py:18-155) for the pipelines the hot path covers,
        and the EngineConfig they map to.  image_shape: the dataset's (h, w, 3) network input (model.py:54 takes it
        from the dataset as well)."""
        if len(self.pipeline_names) != 1:
            error("Only single-pipeline models (the LRCN / single-frame hot path) are built; multi-input fusion "
                  "pipelines are outside the hot path (SURVEY 8f #4)")
        net = self.pipelines[self.pipeline_names[0]]
        opt = self.train.optimizer if self.train else defs.optim.sgd
        clip = self.train.clip_norm if self.train else None
        return pipeline_engine_config(net, self.num_classes, fpc, opt, clip, self.get_dropout(), image_shape)

    def initialize(self, init_file):
        if init_file.endswith(".ini"):
            error("ini files deprecated")  # settings_.py:382-383
        with open(init_file) as f:
            config = yaml.safe_load(f)
        if "run" not in config:
            error("No [run] tag in configuration file %s" % init_file)
        self.read_config(config["run"], init_file)
        info("Starting [%s] run [%s]" % (self.phase, self.run_id))
        from .feeder import Feeder
        feeder = Feeder(self)
        return feeder


def pipeline_engine_config(net, num_classes, fpc, optimizer, clip_norm, dropout_keep_prob, image_shape=None):
    """One pipeline description (this package's Network or the reference's settings_.Network: same fields) -> the
    EngineConfig of the device path, under the validity rules of Model.build_pipeline (models/model.py:18-155)."""
    if len(net.input) != 1 or net.input_fusion is not None:
        error("Multi-input pipelines (input_fusion %s) are outside the hot path (SURVEY 8f #4)" % str(net.input_fusion))
    if net.representation != defs.representation.dcnn:
        error("Only the dcnn representation is built for the hot path (got %s)" % net.representation)
    if net.weights_file is not None and not os.path.exists(net.weights_file):
        error("Weights file %s does not exist" % net.weights_file)
    common = dict(num_classes=num_classes, fpc=fpc, optimizer=optimizer, clip_norm=clip_norm,
                  dropout_keep_prob=dropout_keep_prob)
    if image_shape is not None:
        if len(image_shape) != 3 or int(image_shape[2]) != 3:
            error("image_shape %s: the dcnn input is [h, w, 3] (model.py:54)" % str(tuple(image_shape)))
        if min(int(image_shape[0]), int(image_shape[1])) < 67:
            error("image_shape %s is too small for the AlexNet encoder (pool5 would be empty)" % str(tuple(image_shape)))
        common.update(height=int(image_shape[0]), width=int(image_shape[1]))
    fusion_type, fusion_method = net.frame_fusion if net.frame_fusion else (None, None)
    if net.classifier is None and fusion_type == defs.fusion_type.late:
        error("Specified late fusion with no classifier selected")  # model.py:37-38
    if net.classifier == defs.classifier.lstm:
        if fpc <= 1:
            error("The LSTM classifier requires an fpc greater than 1")  # model.py:121
        if fusion_type is not None and fusion_type != defs.fusion_type.none:
            error("The LSTM classifier should be used only with [%s] fusion, but it's [%s]" % (
                defs.fusion_type.none, fusion_type))  # model.py:125
        hidden, layers, fusion = net.lstm_params[:3]
        return EngineConfig(workflow="lrcn", frame_encoding_layer=net.frame_encoding_layer, lstm_hidden=hidden,
                            lstm_layers=layers, fusion=fusion, **common)
    if net.classifier != defs.classifier.fc:
        error("A classifier (fc or lstm) is required: feature-only pipelines feed other pipelines (SURVEY 8f #4)")
    # classifier fc (model.py:103-117,149-151): optional EARLY fusion of the frame features, convert_dim_fc to the class
    # count when the widths differ, optional LATE fusion of the logits
    early = fusion_method if (fusion_type == defs.fusion_type.early and fpc > 1) else None
    late = fusion_method if (fusion_type == defs.fusion_type.late and fpc > 1) else None
    if early is None and late is None and fpc > 1:
        error("A pipeline with an fc classifier and fpc %d needs frame_fusion [early|late, avg|last]: labels are per "
              "clip (dataset_.py:400-408)" % fpc)
    layer = net.frame_encoding_layer if net.frame_encoding_layer in ("fc6", "fc7") else "fc8"
    if layer == "fc8":
        # fc8 logits per frame: the fc classifier is the identity (model.py:115-117), so fusing before it (early) or
        # after it (late) is the same reduction over the clip's frames
        return EngineConfig(workflow="singleframe", fusion=late or early, **common)
    return EngineConfig(workflow="fc", frame_encoding_layer=layer, fusion=late or early, early_fusion=early is not None,
                        **common)
