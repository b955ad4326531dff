"""The two graph-building classes of the reference, over the CUDA engine: the Python-level seam of SURVEY 8b.

    reference                                         here
    models/model.py:157-175  Model(settings)          Model(settings)  .logits .required_input .get_output()
                                                                        .get_ignorable_variable_names()
    train.py:112-149  Train(settings, feeder,         Train(settings, feeder, logits, summaries)
                            logits, summaries)          .loss .current_lr .global_step .optimizer .accuracyTrain
                                                        .grads_norm .labels .required_input

Both take the REFERENCE's own objects (its `Settings`, `Feeder`, `Summaries`): a maintainer replaces two import lines
of run_task.py and installs the `tensorflow` stand-in (tfshim.py); `sess.run([...], feed_dict)` then lands in
Engine.train_step / Engine.forward.  Everything else of run_task.py -- Settings, Feeder, Dataset, Validation, the
loops, the save cadence -- is the reference's code, unmodified (tests/test_reference_dropin.py runs exactly that).
"""
import os

import numpy as np

from . import tfshim as tf
from .defs import defs
from .settings import pipeline_engine_config
from .train import precompute_learning_rates
from .utils import error, info


def _dataset(settings, tag):
    dsets = settings.feeder.get_dataset_by_tag(tag)
    if not dsets:
        error("Could not find a dataset with the tag {}, required by the pipeline".format(tag))
    return dsets[0]


class Model(object):
    """models/model.py: builds the pipelines from the settings.  Here: validates the (single, dcnn) pipeline with the
    rules of Model.build_pipeline and keeps the recipe of the engine; the engine itself is created by the session."""

    def __init__(self, settings):
        self.settings = settings
        self.required_input = []
        if len(settings.pipeline_names) != 1:
            error("Only single-pipeline models (the LRCN / single-frame hot path) are built; multi-input fusion "
                  "pipelines are outside the hot path (SURVEY 8f #4)")
        name = settings.pipeline_names[-1]
        self.pipeline = settings.pipelines[name]
        tag = self.pipeline.input[0]
        ds = _dataset(settings, tag)
        shape = tuple(self.pipeline.input_shape[0]) if self.pipeline.input_shape and self.pipeline.input_shape[0] \
            else tuple(ds.get_image_shape())  # model.py:50-54
        self.fpc = int(ds.num_frames_per_clip)
        train = settings.train
        self.cfg = pipeline_engine_config(self.pipeline, int(settings.num_classes), self.fpc,
                                          train.optimizer if train else defs.optim.sgd,
                                          train.clip_norm if train else None, settings.get_dropout(), shape)
        self.cfg.mean = None  # the reference's Dataset hands over float32 frames that are already mean-subtracted
        self.input = tf.placeholder(tf.float32, (None,) + shape, name="%s_%s_input" % (name, tag))
        self.required_input.append((self.input, defs.net_input.visual, tag))
        self.logits = tf.Tensor("logits", "logits", (None, int(settings.num_classes)))
        # capacity: the largest batch in clips (whole items per batch, dataset_.py:562-613)
        cpv = ds.clips_per_video if isinstance(ds.clips_per_video, (list, tuple)) else [ds.clips_per_video] * ds.num_items
        bs = int(settings.get_batch_size())
        self.max_clips = max(int(sum(cpv[i:i + bs])) for i in range(0, len(cpv), bs))
        tf.get_default_graph().reset()
        tf.get_default_graph().model = self
        tf.get_default_graph().placeholders.append(self.input)
        info("vlb200.compat.Model: pipeline [%s] -> %s engine, %d frames per clip, up to %d clips per batch" % (
            name, self.cfg.workflow, self.fpc, self.max_clips))

    def get_output(self):
        return self.logits

    def get_ignorable_variable_names(self):
        return []

    def variable_shapes(self):
        from .engine import variable_shapes
        return variable_shapes(self.cfg)

    # -- used by tfshim.Session ---------------------------------------------------------------------
    def make_engine(self, factory=None):
        if factory is None:
            from .engine import Engine
            engine = Engine(self.cfg, max_clips=self.max_clips)
        else:
            engine = factory(self.cfg, self.max_clips, "cuda:0", 0, 1)
        if self.pipeline.weights_file is not None:
            from . import checkpoint
            checkpoint.load_alexnet_npy(engine, self.pipeline.weights_file)  # alexnet.py:50-52,69-71
        engine.global_step = int(getattr(self.settings, "global_step", 0) or 0)
        return engine

    @staticmethod
    def _frames(feed):
        """feed_dict[placeholder] = list of N float32 HWC arrays (feeder.py:97-100)."""
        return np.ascontiguousarray(np.stack([np.asarray(f, dtype=np.float32) for f in feed], axis=0)) \
            if isinstance(feed, (list, tuple)) else np.asarray(feed, dtype=np.float32)

    def run_forward(self, engine, feed_dict, train, kinds):
        if self.input not in feed_dict:
            raise ValueError("feed_dict lacks the model input placeholder %s" % self.input.name)
        return {"logits": engine.forward(self._frames(feed_dict[self.input]))}


class Train(object):
    """train.py:112-149: loss, learning-rate lookup, optimiser step, training accuracy -- as fetch handles."""

    def __init__(self, settings, feeder, logits, summaries):
        self.required_input = []
        if not settings.train:
            return
        tr = settings.train
        if tr.lr_mult is not None:
            error("lr_mult is not supported: the reference's multi-tier learning is inoperative (train.py:36-37,152-197)")
        if tr.optimizer not in (defs.optim.sgd, defs.optim.adam):
            error("Undefined optimizer %s" % tr.optimizer)
        self.labels = tf.placeholder(tf.int32, [None, settings.num_classes], name="input_labels")
        self.required_input.append((self.labels, defs.net_input.labels, defs.dataset_tag.main))
        num_batches = feeder.get_num_batches()
        schedule = os.path.join(settings.run_folder, settings.run_id + "_lr_decay_schedule.txt")
        self.learning_rates = precompute_learning_rates(tr.base_lr, tr.lr_decay, num_batches, tr.epochs, schedule)
        if len(self.learning_rates) != num_batches * tr.epochs:
            error("Batch length precomputation mismatch")
        self.loss = tf.Tensor("loss", "cross_entropy_loss/total/Mean")
        self.current_lr = tf.Tensor("lr", "lr/strided_slice")
        self.global_step = tf.Tensor("global_step", "global_step", dtype=tf.int32)
        self.optimizer = tf.Operation("optimizer", "optimizer")
        self.accuracyTrain = tf.Tensor("accuracy", "training_accuracy/accuracy_train/Mean")
        self.grads_norm = tf.Tensor("grads_norm", "grads_norm")
        for s in ("loss", "lr", "accuracyTrain", "grads_norm"):
            summaries.train.append(tf.summary.scalar(s))
        tf.get_default_graph().train = self
        tf.get_default_graph().placeholders.append(self.labels)

    def lr_at(self, step):
        """lr = table[global_step] with the pre-increment step (train.py:129-132)."""
        if step >= len(self.learning_rates):
            error("global step %d exceeds the learning rate table (%d)" % (step, len(self.learning_rates)))
        return float(self.learning_rates[step])

    def run_step(self, engine, model, feed_dict):
        if model.input not in feed_dict or self.labels not in feed_dict:
            raise ValueError("feed_dict lacks the input / label placeholders")
        frames = Model._frames(feed_dict[model.input])
        onehot = np.asarray(feed_dict[self.labels], dtype=np.int32)
        loss, lr, gstep, acc, gnorm = engine.train_step(frames, onehot, self.lr_at(engine.global_step))
        return {"loss": np.float32(loss), "lr": np.float32(lr), "global_step": int(gstep), "accuracy": np.float32(acc),
                "grads_norm": np.float32(gnorm)}
