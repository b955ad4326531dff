"""`tensorflow` stand-in for the reference's HOST-side files: the drop-in seam of SURVEY 8b made executable.

The reference drives its hot path through TensorFlow 1.x objects: placeholders, `sess.run(fetches, feed_dict)`,
`tf.train.Saver`, `tf.python_io.tf_record_iterator` / `tf.train.Example`, summaries.  This module provides exactly the
symbols the reference's `run_task.py`, `settings_.py`, `feeder.py`, `dataset_.py`, `val.py`, `utils_.py` and
`tools/inspect_checkpoint.py` touch at run time, as thin handles over the CUDA engine:

    import vlb200; vlb200.tfshim.install()          # before the reference's modules are imported
    from vlb200.compat import Model, Train          # instead of models.model.Model / train.Train (the graph builders)
    ... the rest of run_task.py (Settings, Feeder, Dataset, Validation, do_train, do_test, feeder.save /
        init_saveload) runs unmodified: sess.run([...]) lands in Engine.train_step / Engine.forward.

It is NOT a TensorFlow implementation: there are no ops and no graph.  The two graph-building classes of the reference
(`Model`, `Train`) are replaced by `vlb200.compat`; everything they would have computed runs in the sm_100a kernels.
`tests/test_reference_dropin.py` runs the reference's own `run_task.main` through this shim.
"""
import os
import sys
import types

import numpy as np

float32 = "float32"
int32 = "int32"
int64 = "int64"


class _Graph(object):
    """What `Model` / `Train` registered: the engine recipe and the fetch handles `Session.run` resolves."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.model = None       # compat.Model
        self.train = None       # compat.Train
        self.session = None
        self.placeholders = []


_graph = _Graph()


def get_default_graph():
    return _graph


def reset_default_graph():
    """feeder.py:251 calls this right before saver.restore: the handles stay valid (there is no graph to rebuild)."""
    return None


class Tensor(object):
    """A fetch / feed handle (placeholder, logits, loss, ...).  `kind` tells Session.run what to do with it."""

    def __init__(self, kind, name, shape=None, dtype=float32):
        self.kind, self.name, self.dtype = kind, name + ":0", dtype
        self.shape = tuple(shape) if shape is not None else None
        self.graph = _graph

    def __repr__(self):
        return "<vlb200.tfshim.Tensor %s %s>" % (self.kind, self.name)

    def __hash__(self):
        return id(self)

    def __eq__(self, other):
        return self is other


class Operation(Tensor):
    pass


def placeholder(dtype, shape=None, name=None):
    t = Tensor("placeholder", name or "Placeholder", shape, dtype)
    _graph.placeholders.append(t)
    return t


def global_variables_initializer():
    return Operation("init", "init")


class _Var(object):
    def __init__(self, name):
        self.name = name + ":0"


def global_variables():
    """Names of the model variables (+ global_step, + optimiser slots): the name diff of feeder.py:229-249."""
    if _graph.model is None:
        return []
    names = [n for n, _ in _graph.model.variable_shapes()] + ["global_step"]
    return [_Var(n) for n in names]


all_variables = global_variables


def trainable_variables():
    return global_variables()[:-1]


class errors(object):
    class NotFoundError(Exception):
        pass

    class OutOfRangeError(Exception):
        pass


class _NameScope(object):
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        return self.name

    def __exit__(self, *a):
        return False


def name_scope(name, *a, **k):
    return _NameScope(name)


variable_scope = name_scope


# ----------------------------------------------------------------------------------------------------------
# summaries: accepted and dropped (tensorboard is outside the hot path)
# ----------------------------------------------------------------------------------------------------------
class summary(object):
    @staticmethod
    def scalar(name, tensor=None, *a, **k):
        return Tensor("summary", "summary/" + str(name))

    @staticmethod
    def histogram(name, tensor=None, *a, **k):
        return Tensor("summary", "summary/" + str(name))

    @staticmethod
    def merge(inputs, *a, **k):
        return Tensor("summary", "summary/merged")

    @staticmethod
    def merge_all(*a, **k):
        return Tensor("summary", "summary/merged")

    class FileWriter(object):
        def __init__(self, logdir=None, graph=None, *a, **k):
            self.logdir = logdir

        def add_summary(self, *a, **k):
            pass

        def flush(self):
            pass

        def close(self):
            pass


# ----------------------------------------------------------------------------------------------------------
# tf.python_io / tf.train.Example: the serialized datasets, read without TensorFlow (tfrecord.py)
# ----------------------------------------------------------------------------------------------------------
class _RecordIterator(object):
    def __init__(self, path):
        from . import tfrecord
        self._it = tfrecord.read_records(path, "length")

    def __iter__(self):
        return self

    def __next__(self):
        return next(self._it)

    def close(self):
        self._it = iter(())


class python_io(object):
    @staticmethod
    def tf_record_iterator(path=None, options=None):
        return _RecordIterator(path)


class _ValueList(object):
    def __init__(self, values):
        self.value = values


class _Feature(object):
    def __init__(self, values):
        is_bytes = bool(values) and isinstance(values[0], (bytes, bytearray))
        is_float = bool(values) and isinstance(values[0], float)
        self.bytes_list = _ValueList(values if is_bytes else [])
        self.float_list = _ValueList(values if is_float else [])
        self.int64_list = _ValueList(values if not (is_bytes or is_float) else [])


class _Features(object):
    def __init__(self, parsed):
        self.feature = {k: _Feature(v) for k, v in parsed.items()}


class _Example(object):
    def __init__(self):
        self.features = _Features({})

    def ParseFromString(self, payload):
        from . import tfrecord
        self.features = _Features(tfrecord.parse_example(payload))


# ----------------------------------------------------------------------------------------------------------
# tf.train.Saver: checkpoints keyed by the TF variable names (checkpoint.py), plus the files the reference's
# feeder.init_saveload insists on (`<prefix>.meta`, `<prefix>.index`) and the `checkpoints/checkpoint` index
# ----------------------------------------------------------------------------------------------------------
class _Saver(object):
    def __init__(self, var_list=None, max_to_keep=5, *a, **k):
        self.max_to_keep = max_to_keep

    def save(self, sess, save_path, global_step=None, *a, **k):
        from . import checkpoint
        prefix = save_path if global_step is None else "%s-%d" % (save_path, int(global_step))
        checkpoint.save_prefix(sess.engine, prefix, self.max_to_keep)
        for ext in (".meta", ".index"):  # existence is all feeder.py:217-222 asks of them
            with open(prefix + ext, "w") as f:
                f.write("vlb200 checkpoint: variables in %s.npz\n" % os.path.basename(prefix))
        return prefix

    def restore(self, sess, save_path):
        from . import checkpoint
        if not os.path.exists(save_path + ".npz"):
            raise errors.NotFoundError("no checkpoint at %s" % save_path)
        checkpoint.restore(sess.engine, save_path, is_validation=_graph.train is None, read_snap=False)


class train(object):
    Saver = _Saver
    Example = _Example


def _checkpoint_tensor_names(file_name):
    """tools/inspect_checkpoint.get_checkpoint_tensor_names through pywrap_tensorflow.NewCheckpointReader."""
    blob = np.load(file_name + ".npz")
    return {k.replace("|", "/"): tuple(blob[k].shape) for k in blob.files}


class _CheckpointReader(object):
    def __init__(self, file_name):
        self._shapes = _checkpoint_tensor_names(file_name)
        self._file = file_name

    def get_variable_to_shape_map(self):
        return dict(self._shapes)

    def get_tensor(self, name):
        return np.load(self._file + ".npz")[name.replace("/", "|")]

    def debug_string(self):
        return "\n".join("%s %s" % kv for kv in sorted(self._shapes.items())).encode()


# ----------------------------------------------------------------------------------------------------------
# tf.Session
# ----------------------------------------------------------------------------------------------------------
class Session(object):
    """`sess.run(fetches, feed_dict)` over the CUDA engine.

    run_task.py:44   sess.run([summaries, loss, lr, global_step, optimizer], fdict)  -> Engine.train_step
    run_task.py:95   sess.run(model.logits, fdict)                                   -> Engine.forward
    run_task.py:133  sess.run(tf.global_variables_initializer())                     -> the engine is built (random init)
    """

    engine_factory = None  # tests of the host workflow inject a CPU stand-in; None = the CUDA Engine (no fallback)

    def __init__(self, *a, **k):
        self.graph = _graph
        self._engine = None
        _graph.session = self

    @property
    def engine(self):
        if self._engine is None:
            if _graph.model is None:
                raise RuntimeError("vlb200.tfshim.Session: no vlb200.compat.Model has been built")
            self._engine = _graph.model.make_engine(Session.engine_factory)
        return self._engine

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def run(self, fetches, feed_dict=None, *a, **k):
        single = not isinstance(fetches, (list, tuple))
        flist = [fetches] if single else list(fetches)
        kinds = {f.kind for f in flist}
        values = {}
        if "init" in kinds:
            self.engine  # build (variables get the reference's random initialisation)
        if "optimizer" in kinds:
            values = _graph.train.run_step(self.engine, _graph.model, feed_dict or {})
        elif kinds & {"logits", "loss", "accuracy"}:
            values = _graph.model.run_forward(self.engine, feed_dict or {}, _graph.train, kinds)
        out = []
        for f in flist:
            if f.kind in ("summary", "init"):
                out.append(b"" if f.kind == "summary" else None)
            elif f.kind == "global_step":
                out.append(values.get("global_step", self.engine.global_step))
            elif f.kind == "lr":
                out.append(values["lr"] if "lr" in values else _graph.train.lr_at(self.engine.global_step))
            elif f.kind in values:
                out.append(values[f.kind])
            elif f.kind == "optimizer":
                out.append(None)
            else:
                raise ValueError("vlb200.tfshim.Session.run: cannot fetch %r" % (f,))
        return out[0] if single else out


InteractiveSession = Session


def install():
    """Make `import tensorflow` resolve to this module (and `tensorflow.python.pywrap_tensorflow` to the checkpoint
    reader) for the reference's host files.  Call before importing them."""
    me = sys.modules[__name__]
    sys.modules["tensorflow"] = me
    python = types.ModuleType("tensorflow.python")
    pywrap = types.ModuleType("tensorflow.python.pywrap_tensorflow")
    pywrap.NewCheckpointReader = _CheckpointReader
    python.pywrap_tensorflow = pywrap
    sys.modules["tensorflow.python"] = python
    sys.modules["tensorflow.python.pywrap_tensorflow"] = pywrap
    me.python = python
    return me


__version__ = "1.x-shim (vlb200)"
