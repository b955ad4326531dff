"""Checkpoints: variables keyed by the reference's TF variable names + the `.snap` progress pickle.

Reference: feeder.py:143-194 (resume_snap), :198-257 (init_saveload), :263-288 (save).  TensorFlow's V2 checkpoint
format cannot be written without TensorFlow; variables go into `<prefix>-<global_step>.npz` (same keys), while the
folder layout, the file-name pattern `<ddmmyy_HHMMSS>_ep_E_btch_B_gs_G.graph-G`, the `.snap` pickle
`[batch_index, epoch_index, global_step]` and `resume_file: latest` keep the reference's semantics.
"""
import glob
import os
import pickle
import time

import numpy as np

from .utils import error, info, warning


def checkpoint_folder(run_folder):
    return os.path.join(run_folder, "checkpoints")


OPTIMIZER_SUFFIXES = ("/Adam", "/Adam_1")
OPTIMIZER_SCALARS = ("beta1_power", "beta2_power")


def _is_optimizer_key(key):
    return key in OPTIMIZER_SCALARS or key.endswith(OPTIMIZER_SUFFIXES)


def _write_atomic(path, writer, mode="wb"):
    """Write through a temporary file in the same folder and os.replace it: a crash never leaves a partial file
    under the final name."""
    tmp = path + ".tmp%d" % os.getpid()
    with open(tmp, mode) as f:
        writer(f)
        f.flush()
        os.fsync(f.fileno())
    os.replace(tmp, path)


def _write_variables(engine, prefix):
    sd = dict(engine.state_dict())
    sd.update(engine.optimizer_state_dict())
    _write_atomic(prefix + ".npz", lambda f: np.savez(f, **{k.replace("/", "|"): v for k, v in sd.items()}))


def read_index(folder):
    """Checkpoint prefixes named by `<folder>/checkpoint`, oldest first.  The file has TensorFlow's layout, which the
    reference's own reader expects (feeder.py:148-156 takes the quoted path of the FIRST line):
        model_checkpoint_path: "<latest prefix>"
        all_model_checkpoint_paths: "<prefix>"      (one line per kept checkpoint, oldest first)
    Plain one-name-per-line files (written by round 1 of this package) are still understood."""
    index = os.path.join(folder, "checkpoint")
    if not os.path.exists(index):
        return []
    latest, kept, plain = None, [], []
    with open(index) as f:
        for line in f:
            line = line.strip()
            if not line:
                continue
            if line.startswith("model_checkpoint_path:"):
                latest = line.split(":", 1)[1].strip().strip('"')
            elif line.startswith("all_model_checkpoint_paths:"):
                kept.append(line.split(":", 1)[1].strip().strip('"'))
            else:
                plain.append(os.path.join(folder, line))
    if latest is None:
        return plain
    if latest not in kept:
        kept.append(latest)
    return kept


def _update_index(folder, prefix, max_to_keep=None):
    prefix = os.path.abspath(prefix)
    names = [n for n in read_index(folder) if n != prefix] + [prefix]
    stale = names[:-max_to_keep] if max_to_keep else []
    if max_to_keep:
        names = names[-max_to_keep:]
    lines = ['model_checkpoint_path: "%s"' % names[-1]] + ['all_model_checkpoint_paths: "%s"' % n for n in names]
    _write_atomic(os.path.join(folder, "checkpoint"), lambda f: f.write("\n".join(lines) + "\n"), mode="w")
    for old in stale:  # only after the index stopped naming them
        for ext in (".npz", ".snap", ".meta", ".index"):
            if os.path.exists(old + ext):
                os.remove(old + ext)


def save_prefix(engine, prefix, max_to_keep=None):
    """`tf.train.Saver.save` of the shim (tfshim.py): variables + optimiser slots to `<prefix>.npz`, then the index.
    The reference's feeder.save writes the `.snap` itself afterwards (feeder.py:263-288)."""
    os.makedirs(os.path.dirname(prefix), exist_ok=True)
    _write_variables(engine, prefix)
    _update_index(os.path.dirname(prefix), prefix, max_to_keep)
    return prefix


def save(engine, run_folder, progress_str, batch_index, epoch_index, max_to_keep=None):
    """Variables + optimiser slots (tf.train.Saver saves all global variables: feeder.py:263-288) + the `.snap`
    progress pickle.  Order: weights, snap, and only then the `checkpoint` index, each through an atomic rename, so
    `resume_file: latest` never points at a partial checkpoint."""
    folder = checkpoint_folder(run_folder)
    os.makedirs(folder, exist_ok=True)
    gs = engine.global_step
    prefix = os.path.abspath(os.path.join(folder, "%s_%s.graph-%d" % (time.strftime("%d%m%y_%H%M%S"), progress_str, gs)))
    _write_variables(engine, prefix)
    _write_atomic(prefix + ".snap", lambda f: pickle.dump([batch_index, epoch_index, gs], f))
    _update_index(folder, prefix, max_to_keep)
    info("Saved checkpoint %s" % prefix)
    return prefix


def resolve(run_folder, resume_file):
    folder = checkpoint_folder(run_folder)
    if resume_file == "latest":
        if not os.path.exists(os.path.join(folder, "checkpoint")):
            error("No checkpoint index in %s to resume `latest` from" % folder)
        names = read_index(folder)
        if not names:
            error("Empty checkpoint index %s" % os.path.join(folder, "checkpoint"))
        return names[-1]
    cand = resume_file if os.path.isabs(resume_file) else os.path.join(folder, resume_file)
    for ext in (".npz", ".snap"):
        if cand.endswith(ext):
            cand = cand[:-len(ext)]
    if not os.path.exists(cand + ".npz"):
        matches = sorted(glob.glob(cand + "*.npz"))
        if not matches:
            error("Checkpoint %s not found" % cand)
        cand = matches[-1][:-4]
    return cand


def restore(engine, prefix, ignorable=("global_step",), is_validation=False, read_snap=True):
    """Load variables (name-diff check like feeder.py:229-249) and return (batch_index, epoch_index, global_step)."""
    blob = np.load(prefix + ".npz")
    sd = {k.replace("|", "/"): blob[k] for k in blob.files}
    want = {name for name, _ in engine.var_shapes}
    opt_sd = {k: sd.pop(k) for k in list(sd) if _is_optimizer_key(k)}  # optimiser slots: outside the name diff
    have = set(sd) - {"global_step"}
    missing, extra = want - have, have - want
    if missing:
        error("Variables missing from checkpoint %s: %s" % (prefix, sorted(missing)))
    if extra:
        warning("Checkpoint variables not in the model (ignored): %s" % sorted(extra))
    if is_validation:
        sd.pop("global_step", None)  # ignorable in validation (feeder.py:226-227)
    engine.load_state_dict(sd)
    if not is_validation:  # ignorable in validation, like global_step
        n_slots = engine.load_optimizer_state_dict(opt_sd)
        if engine.cfg.optimizer == "adam" and n_slots == 0:
            warning("Checkpoint %s holds no Adam slots: the moments restart from zero" % prefix)
    snap = [0, 0, int(sd.get("global_step", 0))]
    if read_snap and os.path.exists(prefix + ".snap"):
        with open(prefix + ".snap", "rb") as f:
            snap = pickle.load(f)
    info("Restored %s: batch %d, epoch %d, global step %d" % (prefix, snap[0], snap[1], snap[2]))
    return snap


def load_alexnet_npy(engine, path, skip=("fc8",)):
    """Initialise the encoder from the Caffe-converted `bvlc_alexnet.npy` the reference uses (alexnet.py:50-52,
    69-71: a pickled dict layer -> [W, b], HWIO filters).  Layers whose shape does not match the model (fc8 for a
    class count other than 1000) or that are listed in `skip` keep their current values."""
    blob = np.load(path, allow_pickle=True, encoding="latin1").item()
    shapes = dict(engine.var_shapes)
    sd = {}
    for layer, wb in blob.items():
        if layer in skip:
            continue
        for suffix, arr in zip(("W", "b"), wb):
            name = "dcnn/%s%s" % (layer, suffix)
            arr = np.asarray(arr, dtype=np.float32)
            if name in shapes and tuple(shapes[name]) == tuple(arr.shape):
                sd[name] = arr
            elif name in shapes:
                warning("bvlc_alexnet.npy: %s has shape %s, model wants %s (kept random init)" % (
                    name, arr.shape, shapes[name]))
    engine.load_state_dict(sd, strict=False)
    info("Loaded %d variables from %s" % (len(sd), path))
    return sorted(sd)
