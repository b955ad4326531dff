"""Training head on the host: learning-rate table, optimiser selection, the `Train` facade run_task uses.

Reference: train.py:50-109 (precompute_learning_rates), :112-149 (Train.__init__), :199-222 (single tier learning).
The arithmetic of the step (loss, gradients, clip, apply) runs on the device inside Engine.train_step.
"""
import math
import os

from .defs import defs
from .utils import error, info


def precompute_learning_rates(base_lr, lr_decay, num_batches, epochs, schedule_file=None):
    """Per-global-step learning rates, `num_batches * epochs` entries (train.py:50-109).

    lr_decay = [strategy, scheme, freq, factor(, constant-lr offset)].  As in the reference the running index
    advances by `freq` per segment, so exp and staircase both give base * factor**k, held for `period` steps where
    period = freq (interval) or ceil(total / freq) (drops)."""
    total = num_batches * epochs
    if lr_decay is None:
        return [base_lr for _ in range(total)]
    decay = tuple(lr_decay)
    offset = 0 if len(decay) == 4 else decay[-1]
    strategy, scheme, freq, factor = decay[:4]
    if strategy == defs.decay.exp:
        staircase = False
    elif strategy == defs.decay.staircase:
        staircase = True
    else:
        error("Undefined decay strategy %s" % strategy)
    if scheme == defs.periodicity.interval:
        period = freq
    elif scheme == defs.periodicity.drops:
        period = math.ceil(total / freq)
    else:
        error("Undefined decay scheme %s" % scheme)
    table = []
    idx = 0
    while len(table) < total:
        fraction = idx // freq if staircase else idx / freq
        table.extend([base_lr * pow(factor, fraction)] * period)
        idx += freq
    table = table[:total]
    if offset:
        table = [base_lr] * offset + table[0:-offset]
    if schedule_file:
        with open(schedule_file, "w") as f:
            step = 0
            for ep in range(epochs):
                for b in range(num_batches):
                    f.write("Epoch %d/%d, batch %d/%d, lr %2.8f\n" % (ep + 1, epochs, b + 1, num_batches, table[step]))
                    step += 1
    return table


class Train(object):
    """What run_task.do_train needs from train.py: the LR table indexed by global_step and the step itself."""

    def __init__(self, settings, feeder, engine):
        self.engine = engine
        self.settings = settings
        if not settings.train:
            return
        tr = settings.train
        if tr.lr_mult is not None:
            # the reference's two-tier path is dead code (train.py:152-197 with empty variable lists)
            error("lr_mult is not supported: the reference's multi-tier learning is inoperative (train.py:36-37,152-197)")
        if tr.optimizer not in (defs.optim.sgd, defs.optim.adam):
            error("Undefined optimizer %s" % tr.optimizer)
        num_batches = feeder.get_num_batches()
        schedule = os.path.join(settings.run_folder, settings.run_id + "_lr_decay_schedule.txt")
        self.learning_rates = precompute_learning_rates(tr.base_lr, tr.lr_decay, num_batches, tr.epochs, schedule)
        if len(self.learning_rates) != num_batches * tr.epochs:
            error("Batch length precomputation mismatch")
        info("Learning rate table: %d steps, first %2.8f, last %2.8f" % (
            len(self.learning_rates), self.learning_rates[0], self.learning_rates[-1]))

    def current_lr(self):
        """lr = table[global_step] with the pre-increment step (train.py:129-132)."""
        step = self.engine.global_step
        if step >= len(self.learning_rates):
            error("global step %d exceeds the learning rate table (%d)" % (step, len(self.learning_rates)))
        return float(self.learning_rates[step])

    def step(self, frames, onehot, crops=None, global_clips=None):
        """(loss, lr, global_step, accuracyTrain, grads_norm) -- the fetches of run_task.py:29,44."""
        return self.engine.train_step(frames, onehot, self.current_lr(), crops=crops, global_clips=global_clips)
