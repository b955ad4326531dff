"""CPU stand-in for vlb200.engine.Engine -- TEST INFRASTRUCTURE for the HOST workflow only (run_task loops, resume,
checkpoints, data-parallel sharding): it has the engine's surface (variables keyed by the reference's TF names,
train_step / forward / forward_device, state dicts) and NO model arithmetic: the loss falls deterministically with
the step and the logits are a fixed function of the frames.  The product has no CPU path; GPU tests use the real one."""
import numpy as np

import vlb200  # noqa: F401
from vlb200 import engine as E


class FakeEngine(object):
    def __init__(self, cfg, max_clips, device="cpu", rank=0, world=1):
        self.cfg, self.max_clips, self.rank, self.world = cfg, max_clips, rank, world
        self.dev = device
        self.var_shapes = E.variable_shapes(cfg)
        rng = np.random.default_rng(3)
        self.vars = {name: rng.standard_normal(shape).astype(np.float32) * 0.01 for name, shape in self.var_shapes}
        self.global_step = 0
        self.adam_t = 0
        self.adam = {}
        self.calls = []  # (kind, clips, lr or None, global_clips or None)
        self.read_resize = None

    # -- variables ---------------------------------------------------------------------------------
    def state_dict(self):
        out = {k: v.copy() for k, v in self.vars.items()}
        out["global_step"] = np.int32(self.global_step)
        return out

    def load_state_dict(self, sd, strict=True):
        for name, shape in self.var_shapes:
            if name not in sd:
                if strict:
                    raise KeyError("missing variable %s" % name)
                continue
            arr = np.asarray(sd[name], np.float32)
            if tuple(arr.shape) != tuple(shape):
                raise ValueError("variable %s: shape %s does not match %s" % (name, arr.shape, shape))
            self.vars[name] = arr.copy()
        if "global_step" in sd:
            self.global_step = int(sd["global_step"])

    def optimizer_state_dict(self):
        if self.cfg.optimizer != "adam":
            return {}
        out = {}
        for name, shape in self.var_shapes:
            out[name + "/Adam"] = self.adam.get(name + "/Adam", np.zeros(shape, np.float32))
            out[name + "/Adam_1"] = self.adam.get(name + "/Adam_1", np.zeros(shape, np.float32))
        out["beta1_power"] = np.float32(0.9 ** (self.adam_t + 1))
        out["beta2_power"] = np.float32(0.999 ** (self.adam_t + 1))
        return out

    def load_optimizer_state_dict(self, sd):
        n = 0
        for k, v in sd.items():
            if k.endswith(("/Adam", "/Adam_1")):
                self.adam[k] = np.asarray(v, np.float32).copy()
                n += 1
        if "beta1_power" in sd:
            self.adam_t = max(0, int(round(np.log(float(sd["beta1_power"])) / np.log(0.9))) - 1)
        return n

    def set_read_resize(self, hw):
        self.read_resize = hw

    # -- the two sess.run calls --------------------------------------------------------------------
    def train_step(self, frames, onehot, lr, dropout_mask=None, apply_update=True, crops=None, global_clips=None):
        clips = len(onehot)
        assert len(frames) == clips * self.cfg.fpc
        self.calls.append(("train", clips, float(lr), global_clips))
        if self.cfg.optimizer == "adam":
            self.adam_t += 1
        self.global_step += 1
        loss = float(np.log(self.cfg.num_classes)) / (1.0 + 0.1 * self.global_step)
        return loss, float(lr), self.global_step, 0.0, 1.0

    def _logits(self, frames):
        f = np.asarray(frames, np.float32).reshape(len(frames) // self.cfg.fpc, -1)
        base = f.mean(axis=1, keepdims=True) / 255.0
        cls = np.arange(self.cfg.num_classes, dtype=np.float32)[None, :]
        return np.cos(base * 37.0 + cls * 0.61).astype(np.float32)

    def forward(self, frames, crops=None):
        self.calls.append(("forward", len(frames) // self.cfg.fpc, None, None))
        return self._logits(frames)

    def forward_device(self, frames, training=False, crops=None):
        import torch
        return torch.from_numpy(self.forward(frames, crops))
