"""Golden vectors of the ARITHMETIC path (SURVEY 8c "golden vectors we must generate ourselves" (1)): outputs of the
float64 torch-autograd evaluation of the reference's graph (oracle/lrcn_torch.py, an implementation independent of
oracle/lrcn_numpy.py and of the CUDA kernels) for seeded random-init weights and seeded inputs, tiny cases of both
workflows.  TensorFlow 1.x cannot run here, so these pin the restatement against itself across implementations and
against regressions - NOT against TensorFlow (arithmetic parity stays "unpinned").

    tests/golden/oracle_arith_golden.npz

Weights and frames are NOT stored: they are regenerated from the seeds (engine.init_variables / numpy default_rng).
Usage: python tests/golden/make_golden_arith.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import vlb200  # noqa: E402,F401
from vlb200 import engine as E  # noqa: E402
from oracle import lrcn_torch as T  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (EngineConfig kwargs, clips, oracle workflow, oracle fusion, layer, weight seed, input seed, unit-range input)
    "lrcn_avg": (dict(workflow="lrcn", fusion="avg", fpc=2, num_classes=101, lstm_hidden=256, clip_norm=10), 2, "lrcn", "avg",
                 "fc7", 7, 8, True),
    "lrcn_last_2layer": (dict(workflow="lrcn", fusion="last", fpc=3, num_classes=11, lstm_hidden=32, lstm_layers=2,
                              clip_norm=10), 2, "lrcn", "last", "fc7", 3, 4, True),
    "singleframe_u8": (dict(workflow="singleframe", fusion="avg", fpc=2, num_classes=101, clip_norm=10,
                            mean=(99.197148, 105.293620, 109.503945)), 2, "singleframe", "avg", "fc8", 5, 6, False),
}


def case_inputs(name):
    kw, clips, wf, fusion, layer, wseed, iseed, unit = CASES[name]
    cfg = E.EngineConfig(**kw)
    params = E.init_variables(cfg, seed=wseed)
    rng = np.random.default_rng(iseed)
    n = clips * cfg.fpc
    if unit:
        frames = rng.uniform(-1, 1, size=(n, 227, 227, 3)).astype(np.float32)
        x = frames
    else:
        frames = rng.integers(0, 256, size=(n, 227, 227, 3), dtype=np.uint8)
        x = frames.astype(np.float32) - np.array(cfg.mean, np.float32)
    labels = rng.integers(0, cfg.num_classes, clips)
    onehot = np.zeros((clips, cfg.num_classes), np.int32)
    onehot[np.arange(clips), labels] = 1
    return cfg, params, frames, x, onehot, (wf, fusion, layer)


def main():
    out = {}
    for name in CASES:
        cfg, params, frames, x, onehot, (wf, fusion, layer) = case_inputs(name)
        P = T.to_torch(params, requires_grad=True, dtype=torch.float64)
        res = T.train_step(P, torch.tensor(x, dtype=torch.float64), torch.tensor(onehot), cfg.fpc, 1e-3, wf, fusion, layer,
                           clip_norm=cfg.clip_norm)
        out[name + "/logits"] = res["logits"].numpy().astype(np.float64)
        out[name + "/loss"] = np.float64(res["loss"])
        out[name + "/global_norm"] = np.float64(res["global_norm"])
        names = sorted(res["grads"])
        out[name + "/grad_names"] = np.array(names)
        out[name + "/grad_l2"] = np.array([float(res["grads"][k].double().norm()) for k in names], np.float64)
        # a fixed sample of gradient entries per variable (first 8 of the flattened tensor)
        out[name + "/grad_head"] = np.stack([np.resize(res["grads"][k].reshape(-1)[:8].numpy(), 8) for k in names])
        print("%s: loss %.6f, global norm %.4e, |logits| max %.3e" % (name, res["loss"], res["global_norm"],
                                                                      np.abs(out[name + "/logits"]).max()))
    np.savez_compressed(os.path.join(OUT, "oracle_arith_golden.npz"), **out)


if __name__ == "__main__":
    main()
