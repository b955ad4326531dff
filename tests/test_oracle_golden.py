"""The numpy oracle (oracle/lrcn_numpy.py, fp32) against the committed golden vectors of the arithmetic path
(tests/golden/oracle_arith_golden.npz, written by tests/golden/make_golden_arith.py from the float64 torch-autograd
implementation): logits, loss, global gradient norm, per-variable gradient norms and gradient samples."""
import importlib.util
import os

import numpy as np
import pytest

from oracle import lrcn_numpy as O

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_golden_arith", os.path.join(HERE, "golden", "make_golden_arith.py"))
G = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(G)
GOLD = np.load(os.path.join(HERE, "golden", "oracle_arith_golden.npz"))


def rel(got, ref):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    return np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-30)


@pytest.mark.parametrize("name", sorted(G.CASES))
def test_numpy_oracle_reproduces_the_golden_vectors(name):
    cfg, params, frames, x, onehot, (wf, fusion, layer) = G.case_inputs(name)
    res = O.train_step({k: v.copy() for k, v in params.items()}, x, onehot, cfg.fpc, 1e-3, wf, fusion, layer,
                       clip_norm=cfg.clip_norm)
    assert rel(res["logits"], GOLD[name + "/logits"]) < 2e-5
    assert abs(float(res["loss"]) - float(GOLD[name + "/loss"])) < 2e-5 * max(1.0, abs(float(GOLD[name + "/loss"])))
    assert np.array_equal(np.argmax(res["logits"], 1), np.argmax(GOLD[name + "/logits"], 1))
    names = [str(n) for n in GOLD[name + "/grad_names"]]
    assert sorted(res["grads"]) == names
    # the oracle's train_step returns the CLIPPED gradients; the golden file holds the raw ones and their global norm
    gn = float(GOLD[name + "/global_norm"])
    scale = cfg.clip_norm / max(gn, cfg.clip_norm) if cfg.clip_norm else 1.0
    assert abs(float(res["global_norm"]) - gn) < 1e-3 * gn
    for i, k in enumerate(names):
        l2 = float(np.sqrt((res["grads"][k].astype(np.float64) ** 2).sum()))
        ref = float(GOLD[name + "/grad_l2"][i]) * scale
        assert abs(l2 - ref) <= 2e-3 * max(ref, 1e-12 * gn), (k, l2, ref)
        head = np.resize(res["grads"][k].reshape(-1)[:8], 8).astype(np.float64)
        ref_head = GOLD[name + "/grad_head"][i] * scale
        assert np.abs(head - ref_head).max() <= 2e-3 * max(np.abs(ref_head).max(), ref * 1e-3), k
