"""The CUDA path against the COMMITTED golden vectors of the arithmetic path (tests/golden/oracle_arith_golden.npz):
bf16 path within 2e-2, fp32-accuracy mode within 1e-3 (north-star tolerances), labels identical, loss of one train step."""
import importlib.util
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_golden_arith", os.path.join(HERE, "golden", "make_golden_arith.py"))
G = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(G)
GOLD = np.load(os.path.join(HERE, "golden", "oracle_arith_golden.npz"))


def rel(got, ref):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    return np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-30)


@pytest.mark.parametrize("name", sorted(G.CASES))
def test_device_path_against_the_golden_vectors(name):
    import vlb200  # noqa: F401
    from vlb200 import engine as E
    cfg, params, frames, x, onehot, _ = G.case_inputs(name)
    ref = GOLD[name + "/logits"]
    eng = E.Engine(cfg, max_clips=onehot.shape[0], params=params)
    e_bf = rel(eng.forward(frames), ref)
    e_32 = rel(eng.forward_fp32(frames), ref)
    print("%s: bf16 path %.3e, fp32 mode %.3e" % (name, e_bf, e_32))
    assert e_bf < 2e-2 and e_32 < 1e-3
    srt = np.sort(ref, axis=1)
    ok = (srt[:, -1] - srt[:, -2]) > 4e-2 * np.abs(ref).max()
    assert np.array_equal(eng.forward(frames).argmax(1)[ok], ref.argmax(1)[ok])
    assert np.array_equal(eng.forward_fp32(frames).argmax(1), ref.argmax(1))
    loss, _, step, _, _ = eng.train_step(frames, onehot, 1e-3)
    gl = float(GOLD[name + "/loss"])
    assert abs(loss - gl) < 2e-2 * max(1.0, abs(gl)) and step == 1
