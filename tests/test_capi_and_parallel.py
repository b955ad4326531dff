"""C-ABI surface (no compute without a GPU) and the data-parallel host logic over gloo (world_size 2, CPU)."""
import os
import re
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import vlb200  # noqa: F401
from vlb200 import _native as nv
from vlb200 import parallel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    with open(os.path.join(ROOT, "include", "vlb200.h")) as f:
        hdr = f.read()
    declared = set(re.findall(r"\b(vl_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(nv.EXPORTS)
    lib = nv.lib()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.vl_version() >= 100
    assert isinstance(lib.vl_last_error(), bytes)


def test_engine_fails_loudly_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from vlb200 import engine as E
    with pytest.raises(nv.NativeError):
        E.Engine(E.EngineConfig(), max_clips=1)


def test_variable_inventory_matches_survey():
    from vlb200 import engine as E
    shapes = dict(E.variable_shapes(E.EngineConfig()))
    assert sum(int(np.prod(s)) for s in shapes.values()) == 61351653  # SURVEY 2.3 parameter count (config-2)
    assert shapes["dcnn/conv2W"] == (5, 5, 48, 256) and shapes["dcnn/fc6W"] == (9216, 4096)
    assert shapes["rnn/multi_rnn_cell/cell_0/basic_lstm_cell/kernel"] == (4352, 1024)
    assert shapes["output_fc_w"] == (256, 101)
    sp = E.encoder_specs(227, 227)
    assert (sp["conv1"].p, sp["conv1"].pad_top, sp["pool1"], sp["pool2"], sp["pool5"]) == (57, 4, (28, 28), (13, 13), (6, 6))


def test_shard_range_covers_all_items():
    for n in (1, 7, 64, 65):
        for w in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _dp_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    from oracle import lrcn_numpy as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # tiny single-frame-free problem: the head of the path (pool -> fc -> CE) is enough to exercise the DP math
    rng = np.random.default_rng(3)
    clips, t_len, d, c = 4, 3, 8, 5
    x = rng.standard_normal((clips, t_len, d)).astype(np.float32)
    w = rng.standard_normal((d, c)).astype(np.float32)
    y = np.zeros((clips, c), np.int32)
    y[np.arange(clips), rng.integers(0, c, clips)] = 1
    lo, hi = parallel.shard_range(clips, rank, world)

    def grads(xs, ys, scale):
        fused = O.temporal_fusion(xs, "avg")
        logits = fused @ w
        loss, dl, _ = O.softmax_ce(logits, ys)
        dl = dl * (scale * len(xs))  # softmax_ce scales by 1/local rows; rescale to the DP convention
        return loss, fused.T @ dl, logits

    loss_l, g_l, logits_l = grads(x[lo:hi], y[lo:hi], parallel.local_grad_scale(hi - lo, world))
    flat = torch.from_numpy(g_l.reshape(-1).copy())
    scal = torch.tensor([float(loss_l), 0.0])
    # the engine's schedule: the tail of the arena is reduced asynchronously first, the head + scalars later
    cut = flat.numel() // 3
    early = parallel.allreduce_async(flat[cut:])
    parallel.allreduce_gradients(flat[:cut], scal)
    parallel.wait(early)
    gathered = parallel.gather_logits(torch.from_numpy(logits_l))
    loss_ref, g_ref, logits_ref = grads(x, y, 1.0 / clips)
    ok = (np.allclose(flat.numpy().reshape(d, c), g_ref, atol=1e-6) and
          abs(scal[0].item() / world - float(loss_ref)) < 1e-6 and
          np.allclose(gathered.numpy(), logits_ref, atol=1e-6))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_data_parallel_math_over_gloo_world2():
    """2 ranks on CPU: summed shard gradients == gradient of the global-batch mean loss; logits gather in order."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, True), (1, True)]
