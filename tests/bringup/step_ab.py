"""A/B of whole train steps (config-2: 64 clips x 16 frames) under environment switches, interleaved in ONE process on ONE
box (boxes of the pool differ by +-4 %).  Usage: python step_ab.py NAME=VALUE[,NAME=VALUE...] ... (one variant per arg;
the empty variant "-" is the baseline)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import numpy as np
import torch
import vlb200  # noqa
from vlb200 import engine as E

variants = sys.argv[1:] or ["-"]
clips, fpc = 64, 16
cfg = E.EngineConfig(workflow="lrcn", fusion="avg", fpc=fpc, num_classes=101, lstm_hidden=256, optimizer="sgd", clip_norm=10,
                     dropout_keep_prob=0.5, mean=(99.197148, 105.293620, 109.503945))
eng = E.Engine(cfg, max_clips=clips)
g = torch.Generator(device="cuda").manual_seed(0)
frames = torch.randint(0, 256, (clips * fpc, 227, 227, 3), dtype=torch.uint8, device="cuda", generator=g)
onehot = torch.zeros(clips, 101, dtype=torch.int32, device="cuda")
onehot[torch.arange(clips), torch.randint(0, 101, (clips,), device="cuda")] = 1

def apply(v):
    keys = []
    if v != "-":
        for kv in v.split(","):
            k, val = kv.split("=")
            os.environ[k] = val
            keys.append(k)
    return keys

def run(steps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        eng.train_step(frames, onehot, 1e-3, frames_ready=True)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps

for _ in range(5):
    eng.train_step(frames, onehot, 1e-3)
results = {v: [] for v in variants}
for rep in range(4):
    for v in variants:
        keys = apply(v)
        run(3)
        results[v].append(run(20))
        for k in keys:
            os.environ.pop(k)
base = min(results[variants[0]])
for v in variants:
    r = results[v]
    print("%-50s min %.3f ms  median %.3f ms  (%.1f %% vs first variant)  %6.0f clips/s" % (
        v, min(r), sorted(r)[len(r) // 2], 100 * (min(r) / base - 1), clips / min(r) * 1e3))
