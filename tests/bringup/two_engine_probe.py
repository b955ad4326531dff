"""Upper bound of micro-batch pipelining: two independent 32-clip engines driven by two host threads on two streams
against one 64-clip engine (config-2 model).  If the pair is not clearly faster, interleaving two half-batch train steps
inside one engine cannot pay either."""
import os, sys, threading, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import torch
import vlb200  # noqa
from vlb200 import engine as E

fpc = 16
cfg = E.EngineConfig(workflow="lrcn", fusion="avg", fpc=fpc, num_classes=101, lstm_hidden=256, optimizer="sgd", clip_norm=10,
                     dropout_keep_prob=0.5, mean=(99.197148, 105.293620, 109.503945))

def make(clips):
    eng = E.Engine(cfg, max_clips=clips)
    g = torch.Generator(device="cuda").manual_seed(0)
    frames = torch.randint(0, 256, (clips * fpc, 227, 227, 3), dtype=torch.uint8, device="cuda", generator=g)
    onehot = torch.zeros(clips, 101, dtype=torch.int32, device="cuda")
    onehot[torch.arange(clips), torch.randint(0, 101, (clips,), device="cuda")] = 1
    return eng, frames, onehot

def run(eng, frames, onehot, steps, stream):
    with torch.cuda.stream(stream):
        for _ in range(steps):
            eng.train_step(frames, onehot, 1e-3, frames_ready=True)
        stream.synchronize()

steps = 60
one = make(64)
s0 = torch.cuda.Stream()
run(*one, 5, s0)
torch.cuda.synchronize()
t0 = time.perf_counter(); run(*one, steps, s0); torch.cuda.synchronize(); t1 = time.perf_counter()
print("one engine, 64 clips : %.3f ms per 64 clips  %7.0f clips/s" % ((t1 - t0) / steps * 1e3, 64 * steps / (t1 - t0)), flush=True)
del one
torch.cuda.empty_cache()
a, b = make(32), make(32)
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
run(*a, 5, sa); run(*b, 5, sb)
torch.cuda.synchronize()
t0 = time.perf_counter(); run(*a, steps, sa); torch.cuda.synchronize(); t1 = time.perf_counter()
print("one engine, 32 clips : %.3f ms per 32 clips  %7.0f clips/s" % ((t1 - t0) / steps * 1e3, 32 * steps / (t1 - t0)), flush=True)
for rep in range(2):
    ta = threading.Thread(target=run, args=(*a, steps, sa)); tb = threading.Thread(target=run, args=(*b, steps, sb))
    torch.cuda.synchronize()
    t0 = time.perf_counter(); ta.start(); tb.start(); ta.join(); tb.join(); torch.cuda.synchronize(); t1 = time.perf_counter()
    print("two engines, 2 x 32  : %.3f ms per 64 clips  %7.0f clips/s" % ((t1 - t0) / steps * 1e3, 64 * steps / (t1 - t0)), flush=True)
