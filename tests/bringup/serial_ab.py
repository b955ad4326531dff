import os, sys, torch
sys.path.insert(0, "/root/repo")
import vlb200
from vlb200 import engine as E
clips = 64
cfg = E.EngineConfig(workflow="lrcn", fusion="avg", fpc=16, num_classes=101, lstm_hidden=256, clip_norm=10,
                     dropout_keep_prob=0.5, optimizer="sgd", mean=(99.197148, 105.293620, 109.503945))
eng = E.Engine(cfg, max_clips=clips)
g = torch.Generator(device="cuda").manual_seed(0)
frames = torch.randint(0, 256, (clips * 16, 227, 227, 3), dtype=torch.uint8, device="cuda", generator=g)
labels = torch.zeros(clips, 101, dtype=torch.int32, device="cuda")
labels[torch.arange(clips), torch.randint(0, 101, (clips,), device="cuda")] = 1
def timeit(iters=10, warm=3):
    for _ in range(warm): eng.train_step(frames, labels, 1e-3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): eng.train_step(frames, labels, 1e-3)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for rep in range(2):
    for serial in (False, True):
        eng.set_serial(serial)
        print("serial=%s  %.3f ms/step" % (serial, timeit()), flush=True)
