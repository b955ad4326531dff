"""BASELINE.json configs[4]: LRCN forward-only inference sweep, 16..1024 clips of 16 frames on one GPU (device-resident
uint8 frames; logits stay on the device).  Prints clips/s per batch size."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import vlb200  # noqa
from vlb200 import engine as E

sizes = [int(a) for a in sys.argv[1:]] or [16, 32, 64, 128, 256, 512, 1024]
cfg = E.EngineConfig(workflow="lrcn", fusion="avg", fpc=16, num_classes=101, lstm_hidden=256, mean=(99.197148, 105.293620, 109.503945))
eng = E.Engine(cfg, max_clips=max(sizes))
g = torch.Generator(device="cuda").manual_seed(0)
print("%8s %10s %12s %10s" % ("clips", "ms", "clips/s", "TFLOP/s"))
for clips in sizes:
    frames = torch.randint(0, 256, (clips * 16, 227, 227, 3), dtype=torch.uint8, device="cuda", generator=g)
    for _ in range(3):
        eng.forward_device(frames)
    torch.cuda.synchronize()
    iters = 10 if clips <= 256 else 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        eng.forward_device(frames)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print("%8d %10.3f %12.1f %10.1f" % (clips, ms, clips / ms * 1e3, clips / ms * 1e3 * 23.983e9 / 1e12), flush=True)
    del frames
print("max mem GB %.1f" % (torch.cuda.max_memory_allocated() / 1e9))
