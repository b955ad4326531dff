"""Development probe: time the config-2 train step (64 clips x 16 frames) and the forward pass on one GPU."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import vlb200  # noqa
from vlb200 import engine as E, _native as nv

clips = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = E.EngineConfig(workflow="lrcn", fusion="avg", fpc=16, num_classes=101, lstm_hidden=256, clip_norm=10,
                     dropout_keep_prob=0.5, optimizer="sgd", mean=(99.197148, 105.293620, 109.503945))
t0 = time.time()
eng = E.Engine(cfg, max_clips=clips)
print("engine init %.1fs" % (time.time() - t0), flush=True)
g = torch.Generator(device="cuda").manual_seed(0)
frames = torch.randint(0, 256, (clips * 16, 227, 227, 3), dtype=torch.uint8, device="cuda", generator=g)
labels = torch.zeros(clips, 101, dtype=torch.int32, device="cuda")
labels[torch.arange(clips), torch.randint(0, 101, (clips,), device="cuda")] = 1


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


ms = timeit(lambda: eng.forward_device(frames))
print("forward   : %.3f ms  %.1f clips/s  (%.1f%% of sustained tensor peak)" % (
    ms, clips / ms * 1e3, clips / ms * 1e3 * 23.983e9 / 1371.0e12 * 100), flush=True)
l0 = nv.lib().vl_launch_count()
ms = timeit(lambda: eng.train_step(frames, labels, 1e-3))
l1 = nv.lib().vl_launch_count()
print("train step: %.3f ms  %.1f clips/s  (%.1f%% of sustained tensor peak), launches/step %d" % (
    ms, clips / ms * 1e3, clips / ms * 1e3 * 68.326e9 / 1371.0e12 * 100, (l1 - l0) // 7), flush=True)
print("loss etc:", eng.train_step(frames, labels, 1e-3))
print("max mem GB", torch.cuda.max_memory_allocated() / 1e9)

if "--serial" in sys.argv:
    eng.set_serial(True)  # every kernel on one stream: durations are those of the kernels alone
if "--profile" in sys.argv:
    from torch.profiler import profile, ProfilerActivity
    import re
    for what, fn in (("forward", lambda: eng.forward_device(frames)), ("train", lambda: eng.train_step(frames, labels, 1e-3))):
        fn(); torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            fn(); torch.cuda.synchronize()
        evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
        evs.sort(key=lambda e: e.time_range.start)
        tot = sum(e.device_time for e in evs)
        print("== %s: %d kernels, %.1f us device time ==" % (what, len(evs), tot))
        t0 = evs[0].time_range.start if evs else 0
        for i, e in enumerate(evs):
            name = re.sub(r"<unnamed>::|\(anonymous namespace\)::|void ", "", e.name)
            name = re.sub(r"\(.*", "", name)[:44]
            print("%3d %-44s start %9.1f  dur %8.1f us" % (i, name, e.time_range.start - t0, e.device_time))
