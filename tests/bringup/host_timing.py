"""Where does the wall time of a train step go on the host?  Enqueue time (Python + ctypes + torch stream/event calls)
against the device time of the step, and the GPU idle gap at the start of a step (the scalar read-back of the previous
step is a full synchronisation)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import vlb200  # noqa
from vlb200 import engine as E

clips = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = E.EngineConfig(workflow="lrcn", fusion="avg", fpc=16, num_classes=101, lstm_hidden=256, clip_norm=10,
                     dropout_keep_prob=0.5, optimizer="sgd", mean=(99.197148, 105.293620, 109.503945))
eng = E.Engine(cfg, max_clips=clips)
g = torch.Generator(device="cuda").manual_seed(0)
frames = torch.randint(0, 256, (clips * 16, 227, 227, 3), dtype=torch.uint8, device="cuda", generator=g)
labels = torch.zeros(clips, 101, dtype=torch.int32, device="cuda")
labels[torch.arange(clips), torch.randint(0, 101, (clips,), device="cuda")] = 1

marks = {}
orig_read = eng.read_step_scalars
def read(lr):
    marks["enq_done"] = time.perf_counter()
    return orig_read(lr)
eng.read_step_scalars = read
for _ in range(3):
    eng.train_step(frames, labels, 1e-3)
torch.cuda.synchronize()
enq, tot = [], []
for _ in range(10):
    t0 = time.perf_counter()
    eng.train_step(frames, labels, 1e-3)
    t1 = time.perf_counter()
    enq.append(marks["enq_done"] - t0)
    tot.append(t1 - t0)
print("host enqueue per step: %.3f ms (min %.3f)   wall per step: %.3f ms" % (
    1e3 * sum(enq) / len(enq), 1e3 * min(enq), 1e3 * sum(tot) / len(tot)))
# the same step without the per-step read-back: the host may run ahead
eng.read_step_scalars = lambda lr: None
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    eng.train_step(frames, labels, 1e-3)
t_enq = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print("10 steps without read-back: enqueue %.3f ms/step, wall %.3f ms/step" % (t_enq * 100, t_all * 100))
