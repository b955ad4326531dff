"""conv1 filter gradient (space-to-depth form): one im2col box per tap (current) against the row-shift form (one tiled
box per filter row, overlapping N atoms).  Correctness against the current kernel and timing."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import vlb200
from vlb200 import _native as nv, kernels as K

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev, bf = "cuda", torch.bfloat16
torch.manual_seed(0)
spec = K.ConvSpec(59, 59, 48, 96, 3, 3, 1, 1, padding="VALID")
x = torch.randn(n, 59, 59, 48, device=dev).to(bf)
dy = torch.randn(n, 57, 57, 96, device=dev).to(bf)

def run(row, split=0):
    dw = torch.zeros(9 * 48, 96, dtype=torch.float32, device=dev)
    K.conv_wgrad_t(spec, x, dy.view(-1, 96), dw, split_k=split, row_shift=row)
    return dw

ref = run(False)
torch.cuda.synchronize()
got = run(True)
torch.cuda.synchronize()
err = ((got - ref).abs().max() / ref.abs().max()).item()
print("row_shift vs im2col: max rel diff %.3e" % err, flush=True)
small = K.ConvSpec(59, 59, 48, 96, 3, 3, 1, 1, padding="VALID")
for row in (False, True):
    for split in (0,):
        for _ in range(2):
            run(row, split)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            run(row, split)
        e1.record(); torch.cuda.synchronize()
        print("row_shift=%s split_k=%d: %.1f us (includes a %.0f KB memset)" % (row, split, e0.elapsed_time(e1) / 3 * 1e3, 9 * 48 * 96 * 4 / 1e3), flush=True)

# dy stored with a 128-channel pitch (256-byte rows): the A boxes then walk aligned rows
dyp = torch.zeros(n, 57, 57, 128, dtype=bf, device=dev)
dyp[..., :96] = dy
def run_p(row):
    dw = torch.zeros(9 * 48, 96, dtype=torch.float32, device=dev)
    K.conv_wgrad_t(spec, x, dyp.view(-1, 128), dw, row_shift=row, a_ld=128)
    return dw
got = run_p(True)
torch.cuda.synchronize()
print("row_shift + dy pitch 128 vs im2col: max rel diff %.3e" % ((got - ref).abs().max() / ref.abs().max()).item(), flush=True)
for row in (False, True):
    for _ in range(2):
        run_p(row)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        run_p(row)
    e1.record(); torch.cuda.synchronize()
    print("dy pitch 128, row_shift=%s: %.1f us" % (row, e0.elapsed_time(e1) / 3 * 1e3), flush=True)
