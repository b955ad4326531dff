"""conv1 filter gradient (space-to-depth form, swapped operands) against the pitch of its operands: x_s2d with 48 or 64
channel pitch (96 B vs 128 B pixel rows), dy with 96 or 128 channel pitch (192 B vs 256 B rows)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import vlb200
from vlb200 import _native as nv, kernels as K

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev, bf = "cuda", torch.bfloat16
torch.manual_seed(0)

def timed(name, fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        fn()
    e1.record(); torch.cuda.synchronize()
    print("%-44s %8.1f us" % (name, e0.elapsed_time(e1) / 3 * 1e3), flush=True)

x48 = torch.randn(n, 59, 59, 48, device=dev).to(bf)
dy96 = torch.randn(n, 57, 57, 96, device=dev).to(bf)
ref = None
for (xp, dp) in ((48, 96), (64, 96), (48, 128), (64, 128)):
    xbuf = torch.zeros(n, 59, 59, xp, dtype=bf, device=dev); xbuf[..., :48] = x48
    dbuf = torch.zeros(n, 57, 57, dp, dtype=bf, device=dev); dbuf[..., :96] = dy96
    spec = K.ConvSpec(59, 59, xp, 96, 3, 3, 1, 1, padding="VALID")
    dw = torch.zeros(9 * xp, 96, dtype=torch.float32, device=dev)
    def run():
        K.conv_wgrad_t(spec, xbuf, dbuf.view(-1, dp), dw, a_ld=dp)
    timed("wgrad_t x pitch %d, dy pitch %d" % (xp, dp), run)
    dw.zero_(); run(); torch.cuda.synchronize()
    got = dw.view(9, xp, 96)[:, :48].clone()
    if ref is None:
        ref = got
    else:
        print("    max rel diff vs 48/96: %.2e" % ((got - ref).abs().max() / ref.abs().max()).item())
