"""conv1 forward (tap-shifted kernel) with the space-to-depth input stored at 48 or 64 channels per pixel (96 B vs 128 B
rows): same 48 contraction channels (3 of 4 K-steps per tap), only the TMA box rows change."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import vlb200
from vlb200 import _native as nv, kernels as K

n, dev, bf = 1024, "cuda", torch.bfloat16
torch.manual_seed(0)
spec = K.ConvSpec(59, 59, 48, 96, 3, 3, 1, 1, padding="VALID")
x48 = torch.randn(n, 59, 59, 48, device=dev).to(bf)
wp = K.pack_conv_weight_host(spec, torch.randn(3, 3, 48, 96, device=dev) * 0.05)
b = torch.full((96,), 0.1, device=dev)
outs = {}
for pitch in (48, 64):
    x = torch.zeros(n, 59, 59, pitch, dtype=bf, device=dev)
    x[..., :48] = x48
    y = torch.empty(n, 57, 57, 96, dtype=bf, device=dev)
    d = nv.ConvFlatDesc()
    d.n, d.h, d.w, d.c = n, 59, 59, pitch
    d.kh, d.kw = 3, 3
    d.pad_top = d.pad_left = d.pad_bottom = d.pad_right = 0
    d.groups, d.cin_g, d.cout_g = 1, 48, 96
    d.flip_taps = 0
    d.w_rows, d.w_ld = 96, spec.k_packed
    d.c_ld = 96
    d.relu = 1
    def run():
        nv.conv_flat(d, x, wp, b, y)
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        run()
    e1.record(); torch.cuda.synchronize()
    print("conv1 fwd, x pitch %d: %.1f us" % (pitch, e0.elapsed_time(e1) / 3 * 1e3), flush=True)
    outs[pitch] = y.clone()
print("identical outputs:", torch.equal(outs[48], outs[64]))
