"""Board power and SM clock of the step's main kernels, each run alone in a loop for ~1 s (NVML samples every 20 ms):
energy per launch = average power x launch time.  The train step runs under sw_power_cap, so its duration follows the
ENERGY of its kernels rather than the sum of their isolated durations."""
import os, sys, threading, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import vlb200  # noqa
from vlb200 import _native as nv, kernels as K
import pynvml

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
n = 1024
dev, bf = "cuda", torch.bfloat16
torch.manual_seed(0)
LRN = (2, 2e-05, 0.75, 1.0)


class Sampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.stop = False
        self.p, self.c = [], []

    def run(self):
        while not self.stop:
            self.p.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1e3)
            self.c.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
            time.sleep(0.02)


def measure(name, fn, flops=0.0, seconds=1.0):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); fn(); fn(); e1.record(); torch.cuda.synchronize()
    us0 = e0.elapsed_time(e1) / 3 * 1e3
    iters = max(10, int(seconds * 1e6 / us0))
    s = Sampler(); s.start()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    s.stop = True; s.join()
    us = e0.elapsed_time(e1) / iters * 1e3
    k = len(s.p) // 3  # skip the ramp
    p = sum(s.p[k:]) / max(1, len(s.p[k:])); c = sorted(s.c[k:])[len(s.c[k:]) // 2] if s.c[k:] else 0
    print("%-30s first launches %7.1f us | sustained %7.1f us  %6.0f W  %4d MHz  %7.1f mJ per launch%s" % (
        name, us0, us, p, c, p * us * 1e-3, ("  %6.0f TFLOP/s" % (flops / us / 1e6)) if flops else ""), flush=True)
    time.sleep(0.5)
    return p * us * 1e-3


cases = []
s1s = K.ConvSpec(59, 59, 48, 96, 3, 3, 1, 1, padding="VALID")
f1 = 2.0 * n * 57 * 57 * 96 * 363
x1 = torch.randn(n, 59, 59, 48, device=dev).to(bf)
w1 = (torch.randn(96, s1s.k_packed, device=dev) * 0.05).to(bf)
b96 = torch.zeros(96, device=dev)
a1 = torch.empty(n, 57, 57, 96, device=dev, dtype=bf)
dy1 = torch.randn(n, 57, 57, 96, device=dev).to(bf)
dw1 = torch.zeros(9 * 48, 96, dtype=torch.float32, device=dev)
cases.append(("conv1 fwd (flat)", lambda: K.conv_fwd_flat(s1s, x1, w1, b96, a1, relu=True, flops=f1), f1))
cases.append(("conv1 wgrad (row-shift)", lambda: K.conv_wgrad_t(s1s, x1, dy1.view(-1, 96), dw1, row_shift=True), f1))
s2 = K.ConvSpec(28, 28, 96, 256, 5, 5, 1, 2)
f2 = K.conv_flops(s2, n)
x2 = torch.randn(n, 28, 28, 96, device=dev).to(bf)
w2k = (torch.randn(256, s2.k_packed, device=dev) * 0.05).to(bf)
b256 = torch.zeros(256, device=dev)
a2 = torch.empty(n, 28, 28, 256, device=dev, dtype=bf)
dy2 = torch.randn(n, 28, 28, 256, device=dev).to(bf)
dw2 = torch.zeros(25 * 48, 256, dtype=torch.float32, device=dev)
w = (torch.randn(5, 5, 48, 256, device=dev) * 0.05)
rows, cols = K.d2s_filter_shape(s2, 2, 2)
wd = torch.empty(rows, cols, dtype=bf, device=dev)
nv.call("vl_pack_dgrad_d2s", w, wd, 5, 5, 48, 128, 2, 2, 2)
dx2 = torch.empty(n, 28, 28, 96, dtype=bf, device=dev)
cases.append(("conv2 fwd (flat)", lambda: K.conv_fwd_flat(s2, x2, w2k, b256, a2, relu=True), f2))
cases.append(("conv2 dgrad (d2s)", lambda: K.conv_dgrad_d2s(s2, dy2, wd, dx2, sh=2, sw=2), f2))
cases.append(("conv2 wgrad (swapped)", lambda: K.conv_wgrad_t(s2, x2, dy2.view(-1, 256), dw2), f2))
s3 = K.ConvSpec(13, 13, 256, 384, 3, 3, 1, 1)
f3 = K.conv_flops(s3, n)
x3 = torch.randn(n, 13, 13, 256, device=dev).to(bf)
w3k = K.pack_conv_weight_host(s3, torch.randn(3, 3, 256, 384, device=dev) * 0.05)
w3d = (torch.randn(9 * 256, 384, device=dev) * 0.05).to(bf)
b384 = torch.zeros(384, device=dev)
a3 = torch.empty(n, 13, 13, 384, device=dev, dtype=bf)
dy3 = torch.randn(n, 13, 13, 384, device=dev).to(bf)
dx3 = torch.empty(n, 13, 13, 256, device=dev, dtype=bf)
dw3 = torch.zeros(9 * 256, 384, dtype=torch.float32, device=dev)
cases.append(("conv3 fwd (im2col)", lambda: K.conv_fwd(s3, x3, w3k, b384, a3), f3))
cases.append(("conv3 dgrad (im2col)", lambda: K.conv_dgrad(s3, dy3, w3d, dx3), f3))
cases.append(("conv3 wgrad", lambda: K.conv_wgrad(s3, x3, dy3, dw3), f3))
xf = torch.randn(n, 9216, device=dev).to(bf)
wf = (torch.randn(9216, 4096, device=dev) * 0.02).to(bf)
b4096 = torch.zeros(4096, device=dev)
of = torch.empty(n, 4096, dtype=bf, device=dev)
cases.append(("fc6 fwd", lambda: K.linear_fwd(xf, wf, b4096, of, relu=True), 2.0 * n * 9216 * 4096))
p1 = torch.empty(n, 28, 28, 96, device=dev, dtype=bf); arg1 = torch.empty(n, 28, 28, 96, device=dev, dtype=torch.uint8)
xa1 = (torch.randn(n, 57, 57, 96, device=dev) * 60).clamp_(min=0).to(bf)
dp1 = torch.randn(n, 28, 28, 96, device=dev).to(bf); dxa1 = torch.empty_like(xa1); db = torch.zeros(96, device=dev)
cases.append(("lrn1 + pool fwd", lambda: nv.call("vl_lrn_pool_fwd", xa1, p1, arg1, n, 57, 57, 96, *LRN), 0))
cases.append(("lrn1 + pool bwd", lambda: nv.call("vl_pool_lrn_bwd", xa1, dp1, arg1, dxa1, db, n, 57, 57, 96, *LRN), 0))
big = torch.empty(61_351_653 + 100, device=dev); g = torch.randn_like(big); sc = torch.ones(8, device=dev)
cases.append(("sgd update 61 M", lambda: nv.call("vl_sgd_update", big, g, 61_351_616, 1e-6, sc, 1.0), 0))
tot = 0.0
for name, fn, fl in cases:
    tot += measure(name, fn, fl)
print("idle: %.0f W" % (pynvml.nvmlDeviceGetPowerUsage(h) / 1e3))
