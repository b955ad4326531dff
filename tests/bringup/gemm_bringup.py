"""GPU bring-up of the tcgen05 contraction core (not a pytest file; run under gpurun).

Each case compares vl_gemm against a torch fp32 computation on bf16-rounded inputs on the same GPU.
This is a development probe: the graded parity tests are tests/test_gpu_*.py (oracle based).
Usage: python tests/bringup/gemm_bringup.py [group ...]
"""
import os
import sys
import time
import traceback

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import vlb200  # noqa: E402
from vlb200 import kernels as K  # noqa: E402

dev = "cuda"
BF16, F32 = torch.bfloat16, torch.float32
results = []


def rel_err(got, ref):
    got = got.float()
    ref = ref.float()
    denom = ref.abs().max().clamp_min(1e-6)
    return ((got - ref).abs().max() / denom).item()


def report(name, got, ref, tol):
    torch.cuda.synchronize()
    e = rel_err(got, ref)
    ok = e <= tol and bool(torch.isfinite(got.float()).all())
    results.append((name, ok, e))
    print("%-60s %s rel_err=%.3e (tol %.1e)" % (name, "OK  " if ok else "FAIL", e, tol), flush=True)
    return ok


def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, device=dev) * scale).to(BF16)


def group_dense():
    torch.manual_seed(0)
    for (m, n, k) in [(128, 128, 64), (128, 128, 256), (256, 256, 512), (1000, 520, 328), (1024, 4096, 1024),
                      (77, 101, 256)]:
        x = rnd(m, k)
        nld = -(-n // 8) * 8
        w_full = torch.zeros(k, nld, device=dev, dtype=BF16)
        w_full[:, :n] = rnd(k, n, scale=0.1)
        w = w_full[:, :n]
        bias = torch.randn(n, device=dev)
        ref = x.float() @ w.float() + bias
        # forward, bf16 out, relu
        out = torch.full((m, n), float("nan"), device=dev, dtype=BF16)
        K.linear_fwd(x, w_full, bias, out, relu=True, n=n)
        report("fwd  A:K  B:MN relu bf16  m%d n%d k%d" % (m, n, k), out, ref.clamp_min(0), 1.5e-2)
        out32 = torch.full((m, n), float("nan"), device=dev, dtype=F32)
        K.linear_fwd(x, w_full, bias, out32, relu=False, n=n)
        report("fwd  A:K  B:MN      f32   m%d n%d k%d" % (m, n, k), out32, ref, 2e-3)
        # dgrad: dx = dy @ w^T  (contraction over n)
        if k % 8 == 0:
            dy_full = torch.zeros(m, nld, device=dev, dtype=BF16)
            dy_full[:, :n] = rnd(m, n)
            dx = torch.full((m, k), float("nan"), device=dev, dtype=BF16)
            mask = rnd(m, k)
            K.linear_dgrad(dy_full, w_full, dx, relu_mask=mask, n_contract=n)
            refdx = (dy_full[:, :n].float() @ w.float().t()) * (mask.float() > 0)
            report("dgrad A:K  B:K  mask bf16  m%d n%d k%d" % (m, n, k), dx, refdx, 1.5e-2)
            # wgrad: dw = x^T dy
            for split in (1, 3):
                dw = torch.zeros(k, n, device=dev, dtype=F32)
                K.linear_wgrad(x, dy_full, dw, split_k=split, n=n)
                refdw = x.float().t() @ dy_full[:, :n].float()
                report("wgrad A:MN B:MN split%d f32 m%d n%d k%d" % (split, m, n, k), dw, refdw, 2e-3)


def conv_case(name, n, h, w, cin, cout, k, stride, groups, do_bwd=True):
    spec = K.ConvSpec(h, w, cin, cout, k, k, stride, groups)
    x = rnd(n, h, w, cin)
    wt = (torch.randn(k, k, cin // groups, cout, device=dev) * 0.05)
    wt_b = wt.to(BF16)
    bias = torch.randn(cout, device=dev)
    # torch reference (NCHW, OIHW), explicit TF-SAME padding
    xp = F.pad(x.float().permute(0, 3, 1, 2), (spec.pad_left, spec.pad_right, spec.pad_top, spec.pad_bottom))
    w_oihw = wt_b.float().permute(3, 2, 0, 1).contiguous()
    xp.requires_grad_(True)
    w_oihw.requires_grad_(True)
    ref = F.conv2d(xp, w_oihw, bias, stride=stride, groups=groups)
    ref_nhwc = ref.permute(0, 2, 3, 1)
    packed = K.pack_conv_weight_host(spec, wt)
    out = torch.full((n, spec.p, spec.q, cout), float("nan"), device=dev, dtype=BF16)
    K.conv_fwd(spec, x, packed, bias, out, relu=True)
    ok = report("conv fwd %s" % name, out, ref_nhwc.clamp_min(0), 1.5e-2)
    if not do_bwd:
        return
    dy = rnd(n, spec.p, spec.q, cout)
    ref.backward(dy.float().permute(0, 3, 1, 2))
    ref_dx = xp.grad[:, :, spec.pad_top:spec.pad_top + h, spec.pad_left:spec.pad_left + w].permute(0, 2, 3, 1)
    ref_dw = w_oihw.grad.permute(2, 3, 1, 0).reshape(k * k * (cin // groups), cout)
    if stride == 1:
        dx = torch.full((n, h, w, cin), float("nan"), device=dev, dtype=BF16)
        w2d = wt_b.reshape(k * k * (cin // groups), cout).contiguous()
        K.conv_dgrad(spec, dy, w2d, dx)
        report("conv dgrad %s" % name, dx, ref_dx, 1.5e-2)
    for split in (1, 4):
        dw = torch.zeros(k * k * (cin // groups), cout, device=dev, dtype=F32)
        K.conv_wgrad(spec, x, dy, dw, split_k=split)
        report("conv wgrad split%d %s" % (split, name), dw, ref_dw, 3e-3)


def group_conv_small():
    torch.manual_seed(1)
    conv_case("3x3 c64->64 13x13 n2", 2, 13, 13, 64, 64, 3, 1, 1)
    conv_case("3x3 c128->192 13x13 n3", 3, 13, 13, 128, 192, 3, 1, 1)
    conv_case("3x3 g2 c128->128 13x13 n4", 4, 13, 13, 128, 128, 3, 1, 2)


def group_conv_alexnet():
    torch.manual_seed(2)
    conv_case("conv3 256->384 n8", 8, 13, 13, 256, 384, 3, 1, 1)
    conv_case("conv4 g2 384->384 n8", 8, 13, 13, 384, 384, 3, 1, 2)
    conv_case("conv5 g2 384->256 n8", 8, 13, 13, 384, 256, 3, 1, 2)
    conv_case("conv2 g2 5x5 96->256 28x28 n4", 4, 28, 28, 96, 256, 5, 1, 2)


def group_conv_strided():
    torch.manual_seed(3)
    conv_case("5x5 s2 c64->64 27x27 n2", 2, 27, 27, 64, 64, 5, 2, 1)
    conv_case("11x11 s4 c64->96 59x59 n2", 2, 59, 59, 64, 96, 11, 4, 1)


def group_perf():
    torch.manual_seed(4)
    for (m, n, k) in [(8192, 8192, 8192), (1024, 4096, 9216), (16384, 4096, 4096)]:
        x = rnd(m, k)
        w = rnd(k, n, scale=0.02)
        out = torch.empty(m, n, device=dev, dtype=BF16)
        for bn in (256, 128):
            for _ in range(3):
                K.linear_fwd(x, w, None, out, block_n=bn)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            iters = 10
            for _ in range(iters):
                K.linear_fwd(x, w, None, out, block_n=bn)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            print("perf fwd m%d n%d k%d bn%d: %.3f ms  %.1f TFLOP/s" % (m, n, k, bn, ms, 2.0 * m * n * k / ms / 1e9),
                  flush=True)
        ref = x.float() @ w.float()
        report("perf-shape check m%d n%d k%d" % (m, n, k), out, ref, 1.5e-2)


GROUPS = {
    "dense": group_dense,
    "conv_small": group_conv_small,
    "conv_alexnet": group_conv_alexnet,
    "conv_strided": group_conv_strided,
    "perf": group_perf,
}

if __name__ == "__main__":
    names = sys.argv[1:] or list(GROUPS)
    print("device:", torch.cuda.get_device_name(0), flush=True)
    for nme in names:
        print("== group %s ==" % nme, flush=True)
        t0 = time.time()
        try:
            GROUPS[nme]()
        except Exception:
            traceback.print_exc()
            print("group %s ABORTED" % nme, flush=True)
            break
        print("== group %s done in %.1fs ==" % (nme, time.time() - t0), flush=True)
    bad = [r for r in results if not r[1]]
    print("SUMMARY: %d cases, %d failed" % (len(results), len(bad)), flush=True)
    for r in bad:
        print("  FAILED:", r[0], "rel_err=%.3e" % r[2], flush=True)
    sys.exit(1 if bad else 0)
