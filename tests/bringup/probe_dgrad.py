import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import vlb200
from vlb200 import kernels as K
torch.manual_seed(0)
def rel(a, b): return ((a.float()-b.float()).abs().max() / b.float().abs().max().clamp_min(1e-30)).item()
for (m, kx, n, use_mask, sparse) in [(4, 4096, 101, True, False), (4, 4096, 101, False, False), (128, 4096, 101, True, False),
                             (4, 256, 101, True, False), (4, 512, 101, True, False), (4, 4096, 128, True, False),
                             (4, 4096, 101, True, True), (4, 4096, 4096, True, False), (1024, 4096, 1024, True, False)]:
    nld = -(-n // 8) * 8
    dy = torch.zeros(m, nld, device="cuda", dtype=torch.bfloat16)
    if sparse:
        dy[:, 5] = 0.25; dy[:, 50] = -0.25
    else:
        dy[:, :n] = torch.randn(m, n, device="cuda").to(torch.bfloat16)
    w = torch.zeros(kx, nld, device="cuda", dtype=torch.bfloat16)
    w[:, :n] = (torch.randn(kx, n, device="cuda") * 0.05).to(torch.bfloat16)
    mask = torch.randn(m, kx, device="cuda").clamp_min(0).to(torch.bfloat16)
    dx = torch.full((m, kx), float("nan"), device="cuda", dtype=torch.bfloat16)
    K.linear_dgrad(dy, w, dx, relu_mask=mask if use_mask else None, n_contract=n)
    torch.cuda.synchronize()
    ref = dy[:, :n].float() @ w[:, :n].float().t()
    if use_mask: ref = ref * (mask.float() > 0)
    e = rel(dx, ref)
    print("m%d kx%d n%d mask%d sparse%d: rel %.3e nan=%d" % (m, kx, n, use_mask, sparse, e, torch.isnan(dx.float()).sum().item()))
    if e > 1e-2:
        bad = ((dx.float()-ref).abs() > 1e-2 * ref.abs().max()).nonzero()
        print("   bad count", bad.shape[0], "first", bad[:8].tolist(), "cols mod 256:", sorted(set((bad[:,1] % 256).tolist()))[:20], "col blocks", sorted(set((bad[:,1]//256).tolist())))
        i, j = bad[0].tolist()
        print("   got", dx[i, j].item(), "ref", ref[i, j].item())
