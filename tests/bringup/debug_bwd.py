"""Development probe: compare every intermediate of the single-frame train step with the oracle."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import vlb200
from vlb200 import engine as E
from oracle import lrcn_numpy as O

def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)

clips, fpc = 2, 2
cfg = E.EngineConfig(workflow="singleframe", fusion="avg", fpc=fpc, num_classes=101, clip_norm=10)
params = E.init_variables(cfg, seed=21)
rng = np.random.default_rng(22)
amp = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
frames = (rng.uniform(-1, 1, size=(clips * fpc, 227, 227, 3)) * amp).astype(np.float32)
labels = rng.integers(0, 101, clips)
onehot = np.zeros((clips, 101), np.int32); onehot[np.arange(clips), labels] = 1
eng = E.Engine(cfg, max_clips=clips, params=params)
eng.train_step(frames, onehot, 1e-3, apply_update=False)
A, G = eng.A, eng.G
n = clips * fpc
q = O.bf16_round
logits, cache = O.singleframe_forward(params, frames, fpc, "avg", True, q)
c = cache["ac"]
def f(t): return t.float().cpu().numpy()
for name in ("a1", "n1", "p1", "a2", "n2", "p2", "a3", "a4", "a5", "p5", "f6", "f7"):
    print("fwd %-4s rel %.3e  absmax %.3e" % (name, rel(f(A[name][:n]), c[name]), np.abs(c[name]).max()))
fl = (c["f7"] @ q(params["dcnn/fc8W"]) + params["dcnn/fc8b"])
print("fwd frame_logits rel %.3e absmax %.3e" % (rel(f(A["frame_logits"][:n]), fl), np.abs(fl).max()))
print("fwd logits rel %.3e" % rel(f(A["logits"][:clips]), logits))
loss, dlog, _ = O.softmax_ce(logits, onehot)
print("dlogits rel %.3e" % rel(f(A["dlogits"][:clips, :101]), dlog))
dfl = O.temporal_fusion_backward((clips, fpc, 101), "avg", dlog).reshape(n, 101)
print("d_frame_logits rel %.3e" % rel(f(A["d_frame_logits_bf16"][:n, :101]), dfl))
P = params
dfl = q(dfl)
df7 = q((dfl @ q(P["dcnn/fc8W"]).T) * (c["f7"] > 0))
print("df7 rel %.3e absmax %.3e" % (rel(f(G["df7"][:n]), df7), np.abs(df7).max()))
got = f(G["df7"][:n])
print("  nonzero got %d ref %d ; ref==0&got!=0: %d ; ref!=0&got==0: %d" % ((got != 0).sum(), (df7 != 0).sum(), ((df7 == 0) & (got != 0)).sum(), ((df7 != 0) & (got == 0)).sum()))
nz = df7 != 0
ratio = got[nz] / df7[nz]
print("  ratio pctiles", np.percentile(ratio, [0, 1, 25, 50, 75, 99, 100]))
unm = dfl @ P["dcnn/fc8W"].T
print("  vs unmasked ref rel %.3e" % rel(got, unm))
print("  rows:", [rel(got[i], df7[i]) for i in range(n)])
print("  got row0[:8]", got[0, :8], "ref", df7[0, :8], "f7", c["f7"][0, :8])
dflg = f(A["d_frame_logits_bf16"][:n])
print("  dfl got nz cols row0", np.nonzero(dflg[0])[0], dflg[0][np.nonzero(dflg[0])[0]], "ref", np.nonzero(dfl[0])[0], dfl[0][np.nonzero(dfl[0])[0]])

df6 = q((df7 @ q(P["dcnn/fc7W"]).T) * (c["f6"] > 0))
print("df6 rel %.3e" % rel(f(G["df6"][:n]), df6))
dp5 = q(df6 @ q(P["dcnn/fc6W"]).T).reshape(c["p5"].shape)
print("dp5 rel %.3e" % rel(f(G["dp5"][:n]), dp5))
da5 = q(O.maxpool_3x3s2_backward(c["a5"].shape, c["arg5"], dp5) * (c["a5"] > 0))
print("da5 rel %.3e" % rel(f(G["da5"][:n]), da5))
da4, _, _ = O.conv2d_same_backward(c["a4"], q(P["dcnn/conv5W"]), da5, 1, 2); da4 = q(da4 * (c["a4"] > 0))
print("da4 rel %.3e" % rel(f(G["da4"][:n]), da4))
da3, _, _ = O.conv2d_same_backward(c["a3"], q(P["dcnn/conv4W"]), da4, 1, 2); da3 = q(da3 * (c["a3"] > 0))
print("da3 rel %.3e" % rel(f(G["da3"][:n]), da3))
dp2, _, _ = O.conv2d_same_backward(c["p2"], q(P["dcnn/conv3W"]), da3, 1, 1); dp2 = q(dp2)
print("dp2 rel %.3e" % rel(f(G["dp2"][:n]), dp2))
dn2 = q(O.maxpool_3x3s2_backward(c["n2"].shape, c["arg2"], dp2))
print("dn2 rel %.3e" % rel(f(G["dn2"][:n]), dn2))
da2 = q(O.lrn_backward(c["a2"], dn2) * (c["a2"] > 0))
print("da2 rel %.3e" % rel(f(G["da2"][:n]), da2))
dp1, _, _ = O.conv2d_same_backward(c["p1"], q(P["dcnn/conv2W"]), da2, 1, 2); dp1 = q(dp1)
print("dp1 rel %.3e" % rel(f(G["dp1"][:n]), dp1))
dn1 = q(O.maxpool_3x3s2_backward(c["n1"].shape, c["arg1"], dp1))
print("dn1 rel %.3e" % rel(f(G["dn1"][:n]), dn1))
da1 = q(O.lrn_backward(c["a1"], dn1) * (c["a1"] > 0))
print("da1 rel %.3e" % rel(f(G["da1"][:n]), da1))
