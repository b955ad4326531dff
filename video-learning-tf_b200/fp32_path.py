"""fp32-accuracy forward mode of the LRCN hot path (north-star tolerance: logits and losses <= 1e-3 relative in fp32).

The reference evaluates `sess.run(model.logits, fdict)` (run_task.py:95) in fp32 (models/alexnet/alexnet.py:60-280,
models/lstm/lstm.py:59-143, tf_util.py:4-30,44-58).  The fast device path stores activations and operands in bf16
(<= 2e-2).  This mode keeps EVERY tensor in fp32 and still runs the contractions on the tensor cores: an fp32 tensor X is
split into X_hi = bf16(X) and X_lo = bf16(X - X_hi), and X @ W = X_hi W_hi + X_lo W_hi + X_hi W_lo (+ a dropped term
<= 2^-18) is ONE launch of the tcgen05 kernel over operands concatenated along the contraction axis
([X_hi | X_lo | X_hi] x [W_hi ; W_hi ; W_lo]) with fp32 accumulation in TMEM, bias / ReLU in the epilogue and fp32 output
(csrc/fp32_path.cu builds the operands; LRN / max-pool run in fp32).  About 3x the tensor work of the bf16 forward; meant
for validation (`Validation` logits, loss read-outs), not for the training step.  No CPU fallback: everything runs in the
CUDA library.
"""
import numpy as np
import torch

from . import _native as nv
from . import kernels as K
from . import shadow_table as ST

BF16, F32 = torch.bfloat16, torch.float32


def _align(n, a=64):
    return -(-n // a) * a


class Fp32Path(object):
    """Forward pass of an `Engine`'s model in fp32 accuracy.  Operand copies are rebuilt when the engine's weights have
    changed (`engine.global_step` / `engine.weights_version`)."""

    def __init__(self, engine, chunk_clips=None):
        self.eng = engine
        cfg = engine.cfg
        # frames per pass: bounded so that the fp32 activations of conv1 (1.25 MB per frame) stay small
        self.chunk_clips = int(chunk_clips or max(1, 64 // cfg.fpc))
        self.nf = self.chunk_clips * cfg.fpc
        self._built_for = None
        self._alloc()

    # ------------------------------------------------------------------------------------------
    def _alloc(self):
        eng, cfg, sp, dev = self.eng, self.eng.cfg, self.eng.sp, self.eng.dev
        nf, b = self.nf, self.chunk_clips
        s1, s1s, s2, s3 = sp["conv1"], sp["conv1_s2d"], sp["conv2"], sp["conv3"]
        (p1h, p1w), (p2h, p2w), (p5h, p5w) = sp["pool1"], sp["pool2"], sp["pool5"]
        # the convolutions of the mode: 3x the input channels, [hi | lo | hi] per group
        self.specs = {
            "conv1": K.ConvSpec(s1s.h, s1s.w, 3 * s1s.cin, 96, s1s.kh, s1s.kw, 1, 1, padding="VALID"),
            "conv2": K.ConvSpec(s2.h, s2.w, 3 * s2.cin, s2.cout, s2.kh, s2.kw, 1, s2.groups),
            "conv3": K.ConvSpec(s3.h, s3.w, 3 * 256, 384, 3, 3, 1, 1),
            "conv4": K.ConvSpec(s3.h, s3.w, 3 * 384, 384, 3, 3, 1, 2),
            "conv5": K.ConvSpec(s3.h, s3.w, 3 * 384, 256, 3, 3, 1, 2),
        }

        def f(*shape):
            return torch.empty(*shape, dtype=F32, device=dev)

        def h(*shape):
            return torch.empty(*shape, dtype=BF16, device=dev)

        A = {}
        A["xs"] = f(nf, s1s.h, s1s.w, s1s.cin)
        A["xs3"] = h(nf, s1s.h, s1s.w, 3 * s1s.cin)
        A["a1"] = f(nf, s1.p, s1.q, 96)
        A["p1"] = f(nf, p1h, p1w, 96)
        A["p1_3"] = h(nf, p1h, p1w, 3 * 96)
        A["a2"] = f(nf, s2.p, s2.q, 256)
        A["p2"] = f(nf, p2h, p2w, 256)
        A["p2_3"] = h(nf, p2h, p2w, 3 * 256)
        A["a3"] = f(nf, s3.p, s3.q, 384)
        A["a3_3"] = h(nf, s3.p, s3.q, 3 * 384)
        A["a4"] = f(nf, s3.p, s3.q, 384)
        A["a4_3"] = h(nf, s3.p, s3.q, 3 * 384)
        A["a5"] = f(nf, s3.p, s3.q, 256)
        A["p5"] = f(nf, p5h, p5w, 256)
        A["p5_3"] = h(nf, 3 * sp["flat"])
        A["f6"] = f(nf, 4096)
        A["f6_3"] = h(nf, 3 * 4096)
        A["f7"] = f(nf, 4096)
        A["f7_3"] = h(nf, 3 * 4096)
        c, hd = cfg.num_classes, cfg.lstm_hidden
        if cfg.workflow == "lrcn":
            for layer in range(cfg.lstm_layers):
                A["gx%d" % layer] = f(nf, 4 * hd)
                A["acts%d" % layer] = f(nf, 4 * hd)
                A["cs%d" % layer] = f(nf, hd)
                A["hseq%d" % layer] = f(nf, hd)
                A["hseq_bf%d" % layer] = h(nf, hd)
                A["hprev_bf%d" % layer] = h(nf, hd)
                A["hseq_3_%d" % layer] = h(nf, 3 * hd)
            A["fused"] = f(b, hd)
            A["fused_3"] = h(b, 3 * hd)
        elif cfg.workflow == "fc" and cfg.early_fusion:
            A["pooled"] = f(b, 4096)
            A["pooled_3"] = h(b, 3 * 4096)
        else:
            A["frame_logits"] = f(nf, c)
        A["logits"] = f(b, c)
        self.A = A

        # ---- operand copies ----
        names = dict(eng.var_shapes)
        W = {}
        plan = []  # (key, shape, table, master variable)
        spx = self.specs
        plan.append(("conv1", (96, spx["conv1"].k_packed),
                     ST.split3_s2d_kmajor(s1.kh, s1.kw, 3, 96, s1.stride, spx["conv1"].cchunks * 64), "dcnn/conv1W"))
        for name in ("conv2", "conv3", "conv4", "conv5"):
            so, sx = sp[name], spx[name]
            plan.append((name, (so.cout, sx.k_packed), ST.split3_kmajor(so.taps, so.cin_g, so.cout, sx.cchunks * 64),
                         "dcnn/%sW" % name))
        total, offs = 0, []
        for key, shape, table, master in plan:
            assert table.size == int(np.prod(shape)), key
            offs.append(total)
            total += _align(table.size)
        table_all = np.full(total, -1, dtype=np.int64)
        for (key, shape, table, master), o in zip(plan, offs):
            lo = table & ST.LO_FLAG
            idx = (table & ~ST.LO_FLAG) + eng.var_off[master]
            assert idx.max() < ST.LO_FLAG
            table_all[o:o + table.size] = np.where(table >= 0, idx | lo, -1)
        self._table = torch.from_numpy(table_all.astype(np.int32)).to(dev)
        self._arena = torch.zeros(total, dtype=BF16, device=dev)
        for (key, shape, table, master), o in zip(plan, offs):
            W[key] = self._arena[o:o + table.size].view(*shape)
        # dense layers: [W_hi ; W_hi ; W_lo] rows, columns padded to a multiple of 8 (TMA row pitch)
        self._dense = []  # (key, master variable, first row, rows)
        self._dense.append(("fc6", "dcnn/fc6W", 0, sp["flat"]))
        if "dcnn/fc7W" in names:
            self._dense.append(("fc7", "dcnn/fc7W", 0, 4096))
        if "dcnn/fc8W" in names:
            self._dense.append(("fc8", "dcnn/fc8W", 0, 4096))
        for fc_name in ("output_fc", "fc_convert"):
            if fc_name + "_w" in names:
                self._dense.append((fc_name, fc_name + "_w", 0, names[fc_name + "_w"][0]))
        if cfg.workflow == "lrcn":
            for layer in range(cfg.lstm_layers):
                kn = "rnn/multi_rnn_cell/cell_%d/basic_lstm_cell/kernel" % layer
                self._dense.append(("lstm%d" % layer, kn, 0, names[kn][0] - hd))  # the x part of the kernel
        for key, master, r0, rows in self._dense:
            cols = names[master][1]
            W[key] = torch.zeros(3 * rows, _align(cols, 8), dtype=BF16, device=dev)
        self.W = W

    def refresh(self):
        """fp32 master weights -> [hi | hi | lo] bf16 operands (after load / after optimiser steps)."""
        eng = self.eng
        nv.call("vl_gather_split_bf16", eng.params, self._table, self._arena, self._arena.numel())
        names = dict(eng.var_shapes)
        for key, master, r0, rows in self._dense:
            w = eng.var(master)
            cols = names[master][1]
            nv.call("vl_split3_weight", w[r0:r0 + rows], self.W[key], rows, cols, self.W[key].shape[1])
        self._built_for = self._version()

    def _version(self):
        return (self.eng.global_step, getattr(self.eng, "weights_version", 0))

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _split(x, out, c, groups=1):
        nv.call("vl_split3_act", x, out, x.numel() // c, c, groups)
        return out

    @staticmethod
    def _dense_fwd(x3, w3, bias, out, relu, n=None):
        K.linear_fwd(x3, w3, bias, out, relu=relu, n=n)
        return out

    def _encoder(self, frames, is_u8, n, crops):
        eng, A, W, sp, spx = self.eng, self.A, self.W, self.eng.sp, self.specs
        from .engine import LRN
        s1, s1s, s2, s3 = sp["conv1"], sp["conv1_s2d"], sp["conv2"], sp["conv3"]
        nv.call("vl_frames_s2d_f32", frames, 1 if is_u8 else 0, eng._mean_dev(), A["xs"][:n], n, int(frames.shape[1]),
                int(frames.shape[2]), crops, eng.cfg.height, eng.cfg.width, s1.stride, s1.pad_top, s1.pad_left, s1s.h, s1s.w)
        self._split(A["xs"][:n], A["xs3"][:n], s1s.cin)
        K.conv_fwd(spx["conv1"], A["xs3"][:n], W["conv1"], eng.var("dcnn/conv1b"), A["a1"][:n], relu=True)
        nv.call("vl_lrn_pool_fwd_f32", A["a1"][:n], A["p1"][:n], n, s1.p, s1.q, 96, LRN["radius"], LRN["alpha"], LRN["beta"],
                LRN["bias"], 1)
        self._split(A["p1"][:n], A["p1_3"][:n], 96, s2.groups)
        K.conv_fwd(spx["conv2"], A["p1_3"][:n], W["conv2"], eng.var("dcnn/conv2b"), A["a2"][:n], relu=True)
        nv.call("vl_lrn_pool_fwd_f32", A["a2"][:n], A["p2"][:n], n, s2.p, s2.q, 256, LRN["radius"], LRN["alpha"], LRN["beta"],
                LRN["bias"], 1)
        self._split(A["p2"][:n], A["p2_3"][:n], 256, 1)
        K.conv_fwd(spx["conv3"], A["p2_3"][:n], W["conv3"], eng.var("dcnn/conv3b"), A["a3"][:n], relu=True)
        self._split(A["a3"][:n], A["a3_3"][:n], 384, 2)
        K.conv_fwd(spx["conv4"], A["a3_3"][:n], W["conv4"], eng.var("dcnn/conv4b"), A["a4"][:n], relu=True)
        self._split(A["a4"][:n], A["a4_3"][:n], 384, 2)
        K.conv_fwd(spx["conv5"], A["a4_3"][:n], W["conv5"], eng.var("dcnn/conv5b"), A["a5"][:n], relu=True)
        nv.call("vl_lrn_pool_fwd_f32", A["a5"][:n], A["p5"][:n], n, s3.p, s3.q, 256, 0, 0.0, 0.0, 1.0, 0)
        flat = sp["flat"]
        self._split(A["p5"][:n].view(n, flat), A["p5_3"][:n], flat)  # HWC-major flatten (alexnet.py:228)
        self._dense_fwd(A["p5_3"][:n], W["fc6"], eng.var("dcnn/fc6b"), A["f6"][:n], True)
        feat, feat3 = A["f6"][:n], A["f6_3"][:n]
        if "fc7" in W:
            self._split(A["f6"][:n], A["f6_3"][:n], 4096)
            self._dense_fwd(A["f6_3"][:n], W["fc7"], eng.var("dcnn/fc7b"), A["f7"][:n], True)
            feat, feat3 = A["f7"][:n], A["f7_3"][:n]
        return feat, feat3

    def _head(self, feat, feat3, n):
        from .engine import POOL
        eng, cfg, A, W = self.eng, self.eng.cfg, self.A, self.W
        b, c = n // cfg.fpc, cfg.num_classes
        logits = A["logits"][:b]
        dim = feat.shape[1]
        if cfg.workflow == "fc" and cfg.early_fusion:
            nv.call("vl_segment_pool_fwd", feat, None, cfg.fpc, b, dim, POOL[cfg.fusion], A["pooled"][:b], None)
            self._split(A["pooled"][:b], A["pooled_3"][:b], dim)
            return self._dense_fwd(A["pooled_3"][:b], W["fc_convert"], eng.var("fc_convert_b"), logits, False, n=c)
        if cfg.workflow in ("singleframe", "fc"):
            key, bname = eng._frame_classifier()
            self._split(feat, feat3, dim)
            fl = A["frame_logits"][:n]
            self._dense_fwd(feat3, W[key], eng.var(bname), fl, False, n=c)
            nv.call("vl_segment_pool_fwd", fl, None, cfg.fpc, b, c, POOL[cfg.fusion], logits, None)
            return logits
        hd, t_len = cfg.lstm_hidden, cfg.fpc
        x, x3 = feat, feat3
        for layer in range(cfg.lstm_layers):
            kern = eng.var("rnn/multi_rnn_cell/cell_%d/basic_lstm_cell/kernel" % layer)
            bias = eng.var("rnn/multi_rnn_cell/cell_%d/basic_lstm_cell/bias" % layer)
            d_in = kern.shape[0] - hd
            self._split(x, x3, d_in)
            gx = A["gx%d" % layer][:n]
            self._dense_fwd(x3, W["lstm%d" % layer], bias, gx, False)
            nv.call("vl_lstm_fwd_cluster" if hd == 256 else "vl_lstm_fwd", gx, kern[d_in:], A["acts%d" % layer][:n],
                    A["cs%d" % layer][:n], A["hseq%d" % layer][:n], A["hseq_bf%d" % layer][:n],
                    A["hprev_bf%d" % layer][:n], b, t_len, hd, 1.0)
            x, x3 = A["hseq%d" % layer][:n], A["hseq_3_%d" % layer][:n]
        state = cfg.fusion == "state"
        nv.call("vl_segment_pool_fwd", x, None, t_len, b, hd, POOL["last" if state else cfg.fusion], A["fused"][:b], None)
        fc = "fc_convert" if state else "output_fc"
        if fc in W:
            self._split(A["fused"][:b], A["fused_3"][:b], hd)
            self._dense_fwd(A["fused_3"][:b], W[fc], eng.var(fc + "_b"), logits, False, n=c)
        else:
            logits.copy_(A["fused"][:b])
        return logits

    def forward_device(self, frames, crops=None):
        """fp32 device logits [clips, C] of `frames` (what Engine.forward_device accepts), evaluated chunk by chunk."""
        eng, cfg = self.eng, self.eng.cfg
        if self._built_for != self._version():
            self.refresh()
        eng._frames_ready = None
        frames, is_u8, n = eng._stage_frames(frames)
        if n % cfg.fpc != 0:
            raise ValueError("number of frames (%d) is not a multiple of num_frames_per_clip (%d)" % (n, cfg.fpc))
        crops = eng._stage_crops(crops, frames, n)
        out = torch.empty(n // cfg.fpc, cfg.num_classes, dtype=F32, device=eng.dev)
        for f0 in range(0, n, self.nf):
            f1 = min(n, f0 + self.nf)
            feat, feat3 = self._encoder(frames[f0:f1], is_u8, f1 - f0, None if crops is None else crops[f0:f1])
            logits = self._head(feat, feat3, f1 - f0)
            out[f0 // cfg.fpc:f1 // cfg.fpc].copy_(logits)
        return out
