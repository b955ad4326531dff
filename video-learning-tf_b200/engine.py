"""Engine: the device-side replacement of the two `sess.run` calls of the reference.

  run_task.py:44   sess.run([summaries, loss, lr, global_step, optimizer], feed_dict)  -> Engine.train_step
  run_task.py:95   sess.run(model.logits, feed_dict)                                   -> Engine.forward

The engine owns one flat fp32 parameter arena (variables keyed by the reference's TF names), the
bf16 operand copies the tensor-core kernels read, all activation buffers, and the launch sequence of the
hand-written kernels (csrc/) through the C-ABI.  PyTorch provides device memory, streams and (for W > 1) the
NCCL process group; it performs no arithmetic of the model.
"""
import math
import os

import numpy as np
import torch

from . import _native as nv
from . import kernels as K
from . import parallel

BF16, F32 = torch.bfloat16, torch.float32
POOL = {"avg": 0, "last": 1, "max": 2}
LRN = dict(radius=2, alpha=2e-05, beta=0.75, bias=1.0)  # alexnet.py:80-89,121-130


def _align(n, a=64):
    return -(-n // a) * a


class EngineConfig(object):
    """The subset of `Settings` (settings_.py) the device path depends on."""

    """workflow:
      lrcn         dcnn(fc6|fc7) -> LSTM -> temporal fusion (avg | last | state) -> dropout -> output fc  (lstm.py:59-99)
      singleframe  dcnn(fc8 logits per frame) -> late fusion over fpc                                    (model.py:149-151)
      fc           dcnn(fc6|fc7) -> [early fusion over fpc] -> convert_dim_fc "fc_convert" -> [late fusion]
                   (classifier fc, model.py:103-117,149-151); early_fusion selects which of the two fusions applies
    fusion `state` (lstm.py:81,91-93; model.py:137-143): the logits come from the final hidden state h of the top
    layer through convert_dim_fc under its default name (fc_convert), without temporal fusion, dropout or output_fc."""

    def __init__(self, num_classes=101, fpc=16, workflow="lrcn", frame_encoding_layer="fc7", lstm_hidden=256,
                 lstm_layers=1, fusion="avg", optimizer="sgd", clip_norm=None, dropout_keep_prob=0.0,
                 height=227, width=227, mean=None, seed=1234, early_fusion=False):
        assert workflow in ("lrcn", "singleframe", "fc")
        self.num_classes = int(num_classes)
        self.fpc = int(fpc)
        self.workflow = workflow
        self.early_fusion = bool(early_fusion) and workflow == "fc"
        self.frame_encoding_layer = frame_encoding_layer if workflow in ("lrcn", "fc") else "fc8"
        if workflow == "fc" and self.frame_encoding_layer not in ("fc6", "fc7"):
            raise ValueError("the fc workflow classifies fc6 / fc7 features (fc8 logits: workflow singleframe)")
        self.lstm_hidden = int(lstm_hidden)
        self.lstm_layers = int(lstm_layers)
        self.fusion = fusion
        self.optimizer = optimizer
        self.clip_norm = clip_norm
        self.dropout_keep_prob = float(dropout_keep_prob or 0.0)
        self.height, self.width = int(height), int(width)
        self.mean = tuple(mean) if mean is not None else None
        self.seed = int(seed)
        if fusion not in ("avg", "last") and not (fusion == "state" and workflow == "lrcn"):
            # `reshape` (tf_util.py:24-27) leaves one row per FRAME, which no label layout of the feeder matches
            # (dataset_.py:400-408: one label row per clip); anything else is undefined in the reference too
            raise ValueError("Undefined frame fusion type : %s" % fusion)
        if optimizer not in ("sgd", "adam"):
            raise ValueError("Undefined optimizer %s" % optimizer)


def variable_shapes(cfg):
    """(name, shape) of every trainable variable in the reference's creation order."""
    sp = encoder_specs(cfg.height, cfg.width)
    out = []
    for name in ("conv1", "conv2", "conv3", "conv4", "conv5"):
        s = sp[name]
        out.append(("dcnn/%sW" % name, (s.kh, s.kw, s.cin_g, s.cout)))
        out.append(("dcnn/%sb" % name, (s.cout,)))
    out.append(("dcnn/fc6W", (sp["flat"], 4096)))
    out.append(("dcnn/fc6b", (4096,)))
    feat = 4096
    fl = cfg.frame_encoding_layer
    if fl != "fc6":
        out.append(("dcnn/fc7W", (4096, 4096)))
        out.append(("dcnn/fc7b", (4096,)))
        if fl != "fc7":
            out.append(("dcnn/fc8W", (4096, cfg.num_classes)))
            out.append(("dcnn/fc8b", (cfg.num_classes,)))
            feat = cfg.num_classes
    if cfg.workflow == "lrcn":
        d_in = feat
        for layer in range(cfg.lstm_layers):
            out.append(("rnn/multi_rnn_cell/cell_%d/basic_lstm_cell/kernel" % layer,
                        (d_in + cfg.lstm_hidden, 4 * cfg.lstm_hidden)))
            out.append(("rnn/multi_rnn_cell/cell_%d/basic_lstm_cell/bias" % layer, (4 * cfg.lstm_hidden,)))
            d_in = cfg.lstm_hidden
        if cfg.lstm_hidden != cfg.num_classes:
            # tf.get_variable ignores name scopes: "output_fc_{w,b}" (lstm.py:90), or convert_dim_fc's default name
            # "fc_convert_{w,b}" when the logits come from the LSTM state (model.py:137-143)
            fc = "fc_convert" if cfg.fusion == "state" else "output_fc"
            out.append((fc + "_w", (cfg.lstm_hidden, cfg.num_classes)))
            out.append((fc + "_b", (cfg.num_classes,)))
    elif cfg.workflow == "fc" and feat != cfg.num_classes:
        out.append(("fc_convert_w", (feat, cfg.num_classes)))  # model.py:115-117 convert_dim_fc(feature_vectors, C)
        out.append(("fc_convert_b", (cfg.num_classes,)))
    return out


def encoder_specs(h, w):
    """Layer geometry of dcnn.create (alexnet.py:60-211) for an h x w input."""
    sp = {}
    sp["conv1"] = K.ConvSpec(h, w, 3, 96, 11, 11, 4, 1)
    # conv1 runs as a 3x3 stride-1 VALID convolution over the space-to-depth input (4x4x3 = 48 channels)
    c1 = sp["conv1"]
    st = c1.stride
    hb, wb = c1.p - 1 + -(-c1.kh // st), c1.q - 1 + -(-c1.kw // st)
    sp["conv1_s2d"] = K.ConvSpec(hb, wb, st * st * 3, 96, -(-c1.kh // st), -(-c1.kw // st), 1, 1, padding="VALID")
    p1h, p1w = K.valid_out(sp["conv1"].p, 3, 2), K.valid_out(sp["conv1"].q, 3, 2)
    sp["conv2"] = K.ConvSpec(p1h, p1w, 96, 256, 5, 5, 1, 2)
    p2h, p2w = K.valid_out(sp["conv2"].p, 3, 2), K.valid_out(sp["conv2"].q, 3, 2)
    sp["conv3"] = K.ConvSpec(p2h, p2w, 256, 384, 3, 3, 1, 1)
    sp["conv4"] = K.ConvSpec(p2h, p2w, 384, 384, 3, 3, 1, 2)
    sp["conv5"] = K.ConvSpec(p2h, p2w, 384, 256, 3, 3, 1, 2)
    p5h, p5w = K.valid_out(p2h, 3, 2), K.valid_out(p2w, 3, 2)
    sp["pool1"] = (p1h, p1w)
    sp["pool2"] = (p2h, p2w)
    sp["pool5"] = (p5h, p5w)
    sp["flat"] = p5h * p5w * 256
    return sp


def init_variables(cfg, seed=None):
    """Random init following make_w_b (alexnet.py:40-46), convert_dim_fc (tf_util.py:44-45) and the
    BasicLSTMCell defaults (glorot-uniform kernel, zero bias); numpy Generator(seed) in creation order."""
    rng = np.random.default_rng(cfg.seed if seed is None else seed)

    def trunc_normal(shape, std=0.05):
        out = rng.standard_normal(size=shape)
        bad = np.abs(out) > 2.0
        while bad.any():
            out[bad] = rng.standard_normal(size=int(bad.sum()))
            bad = np.abs(out) > 2.0
        return (out * std).astype(np.float32)

    params = {}
    for name, shape in variable_shapes(cfg):
        if name.endswith("basic_lstm_cell/kernel"):
            lim = math.sqrt(6.0 / (shape[0] + shape[1]))
            params[name] = rng.uniform(-lim, lim, size=shape).astype(np.float32)
        elif name.endswith("basic_lstm_cell/bias"):
            params[name] = np.zeros(shape, np.float32)
        elif len(shape) == 1:
            params[name] = np.full(shape, 0.1, np.float32)
        else:
            params[name] = trunc_normal(shape)
    return params


class Engine(object):
    def __init__(self, cfg, max_clips, device="cuda:0", params=None, rank=0, world=1, group=None):
        nv.lib()  # fail loudly when the CUDA library is missing: there is no CPU fallback
        if not torch.cuda.is_available():
            raise nv.NativeError("vlb200.Engine needs a CUDA device (sm_100a); no CPU fallback exists")
        self.cfg = cfg
        self.dev = torch.device(device)
        torch.cuda.set_device(self.dev)
        self.rank, self.world, self.group = rank, world, group
        self.max_clips = int(max_clips)
        self.max_frames = self.max_clips * cfg.fpc
        self.sp = encoder_specs(cfg.height, cfg.width)
        self.global_step = 0
        self.adam_t = 0
        self.c_pad = _align(cfg.num_classes, 8)

        # ---- parameter / gradient arenas ----
        self.var_shapes = variable_shapes(cfg)
        self.var_off = {}
        off = 0
        offsets = [0]
        for name, shape in self.var_shapes:
            self.var_off[name] = off
            off += _align(int(np.prod(shape)))
            offsets.append(off)
        self.arena_n = off
        self.params = torch.zeros(off, dtype=F32, device=self.dev)
        # [64-float header | gradient arena | scratch for the filter gradient of the space-to-depth conv1].  The header
        # holds the step scalars ([0..2] clip scalars, [4] loss, [5] correct): [loss, correct] sit right in front of the
        # convolution gradients, so the tail all-reduce of a data-parallel step is ONE collective over
        # [header | conv1..conv5 gradients] instead of two
        s1s = self.sp["conv1_s2d"]
        self._dws_n = _align(s1s.taps * s1s.cin_g * s1s.cout)
        self._hdr = 64
        self.grads_ext = torch.zeros(self._hdr + off + self._dws_n, dtype=F32, device=self.dev)
        self.grads = self.grads_ext[self._hdr:self._hdr + off]
        self.dws1 = self.grads_ext[self._hdr + off:self._hdr + off + s1s.taps * s1s.cin_g * s1s.cout].view(
            s1s.taps * s1s.cin_g, s1s.cout)
        self.seg_offsets = torch.tensor(offsets, dtype=torch.int64, device=self.dev)
        # the variables from fc6W on (89 % of the arena) have their final gradients as soon as fc6's filter gradient is
        # enqueued: their squared norms are taken on a side stream under the convolution backward
        self._k_late = [n for n, _ in self.var_shapes].index("dcnn/fc6W")
        off_late = offsets[self._k_late]
        self._seg_early = torch.tensor(offsets[:self._k_late + 1], dtype=torch.int64, device=self.dev)
        self._seg_late = torch.tensor([o - off_late for o in offsets[self._k_late:]], dtype=torch.int64, device=self.dev)
        # scratch of the two (possibly concurrent) deterministic norm reductions
        n_late_vars = len(self.var_shapes) - self._k_late
        self._ws_late = torch.empty(int(nv.lib().vl_grad_sqnorms_workspace(off - off_late, n_late_vars)), dtype=F32,
                                    device=self.dev)
        self._ws_early = torch.empty(int(nv.lib().vl_grad_sqnorms_workspace(off_late, self._k_late)), dtype=F32,
                                     device=self.dev)
        self.sqnorms = torch.zeros(len(self.var_shapes), dtype=F32, device=self.dev)
        self.scalars = self.grads_ext[:8]
        self.adam_m = self.adam_v = None
        if cfg.optimizer == "adam":
            self.adam_m = torch.zeros(off, dtype=F32, device=self.dev)
            self.adam_v = torch.zeros(off, dtype=F32, device=self.dev)
        self._side = torch.cuda.Stream(device=self.dev)  # filter gradients (off the backward critical path)
        self._side2 = torch.cuda.Stream(device=self.dev)  # second half-batch chain of the conv1/conv2 backward
        # read-back of the step scalars: they are final after vl_clip_scalars, so the host copy waits on an event recorded
        # there (not on the optimiser update / operand refresh that follow) and the next step is enqueued while the
        # tail of this one still runs - the device never idles between steps
        self.read_resize = None
        self._resizer = None
        self._rb_stream = torch.cuda.Stream(device=self.dev)
        self._scalars_host = torch.zeros(8, dtype=F32).pin_memory()
        self._scalars_ready = None
        self._norm_stream = torch.cuda.Stream(device=self.dev)  # early gradient norms (+ the early all-reduce wait)
        self._late_norms_ready = None
        self._stage_stream = torch.cuda.Stream(device=self.dev)  # input staging (independent of the weights)
        self._xs2d_free = None     # recorded after the last reader of the staging buffer (conv1 filter gradient)
        self._frames_ready = None  # optional event of the caller: the frames are complete (else: the current stream)
        self._streams = (self._side, self._side2, self._norm_stream, self._stage_stream)
        self._alloc_shadows()
        self._alloc_activations()
        self.load_state_dict(params if params is not None else init_variables(cfg))

    # ------------------------------------------------------------------------------------------
    # variables
    # ------------------------------------------------------------------------------------------
    def var(self, name, arena=None):
        arena = self.params if arena is None else arena
        shape = dict(self.var_shapes)[name]
        o = self.var_off[name]
        return arena[o:o + int(np.prod(shape))].view(*shape)

    def var2d(self, name, arena=None):
        v = self.var(name, arena)
        return v.reshape(-1, v.shape[-1])

    def state_dict(self):
        """Variables as float32 numpy arrays keyed by the reference's TF variable names (+ global_step)."""
        out = {name: self.var(name).detach().cpu().numpy().copy() for name, _ in self.var_shapes}
        out["global_step"] = np.int32(self.global_step)
        return out

    def load_state_dict(self, sd, strict=True):
        for name, shape in self.var_shapes:
            if name not in sd:
                if strict:
                    raise KeyError("missing variable %s" % name)
                continue
            arr = np.asarray(sd[name], dtype=np.float32)
            if tuple(arr.shape) != tuple(shape):
                raise ValueError("variable %s: shape %s does not match %s" % (name, arr.shape, shape))
            self.var(name).copy_(torch.from_numpy(arr))
        if "global_step" in sd:
            self.global_step = int(sd["global_step"])
        self.refresh_shadows()

    def optimizer_state_dict(self):
        """Optimiser slots under the names tf.train.Saver gives them (the reference saves all global variables,
        feeder.py:263-288): `<var>/Adam` (m), `<var>/Adam_1` (v), `beta1_power` / `beta2_power` (= beta^(t+1) after t
        steps).  Empty for SGD."""
        if self.cfg.optimizer != "adam":
            return {}
        out = {}
        for name, _ in self.var_shapes:
            out[name + "/Adam"] = self.var(name, self.adam_m).detach().cpu().numpy().copy()
            out[name + "/Adam_1"] = self.var(name, self.adam_v).detach().cpu().numpy().copy()
        out["beta1_power"] = np.float32(0.9 ** (self.adam_t + 1))
        out["beta2_power"] = np.float32(0.999 ** (self.adam_t + 1))
        return out

    def load_optimizer_state_dict(self, sd):
        """Inverse of optimizer_state_dict; slots that are absent keep their zero initialisation (a checkpoint written
        by an SGD run, or by a version that did not save them)."""
        if self.cfg.optimizer != "adam":
            return 0
        loaded = 0
        for name, shape in self.var_shapes:
            for suffix, arena in (("/Adam", self.adam_m), ("/Adam_1", self.adam_v)):
                if name + suffix in sd:
                    arr = np.asarray(sd[name + suffix], dtype=np.float32)
                    if tuple(arr.shape) != tuple(shape):
                        raise ValueError("optimizer slot %s: shape %s does not match %s" % (name + suffix, arr.shape, shape))
                    self.var(name, arena).copy_(torch.from_numpy(arr))
                    loaded += 1
        if "beta1_power" in sd:
            self.adam_t = max(0, int(round(math.log(float(sd["beta1_power"])) / math.log(0.9))) - 1)
        return loaded

    def gradient_dict(self):
        return {name: self.var(name, self.grads).detach().cpu().numpy().copy() for name, _ in self.var_shapes}

    # ------------------------------------------------------------------------------------------
    # bf16 operand copies of the weights
    # ------------------------------------------------------------------------------------------
    def _alloc_shadows(self):
        """bf16 operand copies of the weights.  The permuted / padded copies of the convolution filters (and of the
        narrow fc8 / output_fc matrices) live in ONE bf16 arena described by an index table (shadow_table.py): one
        vl_gather_bf16 launch refreshes all of them.  The big same-layout copies (fc6, fc7, LSTM kernels) are separate
        tensors written by the optimiser kernel itself (vl_sgd_update_shadow) or by vl_cast_f32_to_bf16."""
        from . import shadow_table as ST
        cfg, sp, dev = self.cfg, self.sp, self.dev
        names = dict(self.var_shapes)
        plan = []  # (key, shape, table of source indices relative to the master variable, master variable name)
        s1, s1s = sp["conv1"], sp["conv1_s2d"]
        plan.append(("conv1_fwd", (96, s1s.k_packed),
                     ST.s2d_filter_kmajor(s1.kh, s1.kw, 3, 96, s1.stride, s1s.cchunks * 64), "dcnn/conv1W"))
        for name in ("conv2", "conv3", "conv4", "conv5"):
            sc = sp[name]
            rows = sc.taps * sc.cin_g
            plan.append((name, (rows, sc.cout), ST.identity(rows * sc.cout), "dcnn/%sW" % name))  # HWIO as 2D (dgrad)
            plan.append((name + "_fwd", (sc.cout, sc.k_packed),
                         ST.kmajor_padded(sc.taps, sc.cin_g, sc.cout, sc.cchunks * 64), "dcnn/%sW" % name))
        # conv2's data gradient runs as a stride-2 forward convolution over dy with a depth-to-space epilogue
        # (kernels.conv_dgrad_d2s): N = 2*2*48 = 192 columns per UMMA instead of 48
        s2 = sp["conv2"]
        plan.append(("conv2_d2s", K.d2s_filter_shape(s2, 2, 2),
                     ST.dgrad_d2s(s2.kh, s2.kw, s2.cin_g, s2.cout_g, s2.groups, 2, 2), "dcnn/conv2W"))
        if "dcnn/fc8W" in names:
            plan.append(("fc8", (4096, self.c_pad), ST.col_padded(4096, cfg.num_classes, self.c_pad), "dcnn/fc8W"))
        for fc_name in ("output_fc", "fc_convert"):
            if fc_name + "_w" in names:
                rows = names[fc_name + "_w"][0]
                plan.append((fc_name, (rows, self.c_pad), ST.col_padded(rows, cfg.num_classes, self.c_pad),
                             fc_name + "_w"))
        total = 0
        offs = []
        for key, shape, table, master in plan:
            assert table.size == int(np.prod(shape)), key
            offs.append(total)
            total += _align(table.size)
        table_all = np.full(total, -1, dtype=np.int32)
        for (key, shape, table, master), o in zip(plan, offs):
            table_all[o:o + table.size] = np.where(table >= 0, table + self.var_off[master], -1)
        self._shadow_table = torch.from_numpy(table_all).to(dev)
        self._shadow_arena = torch.zeros(total, dtype=BF16, device=dev)
        sh = {}
        for (key, shape, table, master), o in zip(plan, offs):
            sh[key] = self._shadow_arena[o:o + table.size].view(*shape)
        sh["fc6"] = torch.zeros(sp["flat"], 4096, dtype=BF16, device=dev)
        if "dcnn/fc7W" in names:
            sh["fc7"] = torch.zeros(4096, 4096, dtype=BF16, device=dev)
        if cfg.workflow == "lrcn":
            for layer in range(cfg.lstm_layers):
                kn = "rnn/multi_rnn_cell/cell_%d/basic_lstm_cell/kernel" % layer
                rows, cols = names[kn]
                sh["lstm%d" % layer] = torch.zeros(rows, cols, dtype=BF16, device=dev)
                sh["lstm%d_wht" % layer] = torch.zeros(cols, cfg.lstm_hidden, dtype=F32, device=dev)
        self.sh = sh

    def _plain_shadow_segments(self):
        """(begin, end, shadow) of the variables whose bf16 operand copy has the master's layout (fc6, fc7 and the LSTM
        kernels): vl_sgd_update_shadow writes them while the updated weights are in registers."""
        segs = []
        for name, key in (("dcnn/fc6W", "fc6"), ("dcnn/fc7W", "fc7")):
            if key in self.sh and name in self.var_off:
                segs.append((self.var_off[name], self.var_off[name] + self.sh[key].numel(), self.sh[key]))
        if self.cfg.workflow == "lrcn":
            for layer in range(self.cfg.lstm_layers):
                name = "rnn/multi_rnn_cell/cell_%d/basic_lstm_cell/kernel" % layer
                t = self.sh["lstm%d" % layer]
                segs.append((self.var_off[name], self.var_off[name] + t.numel(), t))
        segs = [sg for sg in segs if sg[0] % 4 == 0 and sg[1] % 4 == 0]
        return segs[:4] if len(segs) <= 4 else None

    def refresh_shadows(self, plain_done=False):
        """fp32 master -> bf16 tensor-core operands (after load and after every optimiser step).  plain_done: the
        same-layout copies (fc6, fc7, LSTM kernels) were already written by vl_sgd_update_shadow."""
        sh = self.sh
        self.weights_version = getattr(self, "weights_version", 0) + 1  # derived operand sets (fp32_path) rebuild lazily
        nv.call("vl_gather_bf16", self.params, self._shadow_table, self._shadow_arena, self._shadow_arena.numel())
        for name in ("fc6", "fc7"):
            if name in sh and not plain_done:
                w = self.var("dcnn/%sW" % name)
                nv.call("vl_cast_f32_to_bf16", w, sh[name], w.numel())
        if self.cfg.workflow == "lrcn":
            hdim = self.cfg.lstm_hidden
            for layer in range(self.cfg.lstm_layers):
                kern = self.var("rnn/multi_rnn_cell/cell_%d/basic_lstm_cell/kernel" % layer)
                if not plain_done:
                    nv.call("vl_cast_f32_to_bf16", kern, sh["lstm%d" % layer], kern.numel())
                if hdim != 256:  # the generic BPTT kernel reads w_h^T; the cluster kernel (H = 256) reads w_h itself
                    d_in = kern.shape[0] - hdim
                    nv.call("vl_transpose_f32", kern[d_in:], sh["lstm%d_wht" % layer], hdim, 4 * hdim)

    # ------------------------------------------------------------------------------------------
    # buffers
    # ------------------------------------------------------------------------------------------
    def _alloc_activations(self):
        cfg, sp, dev = self.cfg, self.sp, self.dev
        n, b = self.max_frames, self.max_clips
        s1, s2, s3 = sp["conv1"], sp["conv2"], sp["conv3"]
        (p1h, p1w), (p2h, p2w), (p5h, p5w) = sp["pool1"], sp["pool2"], sp["pool5"]

        def act(*shape, dtype=BF16):
            return torch.empty(*shape, dtype=dtype, device=dev)

        A = {}
        self._frame_bufs = {}  # host-fed frames are staged per stored shape / dtype (allocated on first use)
        self._crops_dev = torch.zeros(n, 3, dtype=torch.int32, device=dev)
        s1s = sp["conv1_s2d"]
        A["x_s2d"] = act(n, s1s.h, s1s.w, s1s.cin)
        A["a1"] = act(n, s1.p, s1.q, 96)
        A["p1"] = act(n, p1h, p1w, 96)
        A["arg1"] = act(n, p1h, p1w, 96, dtype=torch.uint8)
        A["a2"] = act(n, s2.p, s2.q, 256)
        A["p2"] = act(n, p2h, p2w, 256)
        A["arg2"] = act(n, p2h, p2w, 256, dtype=torch.uint8)
        A["a3"] = act(n, s3.p, s3.q, 384)
        A["a4"] = act(n, s3.p, s3.q, 384)
        A["a5"] = act(n, s3.p, s3.q, 256)
        A["p5"] = act(n, p5h, p5w, 256)
        A["arg5"] = act(n, p5h, p5w, 256, dtype=torch.uint8)
        A["f6"] = act(n, 4096)
        A["f7"] = act(n, 4096)
        c, cp, hd = cfg.num_classes, self.c_pad, cfg.lstm_hidden
        if cfg.workflow == "fc" and cfg.early_fusion:
            A["pooled"] = act(b, 4096, dtype=F32)
            A["pooled_bf"] = act(b, 4096)
            A["dpooled"] = act(b, 4096, dtype=F32)
        elif cfg.workflow in ("singleframe", "fc"):
            A["frame_logits"] = act(n, c, dtype=F32)
            A["d_frame_logits"] = act(n, cp, dtype=F32)
            A["d_frame_logits_bf16"] = act(n, cp)
        else:
            for layer in range(cfg.lstm_layers):
                A["gx%d" % layer] = act(n, 4 * hd, dtype=F32)
                A["acts%d" % layer] = act(n, 4 * hd, dtype=F32)
                A["cs%d" % layer] = act(n, hd, dtype=F32)
                A["hseq%d" % layer] = act(n, hd, dtype=F32)
                A["hseq_bf%d" % layer] = act(n, hd)
                A["hprev_bf%d" % layer] = act(n, hd)
                A["dhseq%d" % layer] = act(n, hd, dtype=F32)
            A["dg"] = act(n, 4 * hd)
            A["fused"] = act(b, hd, dtype=F32)
            A["fused_bf"] = act(b, hd)
            A["mask"] = act(b, hd, dtype=F32)
            A["dropped"] = act(b, hd, dtype=F32)
            A["dropped_bf"] = act(b, hd)
            A["ddropped"] = act(b, hd, dtype=F32)
            A["dfused"] = act(b, hd, dtype=F32)
        A["logits"] = act(b, c, dtype=F32)
        A["labels"] = act(b, c, dtype=torch.int32)
        A["row_loss"] = act(2 * b, dtype=F32)
        A["dlogits"] = torch.zeros(b, cp, dtype=F32, device=dev)
        A["dlogits_bf"] = torch.zeros(b, cp, dtype=BF16, device=dev)
        self.A = A
        self.G = None  # gradient-side activation buffers, allocated on the first train_step
        self._pinned = {}
        self._pinned_busy = {}  # pinned staging buffer -> event of the last asynchronous H2D copy that read it

    def _alloc_backward(self):
        sp, dev = self.sp, self.dev
        n = self.max_frames
        s1, s2, s3 = sp["conv1"], sp["conv2"], sp["conv3"]
        (p1h, p1w), (p2h, p2w), (p5h, p5w) = sp["pool1"], sp["pool2"], sp["pool5"]

        def act(*shape):
            return torch.empty(*shape, dtype=BF16, device=dev)

        G = {}
        G["df7"] = act(n, 4096)
        G["df6"] = act(n, 4096)
        G["dp5"] = act(n, p5h, p5w, 256)
        G["da5"] = act(n, s3.p, s3.q, 256)
        G["da4"] = act(n, s3.p, s3.q, 384)
        G["da3"] = act(n, s3.p, s3.q, 384)
        G["dp2"] = act(n, p2h, p2w, 256)
        G["da2"] = act(n, s2.p, s2.q, 256)
        G["dp1"] = act(n, p1h, p1w, 96)
        G["da1"] = act(n, s1.p, s1.q, 96)
        self.G = G

    # ------------------------------------------------------------------------------------------
    # input staging
    # ------------------------------------------------------------------------------------------
    def _frame_buffer(self, shape_hw, dtype):
        """Device staging buffer [max_frames, h, w, 3] for host-fed frames of one stored shape (allocated on first use)."""
        key = (int(shape_hw[0]), int(shape_hw[1]), dtype)
        buf = self._frame_bufs.get(key)
        if buf is None:
            buf = torch.empty(self.max_frames, key[0], key[1], 3, dtype=dtype, device=self.dev)
            self._frame_bufs[key] = buf
        return buf

    def _stage_frames(self, frames):
        """Accept what `feeder.get_feed_dict` produces (a list / array of float32 HWC frames, feeder.py:97-100),
        a uint8 array (mean subtracted on device), or a device tensor; frames may be stored larger than the network
        input (raw_image_shape) when crop offsets accompany them.  Returns (device tensor, is_u8, n_frames)."""
        if isinstance(frames, (list, tuple)):
            frames = np.stack([np.asarray(f) for f in frames], axis=0)
        pinned_key = None
        if isinstance(frames, np.ndarray):
            if frames.dtype != np.uint8:
                frames = np.ascontiguousarray(frames, dtype=np.float32)
            host = torch.from_numpy(np.ascontiguousarray(frames))
            key = (tuple(host.shape[1:3]), host.dtype)
            pin = self._pinned.get(key)
            if pin is None:
                pin = torch.empty(self.max_frames, host.shape[1], host.shape[2], 3, dtype=host.dtype).pin_memory()
                self._pinned[key] = pin
            if host.shape[0] > self.max_frames:
                raise ValueError("batch of %d frames exceeds the engine capacity %d" % (host.shape[0], self.max_frames))
            ev = self._pinned_busy.get(key)
            if ev is not None:
                ev.synchronize()  # the previous asynchronous H2D copy out of this pinned buffer has finished
            pin[:host.shape[0]].copy_(host)
            frames = pin[:host.shape[0]]
            pinned_key = key
        assert isinstance(frames, torch.Tensor) and frames.is_contiguous() and frames.dtype in (torch.uint8, F32)
        n = frames.shape[0]
        if n > self.max_frames:
            raise ValueError("batch of %d frames exceeds the engine capacity %d" % (n, self.max_frames))
        if not frames.is_cuda:
            # host tensor (ideally pinned): one asynchronous H2D copy straight into the device staging buffer
            dst = self._frame_buffer(frames.shape[1:3], frames.dtype)
            dst[:n].copy_(frames, non_blocking=True)
            if pinned_key is not None:
                ev = torch.cuda.Event()
                ev.record()
                self._pinned_busy[pinned_key] = ev
            frames = dst[:n]
        if self.read_resize is not None and frames.dtype == torch.uint8 and \
                tuple(int(x) for x in frames.shape[1:3]) != tuple(self.read_resize):
            # imgproc raw_resize / resize of the reference (dataset_.py:481-491): scipy.misc.imresize at read time
            if self._resizer is None:
                from .resize import DeviceResizer
                self._resizer = DeviceResizer(self.dev)
            frames = self._resizer.resize(frames, int(self.read_resize[0]), int(self.read_resize[1]))
            self._frames_ready = None  # produced on the caller's stream just now
        return frames, frames.dtype == torch.uint8, n

    def set_read_resize(self, hw):
        """Resample uint8 frames to (h, w) on the device before crop / staging (None: frames arrive at their final
        stored size).  Mirrors `imresize(image, raw_image_shape)` / `imresize(image, desired_image_shape)`."""
        self.read_resize = None if hw is None else (int(hw[0]), int(hw[1]))

    def _stage_crops(self, crops, frames, n):
        """Per-frame (y0, x0, mirror) int32 triples on the device; None when the frames already have the input shape."""
        hr, wr = int(frames.shape[1]), int(frames.shape[2])
        if crops is None:
            if (hr, wr) != (self.cfg.height, self.cfg.width):
                raise ValueError("frames are stored %dx%d but the network input is %dx%d: crop offsets are required" % (
                    hr, wr, self.cfg.height, self.cfg.width))
            return None
        if not isinstance(crops, torch.Tensor):
            crops = torch.from_numpy(np.ascontiguousarray(np.asarray(crops, dtype=np.int32)))
        if tuple(crops.shape) != (n, 3):
            raise ValueError("crops must be int32 [%d, 3] (y0, x0, mirror), got %s" % (n, tuple(crops.shape)))
        if not crops.is_cuda:
            c = crops.numpy()
            if (c[:, 0] < 0).any() or (c[:, 1] < 0).any() or (c[:, 0] + self.cfg.height > hr).any() or \
                    (c[:, 1] + self.cfg.width > wr).any():
                raise ValueError("crop window leaves the stored %dx%d frame" % (hr, wr))
            self._crops_dev[:n].copy_(crops, non_blocking=True)
            self._frames_ready = None  # the offsets travel on the caller's stream: the staging must wait for it
            return self._crops_dev[:n]
        # device-resident offsets: not validated here (that would synchronise); the staging kernel clamps them
        # into the stored frame, so an out-of-range window cannot read outside the buffer
        return crops.to(torch.int32).contiguous()

    def prefetch(self, frames_pinned, onehot_pinned, slot):
        """Enqueue the H2D copy of one batch (pinned host tensors: uint8 frames [n,H,W,3], int32 one-hot [b,C]) on the
        copy stream into device slot `slot` (any small index; one buffer pair per slot) and return (frames_dev, onehot_dev, copied, consumed): the compute
        stream must wait for the event `copied` before `train_step(frames_dev, onehot_dev, ...)` and record
        `consumed` after it (the next copy into this slot waits for it); with two slots the copy of batch i+1
        overlaps the step of batch i, with three the copy engine never waits for a slot."""
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream(device=self.dev)
            self._slots = {}
        n, b = frames_pinned.shape[0], onehot_pinned.shape[0]
        key = (slot, frames_pinned.dtype)
        if key not in self._slots:
            self._slots[key] = (torch.empty(self.max_frames, self.cfg.height, self.cfg.width, 3,
                                            dtype=frames_pinned.dtype, device=self.dev),
                                torch.empty(self.max_clips, self.cfg.num_classes, dtype=torch.int32, device=self.dev),
                                torch.cuda.Event())
        fd, od, done = self._slots[key]
        self._copy_stream.wait_event(done)  # the step that last read this slot has finished
        with torch.cuda.stream(self._copy_stream):
            fd[:n].copy_(frames_pinned, non_blocking=True)
            od[:b].copy_(onehot_pinned, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        return fd[:n], od[:b], ev, done

    # ------------------------------------------------------------------------------------------
    # forward
    # ------------------------------------------------------------------------------------------
    def _mean_dev(self):
        if not hasattr(self, "_mean_t"):
            m = self.cfg.mean if self.cfg.mean is not None else (0.0, 0.0, 0.0)
            self._mean_t = torch.tensor(m, dtype=F32, device=self.dev)
        return self._mean_t

    def _encoder_fwd(self, frames, is_u8, n, training, crops=None):
        A, sp, sh = self.A, self.sp, self.sh
        s1 = sp["conv1"]
        (p1h, p1w), (p2h, p2w), (p5h, p5w) = sp["pool1"], sp["pool2"], sp["pool5"]
        s1s = sp["conv1_s2d"]
        xs = A["x_s2d"][:n]
        # The staging kernel depends on the frames only, not on the weights: it runs on its own stream and waits for
        # (a) the frames (whatever the caller's stream has enqueued so far) and (b) the last reader of the staging
        # buffer (conv1's filter gradient of the previous step) - NOT for the optimiser tail of the previous step, so
        # the HBM-bound staging of step i+1 runs next to the norm / update / operand-refresh kernels of step i.
        main = torch.cuda.current_stream()
        # (stream OBJECTS differ from call to call of torch.cuda.current_stream(): compare the handles)
        overlap = self._stage_stream.cuda_stream != main.cuda_stream and os.environ.get("VL_STAGE_OVERLAP", "1") != "0"
        st = self._stage_stream if overlap else main
        if overlap:
            if self._frames_ready is None:  # unknown producer: everything enqueued on the caller's stream so far
                ev = torch.cuda.Event()
                ev.record(main)
                st.wait_event(ev)
            elif self._frames_ready is not True:  # the caller's event (e.g. the H2D copy of Engine.prefetch)
                st.wait_event(self._frames_ready)
            if self._xs2d_free is not None:
                st.wait_event(self._xs2d_free)
        with torch.cuda.stream(st):
            nv.call("vl_frames_s2d_crop", frames, 1 if is_u8 else 0, self._mean_dev(), xs, n, int(frames.shape[1]),
                    int(frames.shape[2]), crops, self.cfg.height, self.cfg.width, s1.stride, s1.pad_top, s1.pad_left,
                    s1s.h, s1s.w)
            if overlap:
                staged = torch.cuda.Event()
                staged.record(st)
        if overlap:
            main.wait_event(staged)
        self._frames_ready = None
        s2, s3 = sp["conv2"], sp["conv3"]

        def conv_chain(lo, hi):
            m = hi - lo
            K.conv_fwd_flat(s1s, xs[lo:hi], sh["conv1_fwd"], self.var("dcnn/conv1b"), A["a1"][lo:hi], relu=True,
                            flops=K.conv_flops(s1, m))  # tap-shifted kernel; algorithmic FLOPs of the 11x11x3 layer
            nv.call("vl_lrn_pool_fwd", A["a1"][lo:hi], A["p1"][lo:hi], A["arg1"][lo:hi], m, s1.p, s1.q, 96, LRN["radius"],
                    LRN["alpha"], LRN["beta"], LRN["bias"])
            K.conv_fwd_flat(s2, A["p1"][lo:hi], sh["conv2_fwd"], self.var("dcnn/conv2b"), A["a2"][lo:hi], relu=True)
            nv.call("vl_lrn_pool_fwd", A["a2"][lo:hi], A["p2"][lo:hi], A["arg2"][lo:hi], m, s2.p, s2.q, 256, LRN["radius"],
                    LRN["alpha"], LRN["beta"], LRN["bias"])
            K.conv_fwd(sp["conv3"], A["p2"][lo:hi], sh["conv3_fwd"], self.var("dcnn/conv3b"), A["a3"][lo:hi], relu=True)
            K.conv_fwd(sp["conv4"], A["a3"][lo:hi], sh["conv4_fwd"], self.var("dcnn/conv4b"), A["a4"][lo:hi], relu=True)
            K.conv_fwd_flat(sp["conv5"], A["a4"][lo:hi], sh["conv5_fwd"], self.var("dcnn/conv5b"), A["a5"][lo:hi], relu=True)
            nv.call("vl_maxpool_fwd", A["a5"][lo:hi], A["p5"][lo:hi], A["arg5"][lo:hi], m, s3.p, s3.q, 256)

        # The convolution stack runs as two half-batch chains on two streams (VL_FWD_HALVES=0: one chain), so that the
        # issue-bound LRN / pool kernels of one half run next to the tensor-bound convolutions of the other: the
        # register-only LRN forward holds no shared memory, so one of its CTAs fits beside a persistent contraction CTA
        # (only one: the contraction kernels hold 41 K of the 64 K registers).  Measured (profiles/r02_step_ab.txt):
        # forward 2.04 -> 1.99 ms, train step -0.2 % (min) / -1.7 % (median) in a same-box A/B.
        fwd2 = self._side2
        if n >= 256 and n % 2 == 0 and fwd2.cuda_stream != main.cuda_stream and \
                os.environ.get("VL_FWD_HALVES", "1") == "1":
            ev = torch.cuda.Event()
            ev.record(main)
            fwd2.wait_event(ev)
            conv_chain(0, n // 2)
            with torch.cuda.stream(fwd2):
                conv_chain(n // 2, n)
                ev2 = torch.cuda.Event()
                ev2.record(fwd2)
            main.wait_event(ev2)
        else:
            conv_chain(0, n)
        flat = A["p5"][:n].view(n, sp["flat"])  # HWC-major flatten (alexnet.py:228)
        K.linear_fwd(flat, sh["fc6"], self.var("dcnn/fc6b"), A["f6"][:n], relu=True)
        feat = A["f6"][:n]
        if "fc7" in sh:
            K.linear_fwd(A["f6"][:n], sh["fc7"], self.var("dcnn/fc7b"), A["f7"][:n], relu=True)
            feat = A["f7"][:n]
        return feat

    def _head_fwd(self, feat, n, training):
        cfg, A, sh = self.cfg, self.A, self.sh
        b = n // cfg.fpc
        c = cfg.num_classes
        logits = A["logits"][:b]
        if cfg.workflow == "fc" and cfg.early_fusion:
            # aggregate_clip_vectors on the features (model.py:103-108), then convert_dim_fc (model.py:115-117)
            dim = feat.shape[1]
            nv.call("vl_segment_pool_fwd_bf16", feat, cfg.fpc, b, dim, POOL[cfg.fusion], A["pooled"][:b],
                    A["pooled_bf"][:b])
            K.linear_fwd(A["pooled_bf"][:b], sh["fc_convert"], self.var("fc_convert_b"), logits, relu=False, n=c)
            return logits
        if cfg.workflow in ("singleframe", "fc"):
            key, bname = self._frame_classifier()
            fl = A["frame_logits"][:n]
            K.linear_fwd(feat, sh[key], self.var(bname), fl, relu=False, n=c)
            nv.call("vl_segment_pool_fwd", fl, None, cfg.fpc, b, c, POOL[cfg.fusion], logits, None)
            return logits
        hd, t_len = cfg.lstm_hidden, cfg.fpc
        x = feat
        for layer in range(cfg.lstm_layers):
            kern = self.var("rnn/multi_rnn_cell/cell_%d/basic_lstm_cell/kernel" % layer)
            bias = self.var("rnn/multi_rnn_cell/cell_%d/basic_lstm_cell/bias" % layer)
            d_in = kern.shape[0] - hd
            gx = A["gx%d" % layer][:n]
            K.linear_fwd(x, sh["lstm%d" % layer][:d_in], bias, gx, relu=False)
            nv.call("vl_lstm_fwd_cluster" if hd == 256 else "vl_lstm_fwd", gx, kern[d_in:], A["acts%d" % layer][:n],
                    A["cs%d" % layer][:n], A["hseq%d" % layer][:n], A["hseq_bf%d" % layer][:n],
                    A["hprev_bf%d" % layer][:n], b, t_len, hd, 1.0)
            x = A["hseq_bf%d" % layer][:n]
        hseq = A["hseq%d" % (cfg.lstm_layers - 1)][:n]
        # fusion `state`: the final hidden state of the top layer = its output at the last step (full-length sequences)
        state = cfg.fusion == "state"
        nv.call("vl_segment_pool_fwd", hseq, None, t_len, b, hd, POOL["last" if state else cfg.fusion], A["fused"][:b],
                A["fused_bf"][:b])
        top, top_bf = A["fused"][:b], A["fused_bf"][:b]
        self._dropout_on = bool(training and cfg.dropout_keep_prob > 0 and not state)  # lstm.py:52; no dropout on state
        if self._dropout_on:
            if self._injected_mask is not None:
                A["mask"][:b].copy_(self._injected_mask)
            else:
                nv.call("vl_dropout_mask", A["mask"][:b], b * hd, cfg.dropout_keep_prob, cfg.seed + 7919 * self.rank,
                        self.global_step * ((self.max_clips * hd + 3) // 4))
            nv.call("vl_mul", top, A["mask"][:b], A["dropped"][:b], A["dropped_bf"][:b], b * hd)
            top, top_bf = A["dropped"][:b], A["dropped_bf"][:b]
        self._top_bf = top_bf
        fc = "fc_convert" if state else "output_fc"
        if fc in sh:
            K.linear_fwd(top_bf, sh[fc], self.var(fc + "_b"), logits, relu=False, n=c)
        else:
            logits.copy_(top)
        return logits

    def _frame_classifier(self):
        """(operand key, bias variable) of the per-frame classifier: fc8 inside the dcnn (single-frame workflow) or
        convert_dim_fc on top of fc6 / fc7 features (classifier fc with late fusion)."""
        return ("fc8", "dcnn/fc8b") if self.cfg.workflow == "singleframe" else ("fc_convert", "fc_convert_b")

    _injected_mask = None
    bwd_smem_reserve = 0  # bytes (see train_step); set from the A/B measurement in profiles/r02_step_ab.txt

    def forward_device(self, frames, training=False, crops=None, frames_ready=None):
        """Enqueue the forward pass; returns the device logits [clips, C] (fp32).  `crops` (int32 [n, 3]: y0, x0, mirror)
        applies the reference's read-time crop / mirror (dataset_.py:444-461,498-500) on the device.  frames_ready: see
        train_step."""
        self._frames_ready = frames_ready if isinstance(frames, torch.Tensor) and frames.is_cuda else None
        frames, is_u8, n = self._stage_frames(frames)
        if n % self.cfg.fpc != 0:
            raise ValueError("number of frames (%d) is not a multiple of num_frames_per_clip (%d)" % (n, self.cfg.fpc))
        feat = self._encoder_fwd(frames, is_u8, n, training, self._stage_crops(crops, frames, n))
        return self._head_fwd(feat, n, training)

    def forward_features(self, frames, crops=None):
        """The dcnn feature vectors of every frame (bf16 device tensor [frames, dim] of the configured
        frame_encoding_layer): the output of a classifier-less pipeline (model.py:110-112) that a later pipeline
        consumes (fusion.py)."""
        frames, is_u8, n = self._stage_frames(frames)
        return self._encoder_fwd(frames, is_u8, n, False, self._stage_crops(crops, frames, n))

    def forward(self, frames, crops=None):
        """`sess.run(model.logits, fdict)` (run_task.py:95): float32 ndarray [clips, C] on the host.
        VLB200_FP32=1 in the environment routes it (and with it the validation workflow of run_task / tfshim) through
        the fp32-accuracy mode."""
        if os.environ.get("VLB200_FP32", "0") == "1":
            return self.forward_fp32(frames, crops)
        return self.forward_device(frames, training=False, crops=crops).cpu().numpy()

    def forward_fp32(self, frames, crops=None):
        """The same logits in fp32 ACCURACY (<= 1e-3 relative against the reference's fp32 graph; north-star tolerance):
        every tensor stays fp32 and the contractions run as hi / lo bf16 split products on the tensor cores
        (fp32_path.py).  About 3x the cost of `forward`; for validation read-outs."""
        if getattr(self, "_fp32", None) is None:
            from .fp32_path import Fp32Path
            self._fp32 = Fp32Path(self)
        return self._fp32.forward_device(frames, crops=crops).cpu().numpy()

    # ------------------------------------------------------------------------------------------
    # backward
    # ------------------------------------------------------------------------------------------
    def _split_k(self, out_rows, out_cols, contraction, groups=1, block_n=None):
        tiles = -(-out_rows // 128) * -(-out_cols // (block_n or min(256, _align(out_cols, 64)))) * groups
        kb = -(-contraction // 64)
        want = max(1, (2 * nv.lib().vl_device_sm_count()) // max(tiles, 1))
        return int(max(1, min(want, kb // 4 if kb >= 8 else 1)))

    def _zero_grads(self, n):
        """Zero the gradient arena where the backward pass ACCUMULATES (split-K red.add filter gradients, atomically
        summed bias gradients).  The fc6 / fc7 filter gradients (89 % of the arena) are plain stores whenever their
        GEMM runs unsplit, which is the case at every batch size that fills the tile grid: they are skipped."""
        skip = []
        for name in ("dcnn/fc6W", "dcnn/fc7W"):
            if name in self.var_off:
                rows, cols = self.var2d(name).shape
                if self._split_k(rows, cols, n) == 1:
                    skip.append((self.var_off[name], self.var_off[name] + rows * cols))
        body = self.grads_ext[self._hdr:]  # the header (step scalars, written by vl_softmax_ce) is not touched
        lo = 0
        total = body.numel()
        for b, e in sorted(skip):
            if b > lo:
                nv.call("vl_zero", body[lo:b], (b - lo) * 4)
            lo = e
        if total > lo:
            nv.call("vl_zero", body[lo:], (total - lo) * 4)

    def _dense_bwd(self, x, dy, wname, bname, n_valid=None):
        """Filter / bias gradient of y = x @ W + b into the gradient arena."""
        dw = self.var2d(wname, self.grads)
        n_cols = dw.shape[1]
        K.linear_wgrad(x, dy, dw, split_k=self._split_k(dw.shape[0], n_cols, x.shape[0]), n=n_cols)
        nv.call("vl_colsum", dy, self.var(bname, self.grads), dy.shape[0], n_cols, dy.stride(0))

    def _conv_bwd(self, name, x, dy, dx, relu_mask, bias_done=False):
        """Gradients of one convolution.  The data gradient stays on the main stream (it is the critical path of the
        backward pass); the filter / bias gradients only feed the optimiser, so they run on the side stream where
        they overlap the issue-bound LRN/pool gradient kernels of the main chain."""
        s = self.sp[name]
        n = x.shape[0]
        dw = self.var2d("dcnn/%sW" % name, self.grads)
        main = torch.cuda.current_stream()
        ready = torch.cuda.Event()
        ready.record(main)  # dy is complete
        if dx is not None:
            K.conv_dgrad(s, dy, self.sh[name], dx, relu_mask=relu_mask)
        self._side.wait_event(ready)
        with torch.cuda.stream(self._side):
            K.conv_wgrad(s, x, dy, dw)  # split-K chosen by the library
            if not bias_done:
                nv.call("vl_colsum", dy, self.var("dcnn/%sb" % name, self.grads), n * s.p * s.q, s.cout, s.cout)

    def _head_bwd(self, n):
        """From d(loss)/d(logits) down to d(loss)/d(frame features); returns the bf16 feature gradient."""
        cfg, A, G, sh = self.cfg, self.A, self.G, self.sh
        b, c = n // cfg.fpc, cfg.num_classes
        feat_name = "f7" if "fc7" in sh else "f6"
        feat = A[feat_name][:n]
        dfeat = G["d" + feat_name][:n]
        if cfg.workflow == "fc" and cfg.early_fusion:
            dim = feat.shape[1]
            dl_bf = A["dlogits_bf"][:b]
            self._dense_bwd(A["pooled_bf"][:b], dl_bf, "fc_convert_w", "fc_convert_b")
            K.linear_dgrad(dl_bf, sh["fc_convert"], A["dpooled"][:b], n_contract=c)
            nv.call("vl_segment_pool_bwd_relu_bf16", A["dpooled"][:b], feat, cfg.fpc, b, dim, POOL[cfg.fusion], dfeat)
            return dfeat
        if cfg.workflow in ("singleframe", "fc"):
            key, bname = self._frame_classifier()
            wname = bname[:-1] + "W" if key == "fc8" else "fc_convert_w"
            dfl = A["d_frame_logits"][:n]
            nv.call("vl_segment_pool_bwd", A["dlogits"][:b], None, cfg.fpc, b, self.c_pad, POOL[cfg.fusion], dfl)
            nv.call("vl_mul", dfl, None, None, A["d_frame_logits_bf16"][:n], n * self.c_pad)
            dfl_bf = A["d_frame_logits_bf16"][:n]
            self._dense_bwd(feat, dfl_bf, wname, bname)
            K.linear_dgrad(dfl_bf, sh[key], dfeat, relu_mask=feat, n_contract=c)
            return dfeat
        hd, t_len = cfg.lstm_hidden, cfg.fpc
        dl_bf = A["dlogits_bf"][:b]
        state = cfg.fusion == "state"
        fc = "fc_convert" if state else "output_fc"
        if fc in sh:
            self._dense_bwd(self._top_bf, dl_bf, fc + "_w", fc + "_b")
            K.linear_dgrad(dl_bf, sh[fc], A["ddropped"][:b], n_contract=c)
            dtop = A["ddropped"][:b]
        else:
            dtop = A["dlogits"][:b, :c].contiguous()
        if self._dropout_on:
            nv.call("vl_mul", dtop, A["mask"][:b], A["dfused"][:b], None, b * hd)
            dtop = A["dfused"][:b]
        top = cfg.lstm_layers - 1
        nv.call("vl_segment_pool_bwd", dtop, None, t_len, b, hd, POOL["last" if state else cfg.fusion],
                A["dhseq%d" % top][:n])
        for layer in range(top, -1, -1):
            kn = "rnn/multi_rnn_cell/cell_%d/basic_lstm_cell/kernel" % layer
            bn = "rnn/multi_rnn_cell/cell_%d/basic_lstm_cell/bias" % layer
            kern_g = self.var(kn, self.grads)
            d_in = kern_g.shape[0] - hd
            dg = A["dg"][:n]
            if hd == 256:
                kern = self.var(kn)
                nv.call("vl_lstm_bwd_cluster", A["dhseq%d" % layer][:n], A["acts%d" % layer][:n],
                        A["cs%d" % layer][:n], kern[d_in:], dg, b, t_len, hd)
            else:
                nv.call("vl_lstm_bwd", A["dhseq%d" % layer][:n], A["acts%d" % layer][:n], A["cs%d" % layer][:n],
                        sh["lstm%d_wht" % layer], dg, b, t_len, hd)
            x = feat if layer == 0 else A["hseq_bf%d" % (layer - 1)][:n]
            K.linear_wgrad(x, dg, kern_g[:d_in], split_k=self._split_k(d_in, 4 * hd, n))
            K.linear_wgrad(A["hprev_bf%d" % layer][:n], dg, kern_g[d_in:], split_k=self._split_k(hd, 4 * hd, n))
            nv.call("vl_colsum", dg, self.var(bn, self.grads), n, 4 * hd, 4 * hd)
            if layer > 0:
                K.linear_dgrad(dg, sh["lstm%d" % layer][:d_in], A["dhseq%d" % (layer - 1)][:n])
            else:
                K.linear_dgrad(dg, sh["lstm0"][:d_in], dfeat, relu_mask=feat)
        return dfeat

    def _reduce_late_gradients(self):
        """fc6/fc7/LSTM/output gradients (89 % of the bytes) are final once fc6's filter gradient is enqueued: their
        all-reduce (W > 1) runs on NCCL's stream and their squared norms on a side stream while the convolution
        gradients are still being computed."""
        late = self.grads[self.var_off["dcnn/fc6W"]:]
        final = torch.cuda.Event()
        final.record()
        self._norm_stream.wait_event(final)
        with torch.cuda.stream(self._norm_stream):
            if self.world > 1:
                self._early_reduce = parallel.allreduce_async(late, self.group)
                parallel.wait(self._early_reduce)
            nv.call("vl_grad_sqnorms", late, late.numel(), self._seg_late, len(self.var_shapes) - self._k_late,
                    self.sqnorms[self._k_late:], self._ws_late, self._ws_late.numel())
            self._late_norms_ready = torch.cuda.Event()
            self._late_norms_ready.record()

    def _encoder_bwd(self, dfeat, n):
        A, G, sp, sh = self.A, self.G, self.sp, self.sh
        s1, s2, s3 = sp["conv1"], sp["conv2"], sp["conv3"]
        if "fc7" in sh:
            self._dense_bwd(A["f6"][:n], dfeat, "dcnn/fc7W", "dcnn/fc7b")
            K.linear_dgrad(dfeat, sh["fc7"], G["df6"][:n], relu_mask=A["f6"][:n])
        df6 = G["df6"][:n]
        flat = A["p5"][:n].view(n, sp["flat"])
        self._dense_bwd(flat, df6, "dcnn/fc6W", "dcnn/fc6b")
        self._reduce_late_gradients()
        K.linear_dgrad(df6, sh["fc6"], G["dp5"][:n].view(n, sp["flat"]))
        nv.call("vl_maxpool_bwd", G["dp5"][:n], A["arg5"][:n], G["da5"][:n], A["a5"][:n], n, s3.p, s3.q, 256)
        self._conv_bwd("conv5", A["a4"][:n], G["da5"][:n], G["da4"][:n], A["a4"][:n])
        self._conv_bwd("conv4", A["a3"][:n], G["da4"][:n], G["da3"][:n], A["a3"][:n])
        self._conv_bwd("conv3", A["p2"][:n], G["da3"][:n], G["dp2"][:n], None)
        # conv2 / conv1 blocks: the issue-bound LRN/pool gradient kernels and the tensor-bound contractions use
        # different pipes, so the batch is split in two halves that run the chain
        #     pool_lrn_bwd2 -> conv2 dgrad -> pool_lrn_bwd1 -> conv1 wgrad
        # on two streams; the LRN kernel of one half overlaps the contraction of the other.  conv2's filter gradient
        # (needs both halves of da2) stays on the filter-gradient stream.
        main = torch.cuda.current_stream()
        s1s = sp["conv1_s2d"]
        dp2_ready = torch.cuda.Event()
        dp2_ready.record(main)
        # VL_BWD_HALVES=1: two half-batch chains on two streams (paid off while the LRN / pool gradient kernels were
        # issue bound at 0.26 of the HBM rate; measured 6.32 ms against 6.29 ms per step with the current kernels)
        halves = [(0, n // 2), (n // 2, n)] if (n >= 2 and os.environ.get("VL_BWD_HALVES", "0") == "1") else [(0, n)]
        da2_ready = []
        for i, (lo, hi) in enumerate(halves):
            st = main if i == 0 else self._side2
            if st is not main:
                st.wait_event(dp2_ready)
            m = hi - lo
            with torch.cuda.stream(st):
                nv.call("vl_pool_lrn_bwd", A["a2"][lo:hi], G["dp2"][lo:hi], A["arg2"][lo:hi], G["da2"][lo:hi],
                        self.var("dcnn/conv2b", self.grads), m, s2.p, s2.q, 256, LRN["radius"], LRN["alpha"],
                        LRN["beta"], LRN["bias"])
                ev = torch.cuda.Event()
                ev.record(st)
                da2_ready.append(ev)
                K.conv_dgrad_d2s(s2, G["da2"][lo:hi], self.sh["conv2_d2s"], G["dp1"][lo:hi], sh=2, sw=2)
                nv.call("vl_pool_lrn_bwd", A["a1"][lo:hi], G["dp1"][lo:hi], A["arg1"][lo:hi], G["da1"][lo:hi],
                        self.var("dcnn/conv1b", self.grads), m, s1.p, s1.q, 96, LRN["radius"], LRN["alpha"],
                        LRN["beta"], LRN["bias"])
                # row-shift form when an output row fits one k-block (q <= 64): the taps of a filter row share one
                # staged input row (kernels.conv_wgrad_t); split-K atomics: halves add up
                K.conv_wgrad_t(s1s, A["x_s2d"][lo:hi], G["da1"][lo:hi], self.dws1,
                               row_shift=(s1s.q <= 64 and s1s.cin_g <= 64 and os.environ.get("VL_WGRAD_ROW", "1") != "0"),
                               flops=K.conv_flops(s1, m))
        for ev in da2_ready:
            self._side.wait_event(ev)
        with torch.cuda.stream(self._side):
            # swapped operands (dy^T on M): 547 us against 570 us at 1024 frames (tests/bringup/wgrad_probe.py)
            K.conv_wgrad_t(s2, A["p1"][:n], G["da2"][:n], self.var2d("dcnn/conv2W", self.grads))
        if len(halves) > 1:
            main.wait_stream(self._side2)
        self._xs2d_free = torch.cuda.Event()
        self._xs2d_free.record()  # every conv1 filter-gradient launch (the last reader of x_s2d) is enqueued before this
        nv.call("vl_s2d_unpack_grad", self.dws1, self.var("dcnn/conv1W", self.grads), s1.kh, s1.kw, 3, 96, s1.stride)
        torch.cuda.current_stream().wait_stream(self._side)  # join: every filter gradient is in the arena

    # ------------------------------------------------------------------------------------------
    # training step
    # ------------------------------------------------------------------------------------------
    def set_serial(self, serial):
        """serial=True runs every kernel of the step on the caller's stream (no overlap between the backward chains):
        used by bench.py to time each contraction launch in isolation for the roofline."""
        if serial:
            cur = torch.cuda.current_stream()
            self._side, self._side2, self._norm_stream, self._stage_stream = cur, cur, cur, cur
        else:
            self._side, self._side2, self._norm_stream, self._stage_stream = self._streams

    def train_step(self, frames, onehot, lr, dropout_mask=None, apply_update=True, crops=None, global_clips=None,
                   frames_ready=None):
        """One `sess.run([loss, lr, global_step, optimizer])` (run_task.py:44, train.py:199-222).

        frames: host/device frames of `clips * fpc` images; onehot: int32 [clips, C] (utils_.labels_to_one_hot).
        global_clips: clips of the GLOBAL batch over all data-parallel ranks (default: local clips x world); the
        loss is the mean over the global batch (train.py:123) whatever the shard sizes are.
        frames_ready (device-resident frames only): a CUDA event after which the frames are complete (the H2D copy of
        Engine.prefetch), or True when they have been complete all along; the input staging then waits for exactly
        that instead of for everything enqueued on the caller's stream, i.e. it overlaps the optimiser tail of the
        previous step.  None: no assumption.
        Returns (loss, lr, global_step, accuracy, grads_norm) with global_step already incremented."""
        self._frames_ready = frames_ready if isinstance(frames, torch.Tensor) and frames.is_cuda else None
        cfg, A = self.cfg, self.A
        if self.G is None:
            self._alloc_backward()
        self._injected_mask = None
        if dropout_mask is not None:
            self._injected_mask = torch.as_tensor(np.asarray(dropout_mask, dtype=np.float32)).to(self.dev)
        if len(frames) == 0:
            # a data-parallel rank whose share of this batch is empty (fewer videos than ranks): zero gradients, zero
            # loss / correct count, and the same collectives as every other rank
            if self.world <= 1 or not global_clips:
                raise ValueError("empty batch")
            self._global_clips = int(global_clips)
            nv.call("vl_zero", self.grads_ext, self.grads_ext.numel() * 4)
            self._reduce_late_gradients()
            return self._finish_step(lr, apply_update)
        frames, is_u8, n = self._stage_frames(frames)
        if n % cfg.fpc != 0:
            raise ValueError("number of frames (%d) is not a multiple of num_frames_per_clip (%d)" % (n, cfg.fpc))
        b, c = n // cfg.fpc, cfg.num_classes
        if isinstance(onehot, torch.Tensor):
            if tuple(onehot.shape) != (b, c):
                raise ValueError("labels shape %s does not match [%d, %d]" % (tuple(onehot.shape), b, c))
            A["labels"][:b].copy_(onehot.to(torch.int32), non_blocking=True)
        else:
            lab = np.ascontiguousarray(np.asarray(onehot, dtype=np.int32))
            if lab.shape != (b, c):
                raise ValueError("labels shape %s does not match [%d, %d]" % (lab.shape, b, c))
            A["labels"][:b].copy_(torch.from_numpy(lab), non_blocking=True)
        feat = self._encoder_fwd(frames, is_u8, n, True, self._stage_crops(crops, frames, n))
        logits = self._head_fwd(feat, n, True)
        self._global_clips = int(global_clips) if global_clips else b * self.world
        nv.call("vl_softmax_ce", logits, A["labels"][:b], b, c, 1.0 / self._global_clips, A["row_loss"],
                self.scalars[4:6], A["dlogits"][:b], A["dlogits_bf"][:b], self.c_pad)
        self._zero_grads(n)
        # During the backward pass the persistent contraction kernels leave a few KB of shared memory per SM unused:
        # the issue-bound LRN / pool gradient CTAs (2 KB each, main stream) then become resident NEXT TO the filter-
        # gradient CTAs of the side stream instead of queueing behind 148 CTAs that own all the shared memory
        reserve = int(os.environ.get("VL_BWD_SMEM_RESERVE", str(self.bwd_smem_reserve)))
        if reserve:
            nv.lib().vl_set_smem_reserve(reserve)
        try:
            dfeat = self._head_bwd(n)
            self._encoder_bwd(dfeat, n)
        finally:
            if reserve:
                nv.lib().vl_set_smem_reserve(0)
        return self._finish_step(lr, apply_update)

    def _finish_step(self, lr, apply_update):
        """Tail of a train step: the remaining all-reduce, global-norm clip, optimiser, operand refresh, read-back."""
        cfg = self.cfg
        if self.world > 1:
            # ONE tail collective: [header with loss / correct | conv1..conv5 gradients] (contiguous by construction)
            parallel.allreduce_gradients(self.grads_ext[:self._hdr + self.var_off["dcnn/fc6W"]], None, self.group)
        early_n = self.var_off["dcnn/fc6W"]
        nv.call("vl_grad_sqnorms", self.grads[:early_n], early_n, self._seg_early, self._k_late, self.sqnorms,
                self._ws_early, self._ws_early.numel())
        torch.cuda.current_stream().wait_event(self._late_norms_ready)  # implies the early all-reduce has completed
        clip = float(cfg.clip_norm) if cfg.clip_norm else 0.0
        nv.call("vl_clip_scalars", self.sqnorms, len(self.var_shapes), clip, 1.0, self.scalars)
        if os.environ.get("VL_EARLY_READ", "1") != "0":
            self._scalars_ready = torch.cuda.Event()
            self._scalars_ready.record()
        if apply_update:
            fused = False
            if cfg.optimizer == "sgd":
                segs = self._plain_shadow_segments()
                if segs:
                    import ctypes
                    k = len(segs)
                    begin = (ctypes.c_int64 * k)(*[b for b, _, _ in segs])
                    end = (ctypes.c_int64 * k)(*[e for _, e, _ in segs])
                    dst = (ctypes.c_void_p * k)(*[t.data_ptr() for _, _, t in segs])
                    nv.call("vl_sgd_update_shadow", self.params, self.grads, self.arena_n, float(lr), self.scalars, 1.0,
                            k, begin, end, dst)
                    fused = True
                else:
                    nv.call("vl_sgd_update", self.params, self.grads, self.arena_n, float(lr), self.scalars, 1.0)
            else:
                self.adam_t += 1
                nv.call("vl_adam_update", self.params, self.grads, self.adam_m, self.adam_v, self.arena_n, float(lr),
                        0.9, 0.999, 1e-8, self.adam_t, self.scalars, 1.0)
            self.refresh_shadows(plain_done=fused)
            self.global_step += 1
        return self.read_step_scalars(lr)

    def read_step_scalars(self, lr):
        """Device -> host read of the step results (the only synchronisation point of a step)."""
        if self._scalars_ready is not None:
            self._rb_stream.wait_event(self._scalars_ready)
            with torch.cuda.stream(self._rb_stream):
                self._scalars_host.copy_(self.scalars, non_blocking=True)
            self._rb_stream.synchronize()
            self._scalars_ready = None
            s = self._scalars_host.numpy().copy()
        else:
            s = self.scalars.cpu().numpy()
        loss = float(s[4])  # sum over ranks of (1 / global clips) * local row losses = mean over the global batch
        acc = float(s[5]) / self._global_clips
        return loss, float(lr), self.global_step, acc, float(s[2])
