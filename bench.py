"""bench.py -- headline benchmark of the LRCN hot path (BASELINE.json: clips/sec, 16x227x227, LRCN train step).

    python bench.py --gpus N --steps K --warmup W              our arm (one process per GPU under torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K ...    reference arm: the CPU port of the reference's graph

Workload (config-2 of BASELINE.json): per GPU 64 clips x 16 frames uint8[1024,227,227,3], LRCN
(AlexNet fc7 -> LSTM(256) -> avg fusion -> output fc), SGD lr 1e-3 decay [exp, interval, 1000, 0.96], clip_norm 10,
dropout 0.5, random-init weights from default_rng(1234), synthetic data.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CLIPS_PER_GPU = 64
FPC = 16
NUM_CLASSES = 101
MEAN_BGR = (99.197148, 105.293620, 109.503945)
FLOP_PER_CLIP_TRAIN = 68.326e9  # BASELINE.md section 3 (fwd + dgrad + wgrad of every GEMM)
FLOP_PER_CLIP_FWD = 23.983e9
CPU_SAMPLE_CLIPS = 8            # bounded sample for the CPU arms


def measured_peaks():
    """(burst bf16 TFLOP/s, sustained bf16 TFLOP/s, HBM GB/s, source).  MEASURED_PEAKS.json is driver-written: cuBLAS
    bf16 8192^3 best-of-10 (burst: a kernel timed alone) and back to back for 4 s (sustained: inside a long step),
    copy bandwidth.  Fallback: the figures B200_PROFILING.md states."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return (d.get("bf16_tflops", 1642.6), d.get("bf16_tflops_sustained", 1371.0), d.get("hbm_gbs", 6555.2),
                "MEASURED_PEAKS.json")
    return 1650.0, 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


def lr_table_value(step, base_lr=0.001, factor=0.96, freq=1000):
    """Train.precompute_learning_rates (train.py:50-109) for lr_decay [exp, interval, 1000, 0.96]."""
    return base_lr * math.pow(factor, step // freq)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        threading.Thread.__init__(self, daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False
        self.proc = None

    def run_nvml(self):
        """In-process NVML polling (same counters as the nvidia-smi query, without a second process taking the driver's
        locks every 100 ms next to the timed region).  Returns False when NVML is not usable."""
        try:
            import pynvml as N
            N.nvmlInit()
            h = N.nvmlDeviceGetHandleByIndex(self.index)
            mx = N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)
            get_reasons = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                N.nvmlDeviceGetCurrentClocksThrottleReasons
            bits = [(0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"),
                    (0x4, "sw_power_cap")]
            N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)
        except Exception:
            return False
        while not self.stop_flag:
            try:
                sm = N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)
                r = int(get_reasons(h))
                self.samples.append([str(sm), str(mx), "0"] + ["Active" if r & b else "Not Active" for b, _ in bits])
            except Exception:
                pass
            time.sleep(0.02)
        return True

    def run(self):
        if os.environ.get("VL_BENCH_CLOCKS", "nvml") == "nvml" and self.run_nvml():
            return
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                parts = [x.strip() for x in line.split(",")]
                if len(parts) >= 7:
                    self.samples.append(parts)
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for p in self.samples:
            try:
                sm.append(float(p[0]))
                mx = max(mx, float(p[1]))
            except ValueError:
                continue
            for name, val in zip(names, p[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        load = [x for x in sm if x > 0.5 * (sm[-1] if sm else 0)]
        med = load[len(load) // 2] if load else (sm[len(sm) // 2] if sm else None)
        return {"sm_mhz": med, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's TF graph (the reference itself needs TensorFlow 1.x, absent here)
# ----------------------------------------------------------------------------------------------------------
def cpu_train_clips_per_sec(steps, warmup, clips=CPU_SAMPLE_CLIPS):
    import numpy as np
    import torch
    from oracle import lrcn_numpy as O
    from oracle import lrcn_torch as T

    torch.set_num_threads(os.cpu_count() or 1)
    params = T.to_torch(O.init_params(1234, NUM_CLASSES, "fc7", 256, 1), requires_grad=True)
    rng = np.random.default_rng(0)
    frames = torch.tensor(rng.integers(0, 256, size=(clips * FPC, 227, 227, 3)).astype(np.float32)
                          - np.array(MEAN_BGR, np.float32))
    onehot = torch.zeros(clips, NUM_CLASSES, dtype=torch.int32)
    onehot[torch.arange(clips), torch.tensor(rng.integers(0, NUM_CLASSES, clips))] = 1
    mask = (torch.rand(clips, 256) < 0.5).float() * 2.0
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        T.train_step(params, frames, onehot, FPC, lr_table_value(i), "lrcn", "avg", "fc7", clip_norm=10,
                     dropout_mask=mask)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return clips * len(times) / total, total / len(times) * 1e3, torch.get_num_threads()


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # other ranks exit 0 without work
    steps = max(1, min(args.steps, 3))
    warm = max(1, min(args.warmup, 1))
    value, ms, cores = cpu_train_clips_per_sec(steps, warm)
    sample = "%d clips x %d frames (bounded sample of the %d-clip/GPU workload), %d timed steps" % (
        CPU_SAMPLE_CLIPS, FPC, CLIPS_PER_GPU, steps)
    out = {
        "impl": "reference", "metric": "clips/sec (16x227x227) LRCN train step", "value": value,
        "unit": "clips/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "LRCN AlexNet-fc7 + LSTM(256) train step, 16-frame 227x227 clips, CPU port of the "
                               "reference's TensorFlow graph (TensorFlow 1.x cannot be installed here)",
                   "clips_per_step": CPU_SAMPLE_CLIPS},
        "cpu_baseline": {"value": value, "unit": "clips/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(out)


# ----------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------
def hbm_bytes(name, a):
    """ALGORITHMIC bytes (compulsory reads + writes) of one launch of an HBM-bound kernel, from the
    arguments of its C-ABI call (include/vlb200.h).  None: not an HBM-roofline kernel."""
    def numel(t):
        return int(t.numel())

    if name == "vl_frames_s2d_crop":  # frames, is_u8, mean, out, n, hr, wr, crops, h, w, ...
        return numel(a[0]) * a[0].element_size() + numel(a[3]) * 2
    if name in ("vl_lrn_pool_fwd", "vl_maxpool_fwd"):  # x, y, arg, n, h, w, c
        n, h, w, c = a[3], a[4], a[5], a[6]
        pq = ((h - 3) // 2 + 1) * ((w - 3) // 2 + 1)
        return n * c * (h * w * 2 + pq * 2 + pq)
    if name == "vl_pool_lrn_bwd":  # x, dy, arg, dx, dbias, n, h, w, c
        n, h, w, c = a[5], a[6], a[7], a[8]
        pq = ((h - 3) // 2 + 1) * ((w - 3) // 2 + 1)
        return n * c * (h * w * 2 + pq * 2 + pq + h * w * 2)
    if name == "vl_maxpool_bwd":  # dy, arg, dx, relu_of, n, h, w, c
        n, h, w, c = a[4], a[5], a[6], a[7]
        pq = ((h - 3) // 2 + 1) * ((w - 3) // 2 + 1)
        return n * c * (pq * 2 + pq + h * w * 2 + (h * w * 2 if a[3] is not None else 0))
    if name == "vl_colsum":  # dy, out, rows, c, ld
        return a[2] * a[3] * 2
    if name == "vl_grad_sqnorms":
        return a[1] * 4
    if name in ("vl_sgd_update", "vl_sgd_update_shadow"):
        return a[2] * 12
    if name == "vl_adam_update":
        return a[4] * 28
    if name == "vl_gather_bf16":
        return a[3] * 6
    if name == "vl_cast_f32_to_bf16":
        return a[2] * 6
    if name == "vl_zero":
        return a[1]
    return None


class LaunchTracer(object):
    """Brackets every launch that goes through the C-ABI binding with CUDA events on the launching stream (torch's
    current stream at call time) and keeps what the roofline needs: algorithmic FLOPs of the contractions (declared by
    the launch wrappers of kernels.py from the real extents), algorithmic bytes of the HBM-bound kernels."""

    def __init__(self, torch):
        self.torch = torch
        self.rows = []

    def begin(self, name, args, meta):
        e0 = self.torch.cuda.Event(enable_timing=True)
        e0.record()
        row = {"name": name, "e0": e0}
        if name in ("vl_gemm", "vl_conv_flat"):
            row["kind"] = "tensor"
            row["label"], row["flops"] = meta if meta else (name, 0.0)
        else:
            b = hbm_bytes(name, args)
            row["kind"] = "hbm" if b is not None else "other"
            row["label"] = name
            row["bytes"] = b
            if name in ("vl_lrn_pool_fwd", "vl_pool_lrn_bwd", "vl_maxpool_fwd", "vl_maxpool_bwd"):
                k = 3 if name.endswith("fwd") else (5 if name == "vl_pool_lrn_bwd" else 4)
                row["label"] = "%s %dx%dx%d" % (name, args[k + 1], args[k + 2], args[k + 3])
        self.rows.append(row)

    def end(self):
        e1 = self.torch.cuda.Event(enable_timing=True)
        e1.record()
        self.rows[-1]["e1"] = e1

    def table(self, peak_tf, peak_hbm):
        out = []
        for r in self.rows:
            us = r["e0"].elapsed_time(r["e1"]) * 1e3
            row = {"kernel": r["label"], "us": round(us, 1), "bound": r["kind"]}
            if r["kind"] == "tensor" and us > 0:
                tf = r["flops"] / (us * 1e-6) / 1e12
                row.update(gflop=round(r["flops"] / 1e9, 2), tflops=round(tf, 1), frac=round(tf / peak_tf, 3))
            elif r["kind"] == "hbm" and us > 0:
                gbs = r["bytes"] / (us * 1e-6) / 1e9
                row.update(mbytes=round(r["bytes"] / 1e6, 2), gbs=round(gbs, 1), frac=round(gbs / peak_hbm, 3))
            out.append(row)
        return out


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import vlb200  # noqa: F401
    from vlb200 import _native as nv
    from vlb200 import engine as E

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d does not match WORLD_SIZE %d" % (args.gpus, world))
    _pin_rank_to_local_cpus(local_rank, world)
    torch.cuda.set_device(local_rank)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    forward_only = args.mode == "forward"
    clips = args.clips if args.clips else CLIPS_PER_GPU
    flop_per_clip = FLOP_PER_CLIP_FWD if forward_only else FLOP_PER_CLIP_TRAIN

    cfg = E.EngineConfig(workflow="lrcn", fusion="avg", fpc=FPC, num_classes=NUM_CLASSES, lstm_hidden=256,
                         lstm_layers=1, optimizer="sgd", clip_norm=10, dropout_keep_prob=0.5, mean=MEAN_BGR, seed=1234)
    eng = E.Engine(cfg, max_clips=clips, device=dev, rank=rank, world=world, group=group)
    rng = np.random.default_rng(rank)
    frames_host = rng.integers(0, 256, size=(clips * FPC, 227, 227, 3), dtype=np.uint8)
    labels = rng.integers(0, NUM_CLASSES, clips)
    onehot_host = np.zeros((clips, NUM_CLASSES), np.int32)
    onehot_host[np.arange(clips), labels] = 1
    frames_dev = torch.from_numpy(frames_host).to(dev)
    onehot_dev = torch.from_numpy(onehot_host).to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    # frames_ready=True: the device-resident inputs are complete before the timed region starts (the engine's input
    # staging of step i+1 may then run next to the optimiser tail of step i)
    if forward_only:
        def step(i):
            eng.forward_device(frames_dev, training=False, frames_ready=True)
    else:
        def step(i):
            eng.train_step(frames_dev, onehot_dev, lr_table_value(eng.global_step), frames_ready=True)

    # ---- device-resident throughput (value): EXACTLY --steps timed steps ----
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    l0 = nv.lib().vl_launch_count()
    ms_total = timed(step, args.steps, args.warmup)
    launches = (nv.lib().vl_launch_count() - l0) * args.steps // (args.steps + args.warmup)
    ms_step = ms_total / args.steps
    value = world * clips / (ms_step * 1e-3)
    # ---- the same loop for >= 2 s: the region the SUSTAINED tensor peak may be compared with (a 0.13 s region runs
    # at burst clocks) ----
    sus_steps = max(args.steps, int(math.ceil(2000.0 / ms_step)))
    ms_sus = timed(step, sus_steps, 0) / sus_steps
    clocks = sampler.finish() if sampler else None

    # ---- end to end from host buffers: every step copies its uint8 frames (+ int32 labels) from PINNED host memory
    # (H2D inside the timed region, double buffered on a copy stream so that the copy of step i+1 overlaps the
    # compute of step i) and reads the step result back (D2H: the step scalars / the logits) ----
    # VL_BENCH_WC=1: write-combined pinned staging buffers (vl_host_alloc) instead of torch's cached pinned memory
    if os.environ.get("VL_BENCH_WC", "0") == "1":
        pin_frames = nv.host_staging_tensor(frames_host.shape, torch.uint8, write_combined=True)
        pin_frames.copy_(torch.from_numpy(frames_host))
        assert pin_frames.is_pinned()
    else:
        pin_frames = torch.from_numpy(frames_host).pin_memory()
    pin_onehot = torch.from_numpy(onehot_host).pin_memory()
    logits_host = torch.empty(clips, NUM_CLASSES, dtype=torch.float32).pin_memory()

    SLOTS = 3  # batches i+1 and i+2 are in flight while batch i computes: the copy engine never waits for a slot

    def e2e_loop(nsteps):
        queue = [eng.prefetch(pin_frames, pin_onehot, j % SLOTS) for j in range(min(SLOTS - 1, nsteps))]
        for i in range(nsteps):
            fd, od, ev, done = queue.pop(0)
            if i + SLOTS - 1 < nsteps:
                queue.append(eng.prefetch(pin_frames, pin_onehot, (i + SLOTS - 1) % SLOTS))
            torch.cuda.current_stream().wait_event(ev)  # (labels; the frames are handed over through frames_ready)
            if forward_only:
                logits_host.copy_(eng.forward_device(fd, training=False, frames_ready=ev), non_blocking=True)
                done.record()
                torch.cuda.current_stream().synchronize()  # the caller holds this step's logits before the next one
            else:
                eng.train_step(fd, od, lr_table_value(eng.global_step), frames_ready=ev)
                done.record()

    e2e_steps = max(2, args.steps)
    e2e_loop(3)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_loop(e2e_steps)
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1)) / e2e_steps
    e2e_value = world * clips / (ms_e2e * 1e-3)
    # host -> device bandwidth of this rank while all ranks copy at once (the 8-GPU end-to-end limiter)
    barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for _ in range(4):
        frames_dev.copy_(pin_frames, non_blocking=True)
    c1.record()
    barrier()
    h2d_gbs = 4 * frames_host.nbytes / (max_over_ranks(c0.elapsed_time(c1)) * 1e-3) / 1e9

    # ---- roofline: ONE extra step on a single stream, every launch bracketed by CUDA events on its stream (each
    # launch alone: co-running kernels would stretch the spans).  Isolated launches are divided by the BURST peak. ----
    peak_burst, peak_sus, peak_hbm, peak_src = measured_peaks()
    tracer = LaunchTracer(torch)
    eng.set_serial(True)
    nv.tracer = tracer
    try:
        step(0)
        torch.cuda.synchronize()
    finally:
        nv.tracer = None
        eng.set_serial(False)
    rows = tracer.table(peak_burst, peak_hbm)
    barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    tens = [r for r in rows if r["bound"] == "tensor"]
    hbm = [r for r in rows if r["bound"] == "hbm"]
    gemm_ms = sum(r["us"] for r in tens) * 1e-3
    gemm_flops = sum(r["gflop"] for r in tens) * 1e9
    achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12
    hbm_ms = sum(r["us"] for r in hbm) * 1e-3
    hbm_bytes_total = sum(r["mbytes"] for r in hbm) * 1e6
    for r in rows:  # stderr: stdout carries exactly one JSON line
        sys.stderr.write("launch %8.1f us  %-6s %s\n" % (r["us"], r["bound"], r["kernel"]))
    cpu = None
    if world == 1 and not args.no_cpu_baseline and not forward_only:
        cpu_value, cpu_ms, cores = cpu_train_clips_per_sec(2, 1)
        cpu = {"value": cpu_value, "unit": "clips/s", "cores": cores, "kind": "port",
               "sample": "%d clips x %d frames train step of the same model (torch-CPU port of the reference's TF "
                         "graph, oracle/lrcn_torch.py), 1 warm-up + 2 timed steps" % (CPU_SAMPLE_CLIPS, FPC)}
    # DRAM traffic of the contraction launches of one step from the committed ncu capture of the same command
    # (bytes per step); null when the capture is absent.  A profiler cannot run inside the timed program.
    traffic, traffic_note = None, None
    tpath = os.path.join(ROOT, "profiles", "contraction_dram_traffic.json")
    if os.path.exists(tpath) and not forward_only:
        try:
            with open(tpath) as fh:
                tj = json.load(fh)
            traffic, traffic_note = tj.get("bytes_per_step"), tj.get("source")
        except Exception:
            traffic, traffic_note = None, None
    what = "forward pass" if forward_only else "train step"
    sus_value = world * clips / (ms_sus * 1e-3)
    out = {
        "metric": "clips/sec (16x227x227) LRCN %s" % what, "value": value, "unit": "clips/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "LRCN AlexNet-fc7 + 1-layer LSTM(256) %s, 16-frame 227x227 clips, batch %d clips/GPU "
                               "(BASELINE.json configs[%d])" % (what, clips, 4 if forward_only else 1),
                   "mode": args.mode, "clips_per_gpu": clips, "frames_per_clip": FPC, "optimizer": "sgd",
                   "clip_norm": 10, "dropout_keep_prob": 0.5, "parallelism": "dp%d" % world,
                   "l2_policy": "inputs larger than L2 (158 MB uint8 frames + >2 GB of activations per step)"},
        "clocks": clocks,
        "sustained": {"steps": sus_steps, "seconds": sus_steps * ms_sus * 1e-3, "ms_per_step": ms_sus,
                      "value": sus_value,
                      "tensor_frac_of_sustained_peak": sus_value / world * flop_per_clip / (peak_sus * 1e12),
                      "note": "the same loop timed for >= 2 s (device-resident); this is the step-level figure that "
                              "may be held against the sustained cuBLAS peak"},
        "e2e": {"value": e2e_value, "unit": "clips/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(frames_host.nbytes + (0 if forward_only else onehot_host.nbytes)),
                "d2h_bytes_per_step": int(logits_host.numel() * 4) if forward_only else 32,
                "h2d_gbs_per_rank_all_ranks_copying": h2d_gbs,
                "host_staging": "write-combined pinned" if os.environ.get("VL_BENCH_WC", "0") == "1" else "pinned",
                "input": "uint8 frames + int32 one-hot labels in pinned host memory, H2D on a copy stream "
                         "(three device slots: two copies in flight) inside the timed region; Engine.prefetch + Engine.%s" % (
                             "forward_device" if forward_only else "train_step")},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_burst, "unit": "TFLOP/s",
                     "frac": achieved / peak_burst, "traffic": traffic, "traffic_note": traffic_note,
                     "peak_source": peak_src + ": bf16_tflops (burst) -- every launch is timed ALONE, so the burst "
                                               "figure is the denominator",
                     "kernel": "tcgen05 contraction kernels (umma_gemm_kernel + conv_flat_kernel), all launches of a step",
                     "launches_per_step": len(tens), "kernel_ms_per_step": gemm_ms,
                     "algorithmic_gflop_per_step": gemm_flops / 1e9,
                     "step_tensor_frac_of_burst_peak": value / world * flop_per_clip / (peak_burst * 1e12),
                     "share_of_step": gemm_ms / ms_step,
                     "note": "achieved = sum of the algorithmic FLOPs of the contraction launches of ONE extra step / sum "
                             "of their CUDA-event durations, that step run on a single stream (each launch alone); in the "
                             "timed steps the filter gradients overlap other kernels on side streams",
                     "launches": tens,
                     "hbm": {"bound": "hbm", "peak": peak_hbm, "unit": "GB/s",
                             "achieved": hbm_bytes_total / (hbm_ms * 1e-3) / 1e9 if hbm_ms > 0 else None,
                             "frac": hbm_bytes_total / (hbm_ms * 1e-3) / 1e9 / peak_hbm if hbm_ms > 0 else None,
                             "kernel_ms_per_step": hbm_ms, "share_of_step": hbm_ms / ms_step,
                             "note": "HBM-bound kernels of the same serial step: algorithmic bytes (compulsory reads + "
                                     "writes) / CUDA-event duration against the measured copy bandwidth",
                             "launches": [r for r in hbm if r["us"] >= 5.0]},
                     "other_ms_per_step": sum(r["us"] for r in rows if r["bound"] == "other") * 1e-3},
        "cpu_baseline": cpu,
    }
    emit(out)
    if world > 1:
        dist.destroy_process_group()


def _pin_rank_to_local_cpus(local_rank, world):
    """One process per GPU: give every rank its own contiguous slice of the host cores (the pinned staging buffers are
    then allocated and touched NUMA-locally: with 8 ranks copying 158 MB per step the host memory system, not the GPUs,
    bounds the end-to-end number).  No effect when the affinity cannot be changed."""
    if world <= 1 or os.environ.get("VL_BENCH_AFFINITY", "1") == "0":
        return
    try:
        cpus = sorted(os.sched_getaffinity(0))
        per = max(1, len(cpus) // world)
        mine = cpus[local_rank * per:(local_rank + 1) * per]
        if mine:
            os.sched_setaffinity(0, mine)
    except Exception:
        pass


_JSON_FD = None


def _reserve_stdout():
    """stdout must carry exactly ONE JSON line: libraries (NCCL prints its version banner to stdout when NCCL_DEBUG is
    set on the box) are redirected to stderr, the JSON line is written to the original stdout at the end."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, line)


def main():
    _reserve_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mode", default="train", choices=["train", "forward"],
                    help="train: BASELINE configs[1] (the headline); forward: configs[4], forward-only inference")
    ap.add_argument("--clips", type=int, default=0, help="clips per GPU and step (default 64; the sweep uses 16..1024)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
