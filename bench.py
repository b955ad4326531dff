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
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return d.get("bf16_tflops_sustained", 1371.0), d.get("hbm_gbs", 6555.2), "measured"
    return 1400.0, 6650.0, "fallback"


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
    torch.cuda.set_device(local_rank)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    cfg = E.EngineConfig(workflow="lrcn", fusion="avg", fpc=FPC, num_classes=NUM_CLASSES, lstm_hidden=256,
                         lstm_layers=1, optimizer="sgd", clip_norm=10, dropout_keep_prob=0.5, mean=MEAN_BGR, seed=1234)
    eng = E.Engine(cfg, max_clips=CLIPS_PER_GPU, device=dev, rank=rank, world=world, group=group)
    rng = np.random.default_rng(rank)
    frames_host = rng.integers(0, 256, size=(CLIPS_PER_GPU * FPC, 227, 227, 3), dtype=np.uint8)
    labels = rng.integers(0, NUM_CLASSES, CLIPS_PER_GPU)
    onehot_host = np.zeros((CLIPS_PER_GPU, NUM_CLASSES), np.int32)
    onehot_host[np.arange(CLIPS_PER_GPU), labels] = 1
    frames_dev = torch.from_numpy(frames_host).to(dev)
    onehot_dev = torch.from_numpy(onehot_host).to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

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
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- device-resident throughput (value) ----
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    l0 = nv.lib().vl_launch_count()
    ms_total = timed(lambda i: eng.train_step(frames_dev, onehot_dev, lr_table_value(eng.global_step)),
                     args.steps, args.warmup)
    launches = (nv.lib().vl_launch_count() - l0) * args.steps // (args.steps + args.warmup)
    clocks = sampler.finish() if sampler else None
    ms_step = ms_total / args.steps
    value = world * CLIPS_PER_GPU / (ms_step * 1e-3)

    # ---- end to end from host buffers: every step copies its uint8 frames + int32 labels from PINNED host memory
    # (H2D inside the timed region, double buffered on a copy stream so that the copy of step i+1 overlaps the
    # compute of step i) and reads the step scalars back (D2H, the step's only synchronisation) ----
    pin_frames = torch.from_numpy(frames_host).pin_memory()
    pin_onehot = torch.from_numpy(onehot_host).pin_memory()

    def e2e_loop(nsteps):
        nxt = eng.prefetch(pin_frames, pin_onehot, 0)
        for i in range(nsteps):
            fd, od, ev, done = nxt
            if i + 1 < nsteps:
                nxt = eng.prefetch(pin_frames, pin_onehot, (i + 1) % 2)
            torch.cuda.current_stream().wait_event(ev)
            eng.train_step(fd, od, lr_table_value(eng.global_step))
            done.record()

    e2e_steps = max(2, args.steps)
    e2e_loop(2)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_loop(e2e_steps)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    ms_e2e /= e2e_steps
    e2e_value = world * CLIPS_PER_GPU / (ms_e2e * 1e-3)

    # ---- roofline of the dominant kernel (umma_gemm_kernel): device time of all its launches in one step ----
    peak_tf, peak_hbm, peak_src = measured_peaks()
    # every rank runs this extra step (the train step all-reduces); rank 0's events are the ones reported
    spans = []
    labels = []
    orig, orig_flat = nv.gemm, nv.conv_flat

    def timed_call(fn):
        def wrapper(*a, **k):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()   # on the stream the kernel is launched on (torch's current stream at call time)
            fn(*a, **k)
            e.record()
            spans.append((s, e))
            d = a[0]
            if hasattr(d, "m"):
                labels.append("gemm a%d b%d m=%d n=%d k=%d g=%d%s" % (d.a_mode, d.b_mode, d.m, d.n, d.k, d.groups,
                                                                     " d2s" if d.d2s_c else ""))
            else:
                labels.append("conv_flat %dx%dx%d k%dx%d cout_g=%d g=%d" % (d.h, d.w, d.c, d.kh, d.kw, d.cout_g, d.groups))
        return wrapper
    nv.gemm, nv.conv_flat = timed_call(orig), timed_call(orig_flat)
    eng.set_serial(True)  # one stream: each launch is timed alone (co-running kernels would stretch the spans)
    try:
        eng.train_step(frames_dev, onehot_dev, lr_table_value(eng.global_step))
        torch.cuda.synchronize()
    finally:
        nv.gemm, nv.conv_flat = orig, orig_flat
        eng.set_serial(False)
    gemm_ms = sum(s.elapsed_time(e) for s, e in spans)
    if rank == 0:  # per-launch list of the contraction kernels (stderr: stdout carries exactly one JSON line)
        for (s_, e_), lab in zip(spans, labels):
            sys.stderr.write("contraction %8.1f us  %s\n" % (s_.elapsed_time(e_) * 1e3, lab))
    n_gemm = len(spans)
    barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    achieved = CLIPS_PER_GPU * FLOP_PER_CLIP_TRAIN / (gemm_ms * 1e-3) / 1e12
    cpu_value, cpu_ms, cores = (None, None, None)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu_value, cpu_ms, cores = cpu_train_clips_per_sec(2, 1)
        cpu = {"value": cpu_value, "unit": "clips/s", "cores": cores, "kind": "port",
               "sample": "%d clips x %d frames train step of the same model (torch-CPU port of the reference's TF "
                         "graph, oracle/lrcn_torch.py), 1 warm-up + 2 timed steps" % (CPU_SAMPLE_CLIPS, FPC)}
    # DRAM traffic of the contraction launches of one step, from the committed ncu capture (bytes per step: the roofline
    # entry aggregates the launches of a step, and so does this number); null when the capture is absent
    traffic, traffic_note = None, None
    tpath = os.path.join(ROOT, "profiles", "contraction_dram_traffic.json")
    if os.path.exists(tpath):
        try:
            with open(tpath) as fh:
                tj = json.load(fh)
            traffic, traffic_note = tj.get("bytes_per_step"), tj.get("source")
        except Exception:
            traffic, traffic_note = None, None
    out = {
        "metric": "clips/sec (16x227x227) LRCN train step", "value": value, "unit": "clips/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "LRCN AlexNet-fc7 + 1-layer LSTM(256) train step, 16-frame 227x227 clips, "
                               "batch 64 clips/GPU (BASELINE.json configs[1])",
                   "clips_per_gpu": CLIPS_PER_GPU, "frames_per_clip": FPC, "optimizer": "sgd", "clip_norm": 10,
                   "dropout_keep_prob": 0.5, "parallelism": "dp%d" % world,
                   "l2_policy": "inputs larger than L2 (158 MB uint8 frames + >2 GB of activations per step)",
                   "tensor_pipe_frac_of_step": value / world * FLOP_PER_CLIP_TRAIN / (peak_tf * 1e12)},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "clips/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(frames_host.nbytes + onehot_host.nbytes), "d2h_bytes_per_step": 32,
                "input": "uint8 frames + int32 one-hot labels in pinned host memory, H2D on a copy stream "
                         "(double buffered) inside the timed region; Engine.prefetch + Engine.train_step"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": achieved / peak_tf, "traffic": traffic, "traffic_note": traffic_note,
                     "peak_source": peak_src + " (sustained bf16)",
                     "kernel": "tcgen05 contraction kernels (umma_gemm_kernel + conv_flat_kernel), all launches of a step",
                     "launches_per_step": n_gemm, "kernel_ms_per_step": gemm_ms,
                     "note": "kernel_ms = sum of the per-launch CUDA-event durations of one extra step run on a single "
                             "stream (each launch alone); in the timed steps the filter gradients overlap other "
                             "kernels on side streams",
                     "share_of_step": gemm_ms / ms_step},
        "cpu_baseline": cpu,
    }
    emit(out)
    if world > 1:
        dist.destroy_process_group()


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
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
