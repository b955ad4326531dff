"""Worker of tests/test_gpu_multi.py: a 2-rank NCCL train step must equal the 1-rank step on the concatenated batch
(SURVEY 8e parity rule: loss / parameters within fp tolerance, reduction order differs)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import vlb200  # noqa: E402,F401
from vlb200 import engine as E  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
if torch.cuda.device_count() >= world:
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
else:
    # a box with fewer GPUs than ranks (the driver's 1-GPU box): both ranks share cuda:0 and exchange through gloo
    # (NCCL refuses two ranks on one device); the engine's data-parallel logic is the same, only the transport differs
    local = 0
    torch.cuda.set_device(0)
    dist.init_process_group("gloo")
print("dp_worker rank %d/%d on cuda:%d, backend %s" % (rank, world, local, dist.get_backend()))
fpc, clips_per_rank, classes = 2, 2, 101
cfg = E.EngineConfig(workflow="lrcn", fusion="avg", fpc=fpc, num_classes=classes, lstm_hidden=256, clip_norm=10,
                     dropout_keep_prob=0.0, optimizer="sgd")
params = E.init_variables(cfg, seed=5)
rng = np.random.default_rng(6)
total = clips_per_rank * world
frames = rng.uniform(-60, 60, size=(total * fpc, 227, 227, 3)).astype(np.float32)
onehot = np.zeros((total, classes), np.int32)
onehot[np.arange(total), rng.integers(0, classes, total)] = 1

eng = E.Engine(cfg, max_clips=clips_per_rank, device="cuda:%d" % local, params=params, rank=rank, world=world)
lo, hi = rank * clips_per_rank, (rank + 1) * clips_per_rank
loss, lr, step, acc, gnorm = eng.train_step(frames[lo * fpc:hi * fpc], onehot[lo:hi], 1e-2)
sd = eng.state_dict()
ok = True
if rank == 0:
    ref = E.Engine(cfg, max_clips=total, device="cuda:%d" % local, params=params)
    rloss, _, _, racc, rgnorm = ref.train_step(frames, onehot, 1e-2)
    rsd = ref.state_dict()
    worst = 0.0
    for k, v in rsd.items():
        if k == "global_step":
            continue
        delta = np.abs(sd[k] - v).max()
        upd = np.abs(v - params[k]).max() + 1e-12  # compare the UPDATE (lr * clipped gradient), not the weights
        worst = max(worst, float(delta / upd))
    print("dp2 vs dp1: loss %.6f vs %.6f, acc %.3f vs %.3f, grads_norm %.5f vs %.5f, worst update rel err %.3e" % (
        loss, rloss, acc, racc, gnorm, rgnorm, worst))
    ok = (abs(loss - rloss) < 1e-4 * max(1.0, abs(rloss)) and abs(acc - racc) < 1e-6 and
          abs(gnorm - rgnorm) < 2e-2 * rgnorm and worst < 5e-2)
# the verdict and the parameter digests travel on the backend's native device (gloo: host tensors; nccl: device tensors)
xdev = "cpu" if dist.get_backend() == "gloo" else "cuda:%d" % local
torch.cuda.synchronize()
flag = torch.tensor([1 if ok else 0], device=xdev)
dist.broadcast(flag, 0)
# both ranks must hold identical parameters after the step
per_var = [float(np.float64(v).sum()) for k, v in sorted(sd.items()) if k != "global_step"]
digest = torch.tensor(per_var, device=xdev, dtype=torch.float64)
both = torch.cat([digest, -digest])
dist.all_reduce(both, op=dist.ReduceOp.MAX)  # max(d) == -max(-d) = min(d) on every variable <=> all ranks agree
k = digest.numel()
same = bool(torch.equal(both[:k], -both[k:]))
if not same or flag.item() != 1:
    print("rank %d: verdict %d, parameters identical across ranks: %s\n  max %s\n  min %s" % (
        rank, int(flag.item()), same, both[:k].tolist(), (-both[k:]).tolist()), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if (flag.item() == 1 and same) else 1)
