"""`python run_task.py <cfg.yml>`: the orchestration loops of the reference (run_task.py:25-152) with the two
`sess.run` calls replaced by Engine.train_step / Engine.forward.  Log lines, nats computation, min-loss tracking,
save cadence and the accuracy file keep the reference's formats.
"""
import argparse
import math
import os
import time

from . import checkpoint, parallel
from .defs import defs
from .engine import Engine
from .settings import Settings
from .train import Train
from .utils import elapsed_str, info
from .val import Validation


def print_iter_info(settings, feeder, num_images, num_labels, padding):
    ds = feeder.main
    padinfo = "(%d padding)" % padding if padding > 0 else ""
    epoch_str = "" if settings.val and not settings.train else "epoch: %2d/%2d," % (
        settings.train.epoch_index + 1, settings.train.epochs)
    info("Mode: [%s], %s batch %4d / %4d : %s images%s, %3d labels" % (
        settings.phase, epoch_str, ds.batch_index, len(ds.batches), str(num_images), padinfo, num_labels))


def do_train(settings, train, feeder, engine, dp=None):
    run_batch_count = 0
    min_train_loss = (1000, -1)
    info("Starting train")
    for _ in range(settings.train.epoch_index, settings.train.epochs):
        while feeder.loop():
            frames, onehot, cpvs, num_data, num_labels, padding = feeder.get_feed_dict()
            print_iter_info(settings, feeder, num_data, num_labels, padding)
            run_batch_count += 1
            crops, global_clips = feeder.last_crops, None
            if dp is not None:
                global_clips = len(onehot)
                frames, onehot, crops = dp.shard_batch(frames, onehot, cpvs, crops, feeder.main.fpc)
            batch_loss, learning_rate, settings.global_step, acc, gnorm = train.step(frames, onehot, crops,
                                                                                     global_clips=global_clips)
            if min_train_loss[0] > batch_loss:
                min_train_loss = (batch_loss, settings.global_step)
            nats = batch_loss / math.log(settings.num_classes)
            info("Learning rate %2.8f, global step: %d, batch loss/nats : %2.5f / %2.3f " % (
                learning_rate, settings.global_step, batch_loss, nats))
            info("Dataset global step %d, epoch index %d, batch sizes %s, batch index train %d" % (
                settings.global_step, settings.train.epoch_index + 1, str(feeder.get_batch_sizes()),
                feeder.get_batch_index()))
            if feeder.should_save(run_batch_count) and (dp is None or dp.rank == 0):
                progress = "ep_%d_btch_%d_gs_%d" % (1 + settings.train.epoch_index, feeder.get_batch_index(),
                                                    settings.global_step)
                checkpoint.save(engine, settings.run_folder, progress, feeder.get_batch_index(),
                                settings.train.epoch_index, feeder.main.num_saves)
        if run_batch_count > 0:
            info("Epoch [%d] training run complete." % (1 + settings.train.epoch_index))
        else:
            info("Resumed epoch [%d] is already complete." % (1 + settings.train.epoch_index))
        settings.train.epoch_index += 1
        feeder.rewind_datasets()
    info("Minimum training loss: %2.2f on global index %d" % (min_train_loss[0], min_train_loss[1]))
    if run_batch_count > 0 and not feeder.should_save(run_batch_count) and (dp is None or dp.rank == 0):
        info("Saving model checkpoint out of turn, since training's finished.")
        progress = "ep_%d_btch_%d_gs_%d" % (1 + settings.train.epoch_index, feeder.get_num_batches(),
                                            settings.global_step)
        checkpoint.save(engine, settings.run_folder, progress, feeder.get_num_batches(), settings.train.epoch_index,
                        feeder.main.num_saves)


def do_test(settings, val, feeder, engine, dp=None):
    tic = time.time()
    settings.global_step = 0
    while feeder.loop():
        frames, onehot, cpvs, num_data, num_labels, padding = feeder.get_feed_dict()
        print_iter_info(settings, feeder, num_data, num_labels, padding)
        if dp is None:
            logits = engine.forward(frames, feeder.last_crops)
        else:
            # every rank forwards its contiguous share of the batch's videos; the clip logits are gathered in rank
            # order = the order of the batch, and every rank runs the (host, integer-exact) fusion on all of them
            my_frames, _, my_crops = dp.shard_batch(frames, onehot, cpvs, feeder.last_crops, feeder.main.fpc)
            logits = dp.gather_logits(engine, my_frames, my_crops)
        val.process_validation_logits(logits, onehot, cpvs, settings.val.clip_fusion_method)
        if dp is None or dp.rank == 0:
            val.save_validation_logits_chunk()
    if dp is None or dp.rank == 0:
        val.save_validation_logits_chunk(save_all=True)
    accuracy = val.get_accuracy()
    info("Validation run complete in [%s], accuracy: %2.5f" % (elapsed_str(tic), accuracy))
    if val.validation_logits_save_interval is not None and (dp is None or dp.rank == 0):
        with open(os.path.join(settings.run_folder, "accuracy_" + settings.run_id), "w") as f:
            f.write(str(accuracy))
    return accuracy


class DataParallel(object):
    """`torchrun --nproc-per-node N run_task.py cfg.yml`: one process per GPU.  Every rank runs the same
    feeder (same seeds, same batches) and keeps the contiguous share of each batch's VIDEOS that parallel.shard_range
    gives it, so clips of one video never straddle ranks; gradients are summed by the engine's all-reduce (the loss is
    the mean over the GLOBAL batch whatever the shard sizes are), validation logits are gathered in rank order."""

    def __init__(self, rank, world, group=None):
        self.rank, self.world, self.group = rank, world, group

    def shard_batch(self, frames, onehot, cpvs, crops, fpc):
        lo_v, hi_v = parallel.shard_range(len(cpvs), self.rank, self.world)
        clip0 = int(sum(cpvs[:lo_v]))
        clip1 = clip0 + int(sum(cpvs[lo_v:hi_v]))
        f0, f1 = clip0 * fpc, clip1 * fpc
        return frames[f0:f1], onehot[clip0:clip1], (None if crops is None else crops[f0:f1])

    def gather_logits(self, engine, frames, crops):
        import torch
        if len(frames):
            local = engine.forward_device(frames, training=False, crops=crops)
        else:
            local = torch.zeros(0, engine.cfg.num_classes, dtype=torch.float32, device=engine.dev)
        return parallel.gather_logits(local, self.group).cpu().numpy()


def init_data_parallel(device):
    """(device, DataParallel or None) from the torchrun environment (RANK / WORLD_SIZE / LOCAL_RANK)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return device, None
    import torch
    import torch.distributed as dist
    rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.device_count() >= world:
        device = "cuda:%d" % local
        torch.cuda.set_device(local)
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    elif not dist.is_initialized():  # fewer GPUs than ranks (tests on a one-GPU box): share the device, exchange by gloo
        dist.init_process_group("gloo")
    return device, DataParallel(rank, world)


def main(init_file, device="cuda:0", engine_factory=None):
    """engine_factory(cfg, max_clips, device, rank, world) -> engine: the CUDA Engine by default (there is no CPU
    fallback); the CPU tests of the host workflow inject a stand-in."""
    device, dp = init_data_parallel(device)
    settings = Settings()
    feeder = settings.initialize(init_file)
    cfg = settings.engine_config(feeder.main.fpc, feeder.main.opts.image_shape)
    mean = feeder.main.opts.mean_image
    if feeder.main.opts.data_format == defs.data_format.tfrecord and defs.imgproc.sub_mean not in feeder.main.imgproc:
        mean = None  # dataset_.py:494-495: the mean is subtracted only when imgproc lists sub_mean
    cfg.mean = tuple(mean) if mean is not None else None
    rank, world = (dp.rank, dp.world) if dp is not None else (0, 1)
    if engine_factory is None:
        engine = Engine(cfg, max_clips=feeder.max_clips_per_batch(), device=device, rank=rank, world=world)
    else:
        engine = engine_factory(cfg, feeder.max_clips_per_batch(), device, rank, world)
    engine.set_read_resize(getattr(feeder.main, "resize_to", None))  # imgproc raw_resize / resize on the device
    net = settings.pipelines[settings.pipeline_names[0]]
    if net.weights_file is not None:
        # dcnn.create(..., weightsFile) (alexnet.py:50-52,69-71): conv1..fc7 start from bvlc_alexnet.npy; a checkpoint
        # restored below overrides them, exactly as saver.restore does after the graph was built from the file
        checkpoint.load_alexnet_npy(engine, net.weights_file)
    if settings.should_resume():
        prefix = checkpoint.resolve(settings.run_folder, settings.resume_file)
        batch_index, epoch_index, gstep = checkpoint.restore(engine, prefix, is_validation=not settings.train)
        if settings.train:
            settings.train.epoch_index = epoch_index
            settings.global_step = gstep
            # like dset.restore(idx, epoch) (feeder.py:143-194): a snapshot taken at the end of an epoch has
            # batch_index == num_batches, the loop finds nothing left, logs "Resumed epoch is already complete"
            # and moves on to the next epoch with the global step (= LR table index) where it belongs
            feeder.main.fast_forward(batch_index)
    result = None
    if settings.train:
        train = Train(settings, feeder, engine)
        do_train(settings, train, feeder, engine, dp)
    elif settings.val:
        val = Validation(settings, use_device=engine_factory is None)
        result = do_test(settings, val, feeder, engine, dp)
    info("Run [%s] complete." % settings.run_id)
    return result


if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("init_file", help="Configuration .yml file for the run.")
    parser.add_argument("--device", default="cuda:0")
    args = parser.parse_args()
    main(args.init_file, args.device)
