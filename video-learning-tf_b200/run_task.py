"""`python run_task.py <cfg.yml>`: the orchestration loops of the reference (run_task.py:25-152) with the two
`sess.run` calls replaced by Engine.train_step / Engine.forward.  Log lines, nats computation, min-loss tracking,
save cadence and the accuracy file keep the reference's formats.
"""
import argparse
import math
import os
import time

from . import checkpoint
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


def do_train(settings, train, feeder, engine):
    run_batch_count = 0
    min_train_loss = (1000, -1)
    info("Starting train")
    for _ in range(settings.train.epoch_index, settings.train.epochs):
        while feeder.loop():
            frames, onehot, cpvs, num_data, num_labels, padding = feeder.get_feed_dict()
            print_iter_info(settings, feeder, num_data, num_labels, padding)
            run_batch_count += 1
            batch_loss, learning_rate, settings.global_step, acc, gnorm = train.step(frames, onehot, feeder.last_crops)
            if min_train_loss[0] > batch_loss:
                min_train_loss = (batch_loss, settings.global_step)
            nats = batch_loss / math.log(settings.num_classes)
            info("Learning rate %2.8f, global step: %d, batch loss/nats : %2.5f / %2.3f " % (
                learning_rate, settings.global_step, batch_loss, nats))
            info("Dataset global step %d, epoch index %d, batch sizes %s, batch index train %d" % (
                settings.global_step, settings.train.epoch_index + 1, str(feeder.get_batch_sizes()),
                feeder.get_batch_index()))
            if feeder.should_save(run_batch_count):
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
    if run_batch_count > 0 and not feeder.should_save(run_batch_count):
        info("Saving model checkpoint out of turn, since training's finished.")
        progress = "ep_%d_btch_%d_gs_%d" % (1 + settings.train.epoch_index, feeder.get_num_batches(),
                                            settings.global_step)
        checkpoint.save(engine, settings.run_folder, progress, feeder.get_num_batches(), settings.train.epoch_index,
                        feeder.main.num_saves)


def do_test(settings, val, feeder, engine):
    tic = time.time()
    settings.global_step = 0
    while feeder.loop():
        frames, onehot, cpvs, num_data, num_labels, padding = feeder.get_feed_dict()
        print_iter_info(settings, feeder, num_data, num_labels, padding)
        logits = engine.forward(frames, feeder.last_crops)
        val.process_validation_logits(logits, onehot, cpvs, settings.val.clip_fusion_method)
        val.save_validation_logits_chunk()
    val.save_validation_logits_chunk(save_all=True)
    accuracy = val.get_accuracy()
    info("Validation run complete in [%s], accuracy: %2.5f" % (elapsed_str(tic), accuracy))
    if val.validation_logits_save_interval is not None:
        with open(os.path.join(settings.run_folder, "accuracy_" + settings.run_id), "w") as f:
            f.write(str(accuracy))
    return accuracy


def main(init_file, device="cuda:0"):
    settings = Settings()
    feeder = settings.initialize(init_file)
    cfg = settings.engine_config(feeder.main.fpc)
    mean = feeder.main.opts.mean_image
    if feeder.main.opts.data_format == defs.data_format.tfrecord and defs.imgproc.sub_mean not in feeder.main.imgproc:
        mean = None  # dataset_.py:494-495: the mean is subtracted only when imgproc lists sub_mean
    cfg.mean = tuple(mean) if mean is not None else None
    engine = Engine(cfg, max_clips=feeder.max_clips_per_batch(), device=device)
    engine.set_read_resize(getattr(feeder.main, "resize_to", None))  # imgproc raw_resize / resize on the device
    if settings.should_resume():
        prefix = checkpoint.resolve(settings.run_folder, settings.resume_file)
        batch_index, epoch_index, gstep = checkpoint.restore(engine, prefix, is_validation=not settings.train)
        if settings.train:
            settings.train.epoch_index = epoch_index
            settings.global_step = gstep
            feeder.main.fast_forward(batch_index if batch_index < feeder.get_num_batches() else 0)
    result = None
    if settings.train:
        train = Train(settings, feeder, engine)
        do_train(settings, train, feeder, engine)
    elif settings.val:
        val = Validation(settings)
        result = do_test(settings, val, feeder, engine)
    info("Run [%s] complete." % settings.run_id)
    return result


if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("init_file", help="Configuration .yml file for the run.")
    parser.add_argument("--device", default="cuda:0")
    args = parser.parse_args()
    main(args.init_file, args.device)
