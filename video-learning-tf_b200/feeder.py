"""Batch feeding contract of the hot path (reference: feeder.py:84-141, dataset_.py:386-420,562-613).

The reference's TFRecord reader and per-image Python preprocessing are outside the hot path ("next").
What the device path depends on is the integer contract, which this module keeps: frames ordered
video -> clip -> frame, ONE label row per clip, whole videos per batch, `padding` always 0.  Two TF-free sources:
`synthetic` (seeded uint8 frames, UCF101-shaped) and `npy` (a .npz with `frames` uint8 [N,H,W,3], `labels`,
`clips_per_video`).
"""
import math

import numpy as np

from .defs import defs
from .utils import error, info, labels_to_one_hot


class Dataset(object):
    def __init__(self, opts, batch_size, num_classes, epochs, save_freq_per_epoch):
        self.opts = opts
        self.batch_size = batch_size  # in VIDEOS (items), like dataset_.py:582-613
        self.num_classes = num_classes
        self.fpc = opts.num_frames_per_clip
        self.input_mode = defs.input_mode.video
        self.batch_item = defs.batch_item.default
        h, w, _ = opts.image_shape
        if opts.data_format == defs.data_format.npy:
            blob = np.load(opts.data_path)
            self.frames = blob["frames"]
            self.labels = [list(np.atleast_1d(l)) for l in blob["labels"]]
            self.clips_per_video = [int(c) for c in blob["clips_per_video"]]
            self.num_items = len(self.clips_per_video)
        elif opts.data_format == defs.data_format.synthetic:
            self.num_items = opts.num_items
            cpv = opts.clips_per_video
            self.clips_per_video = [int(c) for c in cpv] if isinstance(cpv, (list, tuple)) else [int(cpv)] * self.num_items
            rng = np.random.default_rng(opts.seed)
            self.labels = [[int(x)] for x in rng.integers(0, num_classes, self.num_items)]
            self.frames = None
            self.rng = rng
            self.shape = (h, w, 3)
        else:
            error("data_format %s needs the reference's TFRecord/raw readers, which are outside the hot path "
                  "(SURVEY 8f #1); use defs.data_format.synthetic or defs.data_format.npy" % opts.data_format)
        if len(self.clips_per_video) != self.num_items:
            error("clips_per_video has %d entries for %d items" % (len(self.clips_per_video), self.num_items))
        self.num_batches = math.ceil(self.num_items / self.batch_size)
        self.batches = list(range(self.num_batches))
        self.batch_index = 0
        self.clip_offsets = np.concatenate([[0], np.cumsum(self.clips_per_video)]).astype(np.int64)
        # save cadence (dataset_.py:562-568)
        self.save_interval = math.ceil(self.num_batches / save_freq_per_epoch) if save_freq_per_epoch else 0
        self.num_saves = math.ceil(save_freq_per_epoch * epochs) if save_freq_per_epoch else 0

    def rewind(self):
        self.batch_index = 0

    def fast_forward(self, batch_index):
        self.batch_index = batch_index

    def next_batch(self):
        lo = self.batch_index * self.batch_size
        hi = min(self.num_items, lo + self.batch_size)
        self.batch_index += 1
        cpvs = self.clips_per_video[lo:hi]
        clips = int(sum(cpvs))
        n = clips * self.fpc
        if self.frames is not None:
            f0 = int(self.clip_offsets[lo]) * self.fpc
            frames = self.frames[f0:f0 + n]
        else:
            frames = self.rng.integers(0, 256, size=(n,) + self.shape, dtype=np.uint8)
        labels = []
        for v in range(lo, hi):  # one label row per clip, replicated from the video (dataset_.py:400-408)
            labels.extend([self.labels[v]] * self.clips_per_video[v])
        return frames, labels_to_one_hot(labels, self.num_classes), cpvs


class Feeder(object):
    def __init__(self, settings):
        self.settings = settings
        self.datasets = {}
        for o in settings.data:
            if o.phase not in settings.phases:
                continue
            opts_phase = settings.train if o.phase == defs.phase.train else settings.val
            epochs = settings.train.epochs if settings.train else 1
            ds = Dataset(o, opts_phase.batch_size, settings.num_classes, epochs, settings.save_freq_per_epoch)
            self.datasets.setdefault(o.phase, []).append(ds)
        if settings.phase not in self.datasets:
            error("No dataset declared for phase %s" % settings.phase)
        self.main = self.datasets[settings.phase][0]
        info("Dataset [%s]: %d items, %d batches of %d, fpc %d" % (
            self.main.opts.name, self.main.num_items, self.main.num_batches, self.main.batch_size, self.main.fpc))

    def loop(self):
        return self.main.batch_index < self.main.num_batches

    def get_feed_dict(self):
        """(frames uint8 [N,H,W,3], onehot int32 [clips,C], clips-per-video of the batch, num_data, num_labels,
        padding=0) -- feeder.py:84-106 without the placeholder indirection."""
        frames, onehot, cpvs = self.main.next_batch()
        return frames, onehot, cpvs, len(frames), len(onehot), 0

    def get_num_batches(self):
        return self.main.num_batches

    def get_batch_index(self):
        return self.main.batch_index

    def get_batch_sizes(self):
        return [d.batch_size for d in self.datasets[self.settings.phase]]

    def should_save(self, run_batch_count):
        si = self.main.save_interval
        return bool(si) and run_batch_count > 0 and run_batch_count % si == 0

    def rewind_datasets(self):
        for d in self.datasets[self.settings.phase]:
            d.rewind()

    def max_clips_per_batch(self):
        m = 0
        for lo in range(0, self.main.num_items, self.main.batch_size):
            m = max(m, int(sum(self.main.clips_per_video[lo:lo + self.main.batch_size])))
        return m
