"""Batch feeding contract of the hot path (reference: feeder.py:84-141, dataset_.py:386-420,562-613).

What the device path depends on is the integer contract, which this module keeps: frames ordered
video -> clip -> frame, ONE label row per clip, whole videos per batch, `padding` always 0.  Sources:
  tfrecord   the reference's serialized datasets (`<path>.tfrecord` + `<path>.size`, serialize.py:138-151,246-256) read
             WITHOUT TensorFlow (tfrecord.py); the read-time preprocessing of dataset_.py:481-501 is split: crop
             offsets and mirror flags are drawn here with the reference's own RNG calls (`choice`, `randrange`), the
             pixels are cropped / mirrored / mean-subtracted on the device by vl_frames_s2d_crop;
  synthetic  seeded uint8 frames, UCF101-shaped;   npy   a .npz with `frames` uint8 [N,H,W,3], `labels`, `clips_per_video`.
"""
import math
import os
from random import choice, randrange

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
        elif opts.data_format == defs.data_format.tfrecord:
            from . import tfrecord
            base = opts.data_path[:-len(".tfrecord")] if str(opts.data_path).endswith(".tfrecord") else opts.data_path
            self.tfr_path = base + ".tfrecord"
            if not os.path.exists(self.tfr_path):
                error("TFRecord file path does not exist: %s" % self.tfr_path)
            # serialize.py:138-151,664: the side file of `<inp>.tfrecord` is `<inp>.tfrecord.size` (dataset_.py:703 reads
            # `self.path + ".size"` after appending ".tfrecord"); `<inp>.size` is accepted as well
            size_file = self.tfr_path + ".size" if os.path.exists(self.tfr_path + ".size") else base + ".size"
            if not os.path.exists(size_file):
                error("Could not file data size file: %s" % (self.tfr_path + ".size"))
            meta = tfrecord.read_size_file(size_file)
            if meta["type"] != defs.input_mode.video:
                error("Specified input mode is [%s] but the size file contains [%s]" % (defs.input_mode.video, meta["type"]))
            self.num_items = meta["items"]
            self.clips_per_video = [int(c) for c in meta["cpi"]]
            if meta["fpc"] != self.fpc:
                self.fpc = int(meta["fpc"])  # the .size file is authoritative (model.py:56-60)
            self.frames = None
            self.labels = None
            self._tfrecord = tfrecord
            self._records = None
            self.verify = getattr(opts, "verify_records", "length")
            self.imgproc = list(getattr(opts, "imgproc", []) or [])
            self.raw_shape = tuple(getattr(opts, "raw_image_shape", None) or (h, w, 3))
            # imgproc raw_resize / resize (dataset_.py:481-491): scipy.misc.imresize at read time.  The frames stay at
            # their serialized size on the host; the engine resamples them on the device (Engine.set_read_resize,
            # csrc/resize.cu: Pillow's bilinear, bit-exact) before the crop / staging kernel.
            self.resize_to = None
            if defs.imgproc.raw_resize in self.imgproc:
                if getattr(opts, "raw_image_shape", None) is None:
                    error("imgproc raw_resize needs raw_image_shape")
                self.resize_to = tuple(self.raw_shape[:2])
            has_crop = defs.imgproc.rand_crop in self.imgproc or defs.imgproc.center_crop in self.imgproc
            if defs.imgproc.resize in self.imgproc and not has_crop:  # dataset_.py:486-491: crop wins over resize
                if self.resize_to is not None:
                    error("imgproc raw_resize followed by resize resamples every frame twice; serialize the frames at "
                          "raw_image_shape or drop one of the two")
                self.resize_to = (h, w)
                self.raw_shape = (h, w, 3)
            self.stored_shape = None  # serialized frame shape, read from the first record when a resize is active
            self.crop_h = self.crop_w = None
            if defs.imgproc.rand_crop in self.imgproc:  # dataset_.py:571-577: ranges of admissible offsets
                self.crop_h = list(range(0, self.raw_shape[0] - h - 1))
                self.crop_w = list(range(0, self.raw_shape[1] - w - 1))
            elif defs.imgproc.center_crop in self.imgproc:
                self.crop_h = int(math.floor((self.raw_shape[0] - h) / 2))
                self.crop_w = int(math.floor((self.raw_shape[1] - w) / 2))
            elif tuple(self.raw_shape[:2]) != (h, w):
                error("Encountered image shape %s but desired shape is %s" % (str(self.raw_shape), str((h, w, 3))))
            self.net_hw = (h, w)
        else:
            error("data_format %s reads loose image files with scipy.misc.imread, which is outside the hot path; "
                  "serialize them (tfrecord) or use defs.data_format.synthetic / npy" % opts.data_format)
        self.last_crops = None
        if not hasattr(self, "resize_to"):
            self.resize_to = None
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
        self._records = None

    def fast_forward(self, batch_index):
        """Resume inside an epoch (feeder.py:143-194): skip the records of the batches already consumed."""
        self.batch_index = batch_index
        if getattr(self, "_tfrecord", None) is not None:
            self._records = self._tfrecord.read_records(self.tfr_path, self.verify)
            skip = int(self.clip_offsets[min(batch_index * self.batch_size, self.num_items)]) * self.fpc
            for _ in range(skip):
                next(self._records)

    def _next_tfrecord_frames(self, n):
        """n consecutive frames of the serialization (dataset_.py:171-217) + their crop / mirror draws."""
        if self._records is None:
            self._records = self._tfrecord.read_records(self.tfr_path, self.verify)
        frames = None if self.resize_to is not None else np.empty((n,) + tuple(self.raw_shape), np.uint8)
        labels, crops = [], np.zeros((n, 3), np.int32)
        h, w = self.net_hw
        for i in range(n):
            try:
                payload = next(self._records)
            except StopIteration:
                error("Encountered unexpected EOF while reading TFRecord example # %d in the batch." % i)
            image, label = self._tfrecord.deserialize_frame(payload)
            if self.resize_to is not None:
                if frames is None:  # frames keep their serialized size; the device resamples them
                    self.stored_shape = tuple(image.shape)
                    frames = np.empty((n,) + self.stored_shape, np.uint8)
                if image.shape != self.stored_shape:
                    error("Encountered image shape %s after %s: one batch must hold frames of one serialized size" % (
                        str(image.shape), str(self.stored_shape)))
            elif image.shape != tuple(self.raw_shape):
                error("Encountered image shape %s but the dataset declares %s" % (str(image.shape), str(self.raw_shape)))
            frames[i] = image
            labels.append(label)
            # process_image (dataset_.py:481-501): crop first (choice(h) then choice(w)), mirror last
            if defs.imgproc.rand_crop in self.imgproc:
                crops[i, 0], crops[i, 1] = choice(self.crop_h), choice(self.crop_w)
            elif defs.imgproc.center_crop in self.imgproc:
                crops[i, 0], crops[i, 1] = self.crop_h, self.crop_w
            if defs.imgproc.rand_mirror in self.imgproc:
                crops[i, 2] = 0 if randrange(2) else 1
        return frames, labels, crops

    def next_batch(self):
        lo = self.batch_index * self.batch_size
        hi = min(self.num_items, lo + self.batch_size)
        self.batch_index += 1
        cpvs = self.clips_per_video[lo:hi]
        clips = int(sum(cpvs))
        n = clips * self.fpc
        self.last_crops = None
        if getattr(self, "_tfrecord", None) is not None:
            frames, per_frame, self.last_crops = self._next_tfrecord_frames(n)
            labels, f0 = [], 0
            for c in cpvs:  # one label row per clip, from the first frame of the video (dataset_.py:400-408)
                labels.extend([per_frame[f0]] * c)
                f0 += c * self.fpc
            return frames, labels_to_one_hot(labels, self.num_classes), cpvs
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
        self.last_crops = None
        info("Dataset [%s]: %d items, %d batches of %d, fpc %d" % (
            self.main.opts.name, self.main.num_items, self.main.num_batches, self.main.batch_size, self.main.fpc))

    def loop(self):
        return self.main.batch_index < self.main.num_batches

    def get_feed_dict(self):
        """(frames uint8 [N,H,W,3], onehot int32 [clips,C], clips-per-video of the batch, num_data, num_labels,
        padding=0) -- feeder.py:84-106 without the placeholder indirection."""
        frames, onehot, cpvs = self.main.next_batch()
        self.last_crops = self.main.last_crops
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
