"""Validation head: clip -> video fusion, accuracy, logits chunks (reference: val.py:59-203).

The per-video fusion runs on the device through the segmented pooling kernel (vl_segment_pool_fwd), whose
summation order reproduces numpy's `np.mean(rows, axis=0)` bit for bit; `use_device=False` keeps the pure numpy
path (used by the CPU tests against the reference's golden vectors).
"""
import os
import pickle
import time

import numpy as np

from .defs import defs
from .utils import debug, error, info


class Validation(object):
    def __init__(self, settings, num_classes=None, use_device=True):
        self.settings = settings
        nc = num_classes if num_classes is not None else settings.num_classes
        self.item_logits = np.zeros([0, nc], np.float32)
        self.item_labels = np.zeros([0, nc], np.float32)
        self.validation_logits_save_counter = 0
        self.validation_logits_save_interval = None
        self.run_id = getattr(settings, "run_id", "run")
        self.run_folder = getattr(settings, "run_folder", ".")
        self.timestamp = time.strftime("%d%m%y_%H%M%S")
        self.use_device = use_device
        if getattr(settings, "val", None):
            self.validation_logits_save_interval = settings.val.logits_save_interval

    # -- fusion ---------------------------------------------------------------------------------------
    def fuse_videos(self, logits, cpvs, clip_fusion):
        """[sum(cpvs), C] clip logits -> [len(cpvs), C] video logits (avg: np.mean(axis=0); last: last row)."""
        if clip_fusion not in (defs.fusion_method.avg, defs.fusion_method.last):
            # the reference leaves `video_logits` unbound for any other method (val.py:161-165) -> NameError
            error("Undefined clip fusion method : %s" % clip_fusion)
        if self.use_device:
            import torch
            from . import _native as nv
            seg = np.concatenate([[0], np.cumsum(cpvs)]).astype(np.int32)
            x = torch.from_numpy(np.ascontiguousarray(logits, dtype=np.float32)).cuda()
            y = torch.empty(len(cpvs), logits.shape[1], dtype=torch.float32, device="cuda")
            nv.call("vl_segment_pool_fwd", x, torch.from_numpy(seg).cuda(), 0, len(cpvs), logits.shape[1],
                    0 if clip_fusion == defs.fusion_method.avg else 1, y, None)
            return y.cpu().numpy()
        out, off = [], 0
        for cpv in cpvs:
            rows = logits[off:off + cpv]
            out.append(np.mean(rows, axis=0) if clip_fusion == defs.fusion_method.avg else rows[-1])
            off += cpv
        return np.vstack(out).astype(np.float32)

    def process_validation_logits(self, logits, labels, cpvs, clip_fusion, input_mode=defs.input_mode.video):
        """Default batch_item + video mode of val.py:91-110: `cpvs` lists the clips-per-video of the videos of this
        batch; the label of a video is the label row of its first clip."""
        logits = np.asarray(logits, np.float32)
        labels = np.asarray(labels)
        if input_mode != defs.input_mode.video:
            self.add_item_logits_labels(logits, labels)  # frames: simply append (val.py:113)
            return
        if sum(cpvs) != len(logits) or len(labels) != len(logits):
            error("Logits and/or labels non empty at the end of video item mode aggregation!")
        video_logits = self.fuse_videos(logits, cpvs, clip_fusion)
        starts = np.concatenate([[0], np.cumsum(cpvs)[:-1]]).astype(np.int64)
        self.add_item_logits_labels(video_logits, labels[starts])
        debug("Video logits and labels accumulation is now %d,%d video in video batch mode." % (
            len(self.item_logits), len(self.item_labels)))
        info("Incremental accuracy up to current batch: %2.3f" % np.mean(
            self.get_chunk_accuracy(self.item_logits, self.item_labels[-len(self.item_logits):])))

    def add_item_logits_labels(self, logits, label):
        self.item_logits = np.vstack((self.item_logits, logits))
        self.item_labels = np.vstack((self.item_labels, label))

    # -- persistence (val.py:115-156): pickled float32 ndarrays, same file names ---------------------------
    def save_validation_logits_chunk(self, save_all=False):
        if self.validation_logits_save_interval is None or len(self.item_logits) == 0:
            return
        if self.validation_logits_save_interval <= 0:
            if save_all:
                path = os.path.join(self.run_folder, "validation_logits_%s_%s.total" % (self.run_id, self.timestamp))
                info("Saving all %d extracted validation logits to %s" % (len(self.item_logits), path))
                with open(path, "wb") as f:
                    pickle.dump(self.item_logits, f)
            return
        if len(self.item_logits) >= self.validation_logits_save_interval or save_all:
            path = os.path.join(self.run_folder, "validation_logits_%s_%s.part_%d" % (
                self.run_id, self.timestamp, self.validation_logits_save_counter))
            info("Saving a %d-sized chunk of validation logits to %s" % (len(self.item_logits), path))
            with open(path, "wb") as f:
                pickle.dump(self.item_logits, f)
            self.item_logits = np.zeros([0, self.item_logits.shape[-1]], np.float32)
            self.validation_logits_save_counter += 1

    def load_validation_logits_chunk(self, chunk_idx):
        if self.validation_logits_save_interval is None:
            return self.item_logits
        path = os.path.join(self.run_folder, "validation_logits_%s_%s.part_%d" % (self.run_id, self.timestamp, chunk_idx))
        with open(path, "rb") as f:
            return pickle.load(f)

    # -- accuracy (val.py:174-203) ----------------------------------------------------------------------
    def get_chunk_accuracy(self, logits, labels):
        return np.mean(np.equal(np.argmax(logits, axis=1), np.argmax(labels, axis=1)))

    def get_accuracy(self):
        """Unweighted mean of per-chunk accuracies, like the reference (val.py:197)."""
        accuracies, cursor = [], 0
        for idx in range(self.validation_logits_save_counter):
            chunk = self.load_validation_logits_chunk(idx)
            labels = self.item_labels[cursor:cursor + len(chunk), :]
            accuracies.append(self.get_chunk_accuracy(chunk, labels))
            cursor += len(chunk)
        if len(self.item_logits) > 0:
            labels = self.item_labels[cursor:cursor + len(self.item_logits), :]
            accuracies.append(self.get_chunk_accuracy(self.item_logits, labels))
        return np.mean(accuracies)
