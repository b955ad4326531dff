"""Clip / frame index generation per `clipframe_mode` (reference: serialize.py:269-408).

Integer contract: for a given Python `random` state the generated frame indices are identical to the reference's
(the generators draw from `random.choice` in the same order).  Checked bit-exactly against golden vectors produced
by the reference's own functions (tests/golden/reference_host_golden.json).
"""
import random
from os.path import basename

from .defs import defs
from .utils import debug, error


def _report(settings, message, path):
    logger = getattr(settings, "logger", None)
    if logger is not None and hasattr(logger, "add_to_log_storage"):
        logger.add_to_log_storage("generation", (message, path))


def get_random_clips(avail_frame_idxs, settings, path):
    """`clip_offset_or_num` random clips of `num_frames_per_clip` consecutive frames; starts within one clip length
    of an already chosen start are avoided until the pool of starts is exhausted (serialize.py:293-355)."""
    fpc, want = settings.num_frames_per_clip, settings.clip_offset_or_num
    num_frames = len(avail_frame_idxs)
    if num_frames == 0:
        error("No frames for path [%s]" % path)
    missing = fpc - num_frames
    if missing > 0:
        message = "Video %s cannot sustain a number of %d fpc, as it has %d frames" % (basename(path), fpc, num_frames)
        debug(message)
        if settings.generation_error == defs.generation_error.abort:
            error(message)
        _report(settings, message, path)
        if settings.generation_error == defs.generation_error.compromise:
            padded = [0] * missing + list(avail_frame_idxs)  # replicate the first frame in front
            return [padded for _ in range(want)]
        if settings.generation_error != defs.generation_error.report:
            error("Undefined generation error strategy: %s" % settings.generation_error)
    all_starts = list(range(num_frames - fpc + 1))
    short = want - len(all_starts)
    if short > 0:
        message = "Video %s cannot sustain a number of %d cpv as it has %d frames" % (basename(path), want, num_frames)
        debug(message)
        if settings.generation_error == defs.generation_error.abort:
            error(message)
        _report(settings, message, path)
        if settings.generation_error == defs.generation_error.compromise:
            all_starts.extend([random.choice(all_starts) for _ in range(short)])
        elif settings.generation_error == defs.generation_error.report:
            return []
        else:
            error("Undefined generation error strategy: %s" % settings.generation_error)
    starts = []
    pool = list(all_starts)
    for _ in range(want):
        start = random.choice(pool)
        starts.append(start)
        for i in range(start - fpc + 1, start + fpc):  # list.remove drops ONE occurrence, like the reference
            if i in pool:
                pool.remove(i)
        if not pool:
            pool = list(all_starts)
    return [list(range(s, s + fpc)) for s in starts]


def get_sequential_clips(avail_frame_idxs, settings, path):
    """All clips whose starts are `num_frames_per_clip + clip_offset_or_num` apart (serialize.py:357-378)."""
    fpc = settings.num_frames_per_clip
    num_frames = len(avail_frame_idxs)
    missing = fpc - num_frames
    if missing > 0:
        message = "Attempted to get %d-framed sequential clips from video %s which has %d frames." % (
            fpc, basename(path), num_frames)
        if settings.generation_error == defs.generation_error.abort:
            error(message)
        _report(settings, message, path)
        if settings.generation_error == defs.generation_error.compromise:
            avail_frame_idxs.extend([random.choice(avail_frame_idxs) for _ in range(missing)])
        elif settings.generation_error == defs.generation_error.report:
            return []
        else:
            error("Undefined generation error strategy: %s" % settings.generation_error)
    distance = fpc + settings.clip_offset_or_num
    return [list(range(s, s + fpc)) for s in range(0, num_frames - fpc + 1, distance)]


def get_random_frames(avail_frame_idxs, settings, path):
    """The reference's rand_frames generator is broken (`shuffle` returns None, serialize.py:271); this framework
    refuses the mode instead of guessing the intent."""
    error("clipframe_mode rand_frames is not usable in the reference (serialize.py:271) and is not implemented")


def generate_clips(num_frames, settings, path="<video>"):
    """Mode switch of generate_frames_for_video (serialize.py:381-398) on frame indices."""
    idxs = list(range(num_frames))
    mode = settings.clipframe_mode
    if mode == defs.clipframe_mode.rand_clips:
        return get_random_clips(idxs, settings, path)
    if mode == defs.clipframe_mode.iterative:
        return get_sequential_clips(idxs, settings, path)
    if mode == defs.clipframe_mode.rand_frames:
        return get_random_frames(idxs, settings, path)
    return []
