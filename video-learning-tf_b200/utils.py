"""Logging helpers and label utilities of the host shell (reference: utils_.py:41-94,133-169)."""
import logging
import time

import numpy as np

LOGGER_NAME = "vlb200"


def get_logger():
    return logging.getLogger(LOGGER_NAME)


def configure_logging(logfile=None, level=logging.INFO):
    logger = get_logger()
    logger.setLevel(level)
    logger.handlers = []
    fmt = logging.Formatter("%(asctime)s| %(levelname)7s - %(message)s")
    sh = logging.StreamHandler()
    sh.setFormatter(fmt)
    logger.addHandler(sh)
    if logfile:
        fh = logging.FileHandler(logfile)
        fh.setFormatter(fmt)
        logger.addHandler(fh)
    return logger


def info(msg):
    get_logger().info(msg)


def debug(msg):
    get_logger().debug(msg)


def warning(msg):
    get_logger().warning(msg)


def error(msg):
    """Log and raise, like utils_.error (utils_.py:133-136): configuration and runtime errors are exceptions."""
    get_logger().error(msg)
    raise Exception(msg)


def elapsed_str(tic):
    sec = time.time() - tic
    m, s = divmod(int(sec), 60)
    h, m = divmod(m, 60)
    return "%02d:%02d:%02d" % (h, m, s)


def labels_to_one_hot(labels, num_classes):
    """List of per-item label lists -> int32 one/multi-hot [items, num_classes] (utils_.py:160-169)."""
    if not isinstance(labels, list):
        labels = [labels]
    labels = [list(l) if isinstance(l, (list, tuple, np.ndarray)) else [l] for l in labels]
    maxlbl = max(lbl for item in labels for lbl in item)
    if maxlbl >= num_classes:
        error("Encountered label %d but the number of labels was set to %d" % (maxlbl, num_classes))
    onehots = np.zeros((len(labels), num_classes), dtype=np.int32)
    for row, item in enumerate(labels):
        onehots[row, item] = 1
    return onehots
