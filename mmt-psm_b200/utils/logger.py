"""Run logger with the reference's contract (utils/logger.py:7-28): rank 0 writes to stdout and to OUTPUT_DIR/log.txt
(log_test.txt for evaluation runs), the other ranks get a silent logger of the same name."""
import logging
import os
import sys

_FORMAT = "%(asctime)s %(name)s %(levelname)s: %(message)s"


def setup_logger(name, save_dir, distributed_rank, test=False):
    logger = logging.getLogger(name)
    logger.setLevel(logging.DEBUG)
    if distributed_rank > 0:
        logger.addHandler(logging.NullHandler())
        logger.propagate = False                 # nothing from the non-master ranks reaches the root handlers either
        return logger
    sinks = [logging.StreamHandler(stream=sys.stdout)]
    if save_dir:
        os.makedirs(save_dir, exist_ok=True)
        sinks.append(logging.FileHandler(os.path.join(save_dir, "log_test.txt" if test else "log.txt")))
    have = {(type(h), getattr(h, "baseFilename", None)) for h in logger.handlers}
    for h in sinks:
        if (type(h), getattr(h, "baseFilename", None)) in have:      # called twice for the same run: no duplicate lines
            continue
        h.setLevel(logging.DEBUG)
        h.setFormatter(logging.Formatter(_FORMAT))
        logger.addHandler(h)
    logger.propagate = False
    return logger
