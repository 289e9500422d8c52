"""Logging helpers of the reference that the path's callers use (src/utils/tracking.py:10-18, 54-55)."""
import logging
from typing import List


def setup_basic_logger():
    """A root-logger stream handler; unlike the reference it does not wipe handlers other code installed."""
    logger = logging.getLogger("lds_gnn_b200")
    if not logger.handlers:
        handler = logging.StreamHandler()
        handler.setFormatter(logging.Formatter(fmt="%(asctime)s (%(levelname)s): %(message)s", datefmt="%Y-%m-%d %H:%M:%S"))
        logger.addHandler(handler)
        logger.setLevel(logging.WARNING)
    return logger


def get_lr(optimizer) -> List[float]:
    return [group["lr"] for group in optimizer.param_groups]
