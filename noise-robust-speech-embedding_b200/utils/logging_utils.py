"""Logger "nrse" with the reference's semantics (ref:src/utils/logging_utils.py:18-66): one timestamped file under
``training.log_dir`` at ``logging.level`` plus a console handler at ``logging.console_level``."""
from __future__ import annotations

import logging
import os
import time

logger = logging.getLogger("nrse")
logger.addHandler(logging.NullHandler())


def setup_logger(config) -> logging.Logger:
    log_cfg = config.get("logging", {})
    level = getattr(logging, str(log_cfg.get("level", "INFO")).upper(), logging.INFO)
    console_level = getattr(logging, str(log_cfg.get("console_level", "ERROR")).upper(), logging.ERROR)
    logger.setLevel(min(level, console_level))
    for h in list(logger.handlers):
        logger.removeHandler(h)
    fmt = logging.Formatter("%(asctime)s - %(name)s - %(levelname)s - %(message)s")
    log_dir = config.get("training", {}).get("log_dir")
    if log_dir:
        os.makedirs(log_dir, exist_ok=True)
        fh = logging.FileHandler(os.path.join(log_dir, time.strftime("train_%Y%m%d_%H%M%S.log")))
        fh.setLevel(level)
        fh.setFormatter(fmt)
        logger.addHandler(fh)
    ch = logging.StreamHandler()
    ch.setLevel(console_level)
    ch.setFormatter(fmt)
    logger.addHandler(ch)
    logger.propagate = False
    return logger
