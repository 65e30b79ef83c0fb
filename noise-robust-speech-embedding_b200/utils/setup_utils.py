"""ref:src/utils/setup_utils.py:4-7 seeds torch and numpy only; Python's ``random`` (which draws the noise file, the
SNR and the crop offset in the reference) is seeded here as well so that augmentation streams are reproducible."""
import random

import numpy as np
import torch


def set_seed(seed: int) -> None:
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
