"""utils/fixseed.py of the reference (:6-13; imported by the scripts, never called -- SURVEY F9)"""
import random

import numpy as np
import torch


def fixseed(seed):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
