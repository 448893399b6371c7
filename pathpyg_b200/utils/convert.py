"""``to_numpy`` (reference ``src/pathpyG/utils/convert.py``): tensors (also on a GPU), ``EdgeIndex`` objects, lists and
arrays as a host numpy array."""
from __future__ import annotations

import numpy as np
import torch


def to_numpy(values) -> np.ndarray:
    if isinstance(values, torch.Tensor):
        return values.as_subclass(torch.Tensor).detach().cpu().numpy()
    if isinstance(values, np.ndarray):
        return values
    return np.array(values)
