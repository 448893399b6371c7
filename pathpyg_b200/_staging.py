"""Host <-> device staging so that results follow the device of the inputs, as in the reference.

The kernels only run on the GPU.  When a caller hands in host tensors (the reference's default),
they are copied to the current CUDA device (asynchronously when pinned), the CUDA path runs, and
the results are copied back -- this is the ``e2e`` path ``bench.py`` measures.  Without a CUDA
device this raises: there is no CPU implementation to fall back to.
"""
from __future__ import annotations

import torch


def compute_device(*tensors) -> tuple[torch.device, bool]:
    """(device the kernels run on, whether results must be copied back to the host)."""
    for t in tensors:
        if isinstance(t, torch.Tensor) and t.is_cuda:
            return t.device, False
    if not torch.cuda.is_available():
        raise RuntimeError(
            "pathpyg_b200 needs a CUDA device: the lift / DBGNN hot path has no CPU implementation "
            "(use the reference pathpyG on CPU-only machines)")
    return torch.device("cuda", torch.cuda.current_device()), True


def up(t, device):
    if t is None or not isinstance(t, torch.Tensor) or t.device == device:
        return t
    return t.as_subclass(torch.Tensor).to(device, non_blocking=True)


def down(t, to_host: bool):
    if not to_host or t is None or not isinstance(t, torch.Tensor):
        return t
    return t.cpu()
