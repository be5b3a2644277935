"""torch plumbing: zero-copy torch views of handle-owned device buffers (for
torch.distributed collectives over NCCL) and pinned host staging."""
import numpy as np


class _DevView:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def device_view(ptr, n, dtype="float64", device="cuda:0"):
    """A torch tensor aliasing `n` elements of device memory at `ptr` (no copy, no ownership)."""
    import torch
    typestr = np.dtype(dtype).str
    return torch.as_tensor(_DevView(ptr, n, typestr), device=device)
