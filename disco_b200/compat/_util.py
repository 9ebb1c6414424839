import numpy as np
import torch

DEVICE = "cuda"


def dev(a, dtype):
    """NumPy array / torch tensor -> contiguous CUDA tensor of `dtype` (no CPU compute path exists)."""
    if isinstance(a, torch.Tensor):
        return a.to(device=DEVICE, dtype=dtype).contiguous()
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a))).to(device=DEVICE, dtype=dtype).contiguous()
