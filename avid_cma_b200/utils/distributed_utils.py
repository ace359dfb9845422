"""Mirror of utils/distributed_utils.py:12-19 of the reference."""
import torch
from torch import distributed as dist


def is_distributed():
    return dist.is_available() and dist.is_initialized()


def _gather_from_all(tensor):
    """all_gather + cat along dim 0, rank-major order, no autograd (distributed_utils.py:12-19)."""
    world = dist.get_world_size()
    out = torch.empty((world * tensor.shape[0],) + tuple(tensor.shape[1:]), dtype=tensor.dtype, device=tensor.device)
    dist.all_gather_into_tensor(out, tensor.contiguous())
    return out
