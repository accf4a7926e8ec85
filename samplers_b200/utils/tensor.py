"""Small tensor helpers (same behaviour as the reference's samplers/utils/tensor.py:5-19)."""
import torch
from torch import Tensor


def validate_tensor_is_scalar(t: Tensor, name: str) -> None:
    if t.ndim != 0:
        raise ValueError(f"`{name}` must be a scalar (0-D tensor).")


def pad_zeros(x: Tensor, target_last_dim: int) -> Tensor:
    """Zero-pad (or crop) the last axis to ``target_last_dim``."""
    cur = x.shape[-1]
    if cur >= target_last_dim:
        return x[..., :target_last_dim]
    return torch.nn.functional.pad(x, (0, target_last_dim - cur))
