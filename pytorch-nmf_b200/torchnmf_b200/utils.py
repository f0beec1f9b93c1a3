"""Small helpers of the reference's `torchnmf/utils.py` (:5-13), same names and semantics."""
import torch

__all__ = ["normalize", "renorm_"]


def normalize(x: torch.Tensor, axis=0) -> torch.Tensor:
    """x scaled to unit sum along `axis` (utils.py:5-6)."""
    return x / x.sum(axis, keepdim=True)


def renorm_(input: torch.Tensor, dim=0):
    """In place: divide by the sum of squares over every axis but `dim` (utils.py:9-13)."""
    dims = [d for d in range(input.dim()) if d != dim]
    input /= (input * input).sum(dims, keepdim=True)
