"""The two helpers of the reference's `torchnmf/utils.py` (:5-13), same names, arguments and results."""
import torch

__all__ = ["normalize", "renorm_"]


def normalize(x, axis=0):
    """`x` scaled so that it sums to one along `axis`."""
    total = torch.sum(x, dim=axis, keepdim=True)
    return torch.div(x, total)


def renorm_(input, dim=0):
    """In place: every slice along `dim` divided by its sum of squares (taken over all the other axes)."""
    others = tuple(d for d in range(input.dim()) if d != dim)
    input.div_(input.square().sum(others, keepdim=True))
