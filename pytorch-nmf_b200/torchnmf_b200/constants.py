"""`eps` exactly as torchnmf/constants.py:3 -- float32 machine epsilon (2**-23)."""
import torch

eps = torch.finfo(torch.float32).eps
