"""Short single-GPU target for ncu: three NMFD iterations + one loss at the cfg3 shape (tensor-core path)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-nmf_b200")]
import torch
from torchnmf_b200.engine import CudaNmfdEngine
torch.manual_seed(0)
V = torch.rand(1, 1025, 8192).bfloat16().float().cuda()
W = torch.randn(1025, 16, 128).abs().cuda(); H = torch.randn(1, 16, 8065).abs().cuda()
eng = CudaNmfdEngine(V, W, H, "f16")
for _ in range(3):
    eng.update_w(1, 1.0, 0.0, 0.0); eng.update_h(1, 1.0, 0.0, 0.0)
print("loss", eng.loss(1))
