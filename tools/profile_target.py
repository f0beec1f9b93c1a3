"""Short single-GPU target for ncu: a few launches of the fused contraction kernels at the cfg2 shape."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-nmf_b200")]
import torch
from torchnmf_b200.engine import CudaNmfEngine
prec = sys.argv[1] if len(sys.argv) > 1 else "f16_split"
N, C, R = (int(x) for x in (sys.argv[2:5] if len(sys.argv) > 4 else (65536, 4096, 64)))
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
beta = float(sys.argv[6]) if len(sys.argv) > 6 else 1.0
torch.manual_seed(0)
V = torch.rand(N, C, device="cuda").bfloat16().float()
W = torch.randn(C, R, device="cuda").abs(); H = torch.randn(N, R, device="cuda").abs()
eng = CudaNmfEngine(V, W, H, prec)
for _ in range(reps):
    eng.update_w(beta, 1.0, 0.0, 0.0)
    eng.update_h(beta, 1.0, 0.0, 0.0)
print("loss", eng.loss(beta))
torch.cuda.synchronize()
eng.close()
