"""Small tensor-core fits for compute-sanitizer (racecheck / synccheck): every ring configuration of the fused kernel.

    compute-sanitizer --tool racecheck python tools/small_fit.py
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-nmf_b200")]
import torch
from torchnmf_b200 import NMF
torch.manual_seed(0)
for (N, C, R, prec, beta) in ((512, 640, 128, "f16", 1.0), (512, 640, 64, "f16_split", 1.0), (384, 512, 64, "f16", 1.0),
                              (384, 512, 128, "f16_split", 1.0), (384, 512, 48, "f16", 0.5), (384, 512, 64, "f16", 2.0)):
    V = (torch.rand(N, C) + 0.01).cuda()
    m = NMF((N, C), R).cuda()
    n = m.fit(V, beta, float("-inf"), 3, precision=prec)
    print(N, C, R, prec, beta, "->", n, m.last_fit_precision, float(m.W.data.sum()), flush=True)
