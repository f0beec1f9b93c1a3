"""Developer probe: two-output (beta != 1) tensor-core kernels vs the CPU oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-nmf_b200")]
import torch
from oracle import mu_oracle as orc
from torchnmf_b200.engine import CudaNmfEngine
from torchnmf_b200 import NMF

def stats(got, want, name):
    d = (got - want).abs()
    rel = d / (want.abs() + 1e-6 * want.abs().max())
    print(f"  {name}: maxrel {rel.max():.3e} medrel {rel.median():.3e} mean(signed) {((got-want)/(want.abs()+1e-6*want.abs().max())).mean():+.2e}", flush=True)

for (N, C, R) in [(384, 256, 64), (300, 200, 40), (2048, 1024, 64)]:
    torch.manual_seed(0)
    V = (torch.rand(N, C) + 0.01).bfloat16().float()
    W0 = torch.randn(C, R).abs() + 0.01; H0 = torch.randn(N, R).abs() + 0.01
    for beta in (0.0, 0.5, 1.5, 3.0, -1.0):
        print(f"N={N} C={C} R={R} beta={beta}", flush=True)
        Wd, Hd = W0.cuda(), H0.cuda()
        eng = CudaNmfEngine(V.cuda(), Wd, Hd, "f16")
        g = orc.gamma_of(beta)
        eng.update_w(beta, g, 0.0, 0.0)
        Wn = orc.nmf_update_w(V, W0, H0, beta)
        stats(Wd.cpu(), Wn, "W after update_w")
        eng.update_h(beta, g, 0.01, 0.02)
        Hn = orc.nmf_update_h(V, Wn, H0, beta, g, 0.01, 0.02)
        stats(Hd.cpu(), Hn, "H after update_h")
        eng.close()
# multi-iteration drift vs f32 path
torch.manual_seed(1)
N, C, R = 4096, 2048, 64
V = (torch.rand(N, C) + 0.01).bfloat16().float().cuda()
W0 = torch.randn(C, R).abs(); H0 = torch.randn(N, R).abs()
for beta in (0.0, 0.5, 1.5):
    out = {}
    for prec in ("f32", "f16"):
        m = NMF(W=W0, H=H0).cuda(); m.fit(V, beta, float("-inf"), 50, precision=prec); out[prec] = (m.W.data.clone(), m.H.data.clone())
    print(f"50 iterations beta={beta}:", flush=True)
    stats(out["f16"][0].cpu(), out["f32"][0].cpu(), "W tc vs f32")
    stats(out["f16"][1].cpu(), out["f32"][1].cpu(), "H tc vs f32")
print("done")
