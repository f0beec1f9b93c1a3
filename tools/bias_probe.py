"""Per-update bias of the tensor-core path vs the exact fp32 path at the cfg2 shape, from a given state."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-nmf_b200")]
import torch
from torchnmf_b200 import NMF
from torchnmf_b200.engine import CudaNmfEngine
N, C, R = 65536, 4096, 64
torch.manual_seed(0); V = torch.rand(N, C).bfloat16().float().cuda()
torch.manual_seed(1); W0 = torch.randn(C, R).abs(); H0 = torch.randn(N, R).abs()
def st(a, b, nm):
    rel = ((a - b) / b).flatten()
    print(f"     {nm}: mean {rel.mean():+.3e}  std {rel.std():.3e}  max|.| {rel.abs().max():.3e}", flush=True)
for k0 in (0, 1, 30):
    m = NMF(W=W0, H=H0).cuda()
    if k0: m.fit(V, 1, float("-inf"), k0, precision="f32")
    Wk, Hk = m.W.data.clone(), m.H.data.clone()
    def one(prec, which):
        W, H = Wk.clone(), Hk.clone()
        e = CudaNmfEngine(V, W, H, prec)
        (e.update_w if which == 0 else e.update_h)(1, 1.0, 0.0, 0.0)
        torch.cuda.synchronize(); e.close()
        return W if which == 0 else H
    ref = [one("f32", 0), one("f32", 1)]
    for center in ("0", "1"):
        os.environ["NMFB200_CENTER"] = center
        for prec in ("f16_split", "f16"):
            print(f"state after {k0} iters, center={center}, {prec}")
            st(one(prec, 0), ref[0], "W update")
            st(one(prec, 1), ref[1], "H update")
