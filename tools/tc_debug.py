"""Developer probe for the tcgen05 path: compares one W-numerator and one H update against the CPU oracle
on a few shapes and prints error statistics (run on the GPU box)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-nmf_b200")]
import torch
from oracle import mu_oracle as orc
from torchnmf_b200.engine import CudaNmfEngine

def stats(got, want, name):
    d = (got - want).abs()
    rel = d / (want.abs() + 1e-6 * want.abs().max())
    print(f"  {name}: maxabs {d.max():.3e} maxrel {rel.max():.3e} medrel {rel.median():.3e} (|want|max {want.abs().max():.3e})", flush=True)
    return rel.max().item()

def run(N, C, R, prec):
    torch.manual_seed(0)
    V = torch.rand(N, C).bfloat16().float()
    W0 = torch.randn(C, R).abs(); H0 = torch.randn(N, R).abs()
    print(f"shape N={N} C={C} R={R} precision={prec}", flush=True)
    Wd, Hd = W0.cuda(), H0.cuda()
    eng = CudaNmfEngine(V.cuda(), Wd, Hd, prec)
    num, den = orc.nmf_w_contractions(V, W0, H0, 1)
    buf = eng.w_partial(1).cpu()
    bad = stats(buf[:C * R].view(C, R), num, "W numerator")
    stats(buf[C * R:], den.view(-1), "colsum(H)")
    if bad > 1e-2:
        g = buf[:C * R].view(C, R); r = g / num
        print("  ratio got/want corner:\n", r[:4, :8], "\n  rows 64..:", r[64:66, :8] if C > 66 else "", flush=True)
    eng.update_h(1, 1.0, 0.0, 0.0)
    Hn = orc.nmf_update_h(V, W0, H0, 1)
    bad = stats(Hd.cpu(), Hn, "H after update_h")
    if bad > 1e-2:
        r = Hd.cpu() / Hn
        print("  ratio got/want corner:\n", r[:4, :8], flush=True)
    eng.update_w(1, 1.0, 0.0, 0.0)
    Wn = orc.nmf_update_w(V, W0, Hn, 1)
    stats(Wd.cpu(), Wn, "W after update_w (on new H)")
    eng.close()

if __name__ == "__main__":
    prec = sys.argv[1] if len(sys.argv) > 1 else "f16"
    shapes = [(128, 128, 64), (384, 256, 64), (300, 200, 40), (2048, 1024, 64), (512, 384, 128), (300, 260, 100)]
    for s in shapes:
        run(*s, prec)
    print("done", prec)
