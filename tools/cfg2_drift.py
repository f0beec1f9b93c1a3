"""Drift study at the BASELINE cfg2 shape: tensor-core modes vs the exact fp32 CUDA-core path (and the reference
golden at 200 iterations).  Prints max / median / signed-mean relative error of W and H."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-nmf_b200")]
import numpy as np, torch
from torchnmf_b200 import NMF
N, C, R = 65536, 4096, 64
torch.manual_seed(0); V = torch.rand(N, C).bfloat16().float().cuda()
torch.manual_seed(1); W0 = torch.randn(C, R).abs(); H0 = torch.randn(N, R).abs()
z = np.load(os.path.join(ROOT, "tests/golden/nmf_cfg2_kl_200.npz"))
def stats(a, b, nm):
    big = b.abs() > 1e-3 * b.abs().max()
    rel = ((a - b) / b)[big]
    print(f"   {nm}: max|rel| {rel.abs().max():.2e} med|rel| {rel.abs().median():.2e} mean(rel) {rel.mean():+.2e}", flush=True)
its = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [10, 50, 200]
modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["f32", "f16_split", "f16"]
res = {}
for it in its:
    for prec in modes:
        m = NMF(W=W0, H=H0).cuda()
        t0 = time.time(); m.fit(V, 1, float("-inf"), it, precision=prec); torch.cuda.synchronize()
        res[(it, prec)] = (m.W.data.clone(), m.H.data.clone())
        print(f"iters {it} {prec}: {time.time()-t0:.2f}s", flush=True)
        if prec != "f32" and (it, "f32") in res:
            stats(m.W.data, res[(it, "f32")][0], "W vs f32 path")
            stats(m.H.data, res[(it, "f32")][1], "H vs f32 path")
        if it == 200:
            stats(m.W.data.cpu()[::8], torch.from_numpy(z["W_sub"]), "W vs reference golden")
            stats(m.H.data.cpu()[::128], torch.from_numpy(z["H_sub"]), "H vs reference golden")
