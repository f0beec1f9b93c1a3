"""Timing of the BASELINE.json configs that are parity-test cases rather than bench lines:
cfg3 (NMFD 1025x8192 R=16 T=128 beta=1) and cfg5 (beta sweep at the cfg2 shape)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-nmf_b200")]
import torch
from torchnmf_b200 import NMF, NMFD
out = {}
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
# cfg3
torch.manual_seed(0)
V = torch.rand(1, 1025, 8192).cuda()
m = NMFD(V.shape, rank=16, T=128).cuda()
W0, H0 = m.W.data.clone(), m.H.data.clone()
def run3():
    m.W.data.copy_(W0); m.H.data.copy_(H0); m.fit(V, 1, float("-inf"), 20)
dt = timed(run3)
out["cfg3_nmfd_it_per_s"] = 20 / dt
print("cfg3 NMFD 1025x8192 R=16 T=128 beta=1:", round(20 / dt, 1), "it/s", flush=True)
# cfg5
N, C, R = 65536, 4096, 64
torch.manual_seed(0)
V = (torch.rand(N, C).bfloat16().float() + 2 ** -7).cuda()
torch.manual_seed(1); W0 = torch.randn(C, R).abs().cuda(); H0 = torch.randn(N, R).abs().cuda()
m = NMF(W=W0.cpu(), H=H0.cpu()).cuda()
for beta in (0, 0.5, 1, 1.5, 2):
    def run5():
        m.W.data.copy_(W0); m.H.data.copy_(H0); m.fit(V, beta, float("-inf"), 20)
    dt = timed(run5, 2)
    out[f"cfg5_beta{beta}_it_per_s"] = 20 / dt
    print(f"cfg5 beta={beta}: {20/dt:.1f} it/s ({m.last_fit_precision})", flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "other_configs.json"), "w"))
