"""Time the fused contraction kernels and one 200-iteration fit at cfg2 (or another shape) with the library named by
NMFB200_LIB (A/B of two builds inside one gpurun call: same box, same clocks).

    python tools/tc_time.py [precision] [N C R] [beta]
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-nmf_b200")]
import torch
from torchnmf_b200 import NMF, _capi
from torchnmf_b200.engine import CudaNmfEngine, release_workspaces
prec = sys.argv[1] if len(sys.argv) > 1 else "f16"
N, C, R = (int(x) for x in sys.argv[2:5]) if len(sys.argv) > 4 else (65536, 4096, 64)
beta = float(sys.argv[5]) if len(sys.argv) > 5 else 1.0
torch.manual_seed(0)
V = torch.rand(N, C, device="cuda").bfloat16().float()
if beta <= 0:
    V.clamp_(min=2 ** -7)
W = torch.randn(C, R, device="cuda").abs(); H = torch.randn(N, R, device="cuda").abs()
eng = CudaNmfEngine(V, W, H, prec)
out = []
for which in (0, 1):
    for _ in range(3): eng.contract_only(which, beta)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): eng.contract_only(which, beta)
    e1.record(); torch.cuda.synchronize()
    out.append(e0.elapsed_time(e1) / 20 * 1e3)
eng.check_health(); eng.close()
m = NMF(W=W.cpu(), H=H.cpu()).cuda()
W0, H0 = m.W.data.clone(), m.H.data.clone()
iters = 200
for _ in range(2):
    m.W.data.copy_(W0); m.H.data.copy_(H0)
    m.fit(V, beta, float("-inf"), iters, precision=prec)
torch.cuda.synchronize()
l0 = _capi.launch_count()
t0 = time.perf_counter()
for _ in range(3):
    m.W.data.copy_(W0); m.H.data.copy_(H0)
    m.fit(V, beta, float("-inf"), iters, precision=prec)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 3
print(f"lib={os.environ.get('NMFB200_LIB', 'default')} prec={m.last_fit_precision} {N}x{C} R={R} beta={beta}: "
      f"contract W {out[0]:.1f} us H {out[1]:.1f} us | fit {iters} it: {dt * 1e3:.1f} ms = {iters / dt:.0f} it/s "
      f"({dt / iters * 1e6:.1f} us/it, {(_capi.launch_count() - l0) / 3 / iters:.1f} launches/it)", flush=True)
release_workspaces()
