"""Quick check of the tensor-core NMFD path against the fp32 path of the same engine (tiny and cfg3 shapes) + timing."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-nmf_b200")]
import torch
from torchnmf_b200 import NMFD
from torchnmf_b200.engine import CudaNmfdEngine
def run(B, C, L, R, T, iters, seed=0):
    torch.manual_seed(seed)
    V = torch.rand(B, C, L).bfloat16().float().cuda()
    W0 = torch.randn(C, R, T).abs().cuda(); H0 = torch.randn(B, R, L - T + 1).abs().cuda()
    outs = {}
    for prec in ("f32", "f16"):
        W, H = W0.clone(), H0.clone()
        eng = CudaNmfdEngine(V, W, H, prec)
        l0 = eng.loss(1)
        for _ in range(iters):
            eng.update_w(1, 1.0, 0.0, 0.0); eng.update_h(1, 1.0, 0.0, 0.0)
        l1 = eng.loss(1)
        eng.check_health(); eng.close()
        outs[prec] = (W, H, l0, l1)
    eW = ((outs["f16"][0] - outs["f32"][0]).abs() / (1e-3 * outs["f32"][0].abs() + 1e-5 * outs["f32"][0].abs().max())).max().item()
    eH = ((outs["f16"][1] - outs["f32"][1]).abs() / (1e-3 * outs["f32"][1].abs() + 1e-5 * outs["f32"][1].abs().max())).max().item()
    print(f"B{B} C{C} L{L} R{R} T{T} {iters} it: tc vs f32: W {eW:.3f} H {eH:.3f} x tol | loss f32 {outs['f32'][2]:.6g}->{outs['f32'][3]:.6g} tc {outs['f16'][2]:.6g}->{outs['f16'][3]:.6g}", flush=True)
run(1, 40, 70, 4, 6, 1)
run(1, 40, 70, 4, 6, 5)
run(2, 130, 700, 5, 37, 5)
run(1, 300, 1000, 16, 128, 3)
run(1, 1025, 8192, 16, 128, 5)
# timing at cfg3
torch.manual_seed(0)
V = torch.rand(1, 1025, 8192).bfloat16().float().cuda()
for prec in ("f32", "f16"):
    m = NMFD((1, 1025, 8192), 16, 128).cuda()
    m.fit(V, 1, float("-inf"), 20, precision=prec)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    n = m.fit(V, 1, float("-inf"), 100, precision=prec)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"cfg3 {prec} [{m.last_fit_precision}]: {n / dt:.0f} it/s ({dt / n * 1e6:.0f} us/it)", flush=True)
