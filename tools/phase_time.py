"""Where an MU iteration's time goes at cfg2: contraction alone, update (contraction + ratio stage), whole iteration, all
timed in-stream with CUDA events (launch gaps included), against the sum of the kernels' own durations."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-nmf_b200")]
import torch
from torchnmf_b200.engine import CudaNmfEngine, release_workspaces
prec = sys.argv[1] if len(sys.argv) > 1 else "f16"
N, C, R = (int(x) for x in sys.argv[2:5]) if len(sys.argv) > 4 else (65536, 4096, 64)
torch.manual_seed(0)
V = torch.rand(N, C, device="cuda").bfloat16().float()
W = torch.randn(C, R, device="cuda").abs(); H = torch.randn(N, R, device="cuda").abs()
eng = CudaNmfEngine(V, W, H, prec)
def t(fn, n=40):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
cw = t(lambda: eng.contract_only(0, 1.0)); ch = t(lambda: eng.contract_only(1, 1.0))
uw = t(lambda: eng.update_w(1.0, 1.0, 0.0, 0.0)); uh = t(lambda: eng.update_h(1.0, 1.0, 0.0, 0.0))
def both(): eng.update_w(1.0, 1.0, 0.0, 0.0); eng.update_h(1.0, 1.0, 0.0, 0.0)
it = t(both)
it_native = t(lambda: eng.iterate(20, 1.0, 1.0, 0.0, 0.0), 5) / 20
print(f"contract W {cw:.1f} H {ch:.1f} | update W {uw:.1f} H {uh:.1f} (ratio stage + gap: W {uw - cw:.1f} H {uh - ch:.1f}) | "
      f"W+H {it:.1f} | native loop {it_native:.1f} us/it", flush=True)
long_both = t(both, 400)
def ten_sync():
    for _ in range(10): both()
    torch.cuda.synchronize()
ts = t(ten_sync, 40) / 10
def ten_loss():
    for _ in range(10): both()
    eng.loss(1.0)
tl = t(ten_loss, 40) / 10
def ten_native_loss():
    eng.iterate(10, 1.0, 1.0, 0.0, 0.0)
    eng.loss(1.0)
tn = t(ten_native_loss, 40) / 10
print(f"sustained (400 it) W+H {long_both:.1f} | 10 it + sync {ts:.1f} | 10 it + loss {tl:.1f} | iterate(10) + loss {tn:.1f} us/it", flush=True)
eng.check_health(); eng.close(); release_workspaces()
