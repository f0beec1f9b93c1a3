import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-nmf_b200")]
import torch
from torchnmf_b200.engine import CudaNmfEngine
prec, N, C, which = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
R = 64
torch.manual_seed(0)
V = torch.rand(N, C, device="cuda").bfloat16().float()
W = torch.randn(C, R, device="cuda").abs(); H = torch.randn(N, R, device="cuda").abs()
eng = CudaNmfEngine(V, W, H, prec)
torch.cuda.synchronize()
t0 = time.time()
try:
    for i in range(5):
        eng.contract_only(which, 1.0)
        torch.cuda.synchronize()
    print(f"OK   {prec} var={os.environ.get('NMFB200_TC_VARIANT')} N={N} C={C} which={which} {time.time()-t0:.3f}s", flush=True)
except Exception as e:
    print(f"FAIL {prec} var={os.environ.get('NMFB200_TC_VARIANT')} N={N} C={C} which={which} after {time.time()-t0:.3f}s: {str(e)[:60]}", flush=True)
