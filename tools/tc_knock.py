"""Knock-out timing of the fused contraction kernel (tuning build: NMFB200_BUILD_TRACE=1 python pytorch-nmf_b200/build.py --force).

Skips one pipeline stage at a time (results are invalid) to see which stage bounds the kernel:
  1 no MUFU reciprocal | 2 no V shared-memory reads | 4 ratio = TMEM load + store only | 8 no O-MMA | 16 no S-MMA |
  32 no V TMA loads (no HBM stream)
Staged variants with VALID results (A/B them inside one gpurun call): 64 a quarter of the reciprocals on the FMA pipe |
  128 start the next tile's first chunk load before this tile's last chunk
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("NMFB200_LIB", os.path.join(ROOT, "pytorch-nmf_b200", "lib", "trace", "libnmf_b200.so"))   # tuning build
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-nmf_b200")]
import torch
from torchnmf_b200.engine import CudaNmfEngine, release_workspaces
prec = sys.argv[1] if len(sys.argv) > 1 else "f16"
N, C, R = 65536, 4096, 64
torch.manual_seed(0)
V = torch.rand(N, C, device="cuda").bfloat16().float()
W = torch.randn(C, R, device="cuda").abs(); H = torch.randn(N, R, device="cuda").abs()
ks = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1, 2, 3, 4, 8, 16, 24, 32, 33, 35, 36, 40, 56, 60, 63, 0, 64, 128, 192, 0]
for k in ks:
    os.environ["NMFB200_TC_KNOCK"] = str(k)       # read once per engine context
    eng = CudaNmfEngine(V, W, H, prec)
    out = []
    for which in (0, 1):
        for _ in range(3): eng.contract_only(which, 1.0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): eng.contract_only(which, 1.0)
        e1.record(); torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) / 10 * 1e3)
    print(f"knock {k:2d}: W {out[0]:7.1f} us  H {out[1]:7.1f} us", flush=True)
    eng.close(); release_workspaces()
