"""Dump the per-tile event timeline of CTA 0 of the fused contraction kernel (tuning aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("NMFB200_LIB", os.path.join(ROOT, "pytorch-nmf_b200", "lib", "trace", "libnmf_b200.so"))   # tuning build
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-nmf_b200")]
prec = sys.argv[1] if len(sys.argv) > 1 else "f16"
which = int(sys.argv[2]) if len(sys.argv) > 2 else 1
os.environ["NMFB200_TC_TRACE"] = os.path.join(ROOT, "gpurun_out", f"trace_{prec}_{which}.txt")
if len(sys.argv) > 3: os.environ["NMFB200_TC_VARIANT"] = sys.argv[3]
import torch
from torchnmf_b200.engine import CudaNmfEngine
N, C, R = 65536, 4096, 64
torch.manual_seed(0)
V = torch.rand(N, C, device="cuda").bfloat16().float()
W = torch.randn(C, R, device="cuda").abs(); H = torch.randn(N, R, device="cuda").abs()
eng = CudaNmfEngine(V, W, H, prec)
for _ in range(3): eng.contract_only(which, 1.0)
torch.cuda.synchronize()
rows = [[int(x) for x in l.split()] for l in open(os.environ["NMFB200_TC_TRACE"])]
t0 = min(x for r in rows for x in r if x > 0)
names = ["S:wait_g", "S:got_g", "R:prewait", "R:gotV", "R:gotS", "O:wait_p", "O:got_p", "V:issue", "G:issue", "R:end00", "R:lastcmp", "R:endN3", "S:gotG", "S:issued", "O:issued", "O:mma0"]
print("tile " + " ".join(n.rjust(9) for n in names))
for i, r in enumerate(rows[:64]):
    if not any(r): break
    print(f"{i:4d} " + " ".join((str(x - t0) if x else "-").rjust(9) for x in r[:16]))
