"""Ring-depth experiment (tuning build): time the f16 contraction for pipeline variants x L2-prefetch distance."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("NMFB200_LIB", os.path.join(ROOT, "pytorch-nmf_b200", "lib", "trace", "libnmf_b200.so"))   # tuning build
CHILD = r'''
import os, sys
sys.path[:0] = [%r, os.path.join(%r, "pytorch-nmf_b200")]
import torch
from torchnmf_b200.engine import CudaNmfEngine
N, C, R = 65536, 4096, 64
torch.manual_seed(0)
V = torch.rand(N, C, device="cuda").bfloat16().float()
W = torch.randn(C, R, device="cuda").abs(); H = torch.randn(N, R, device="cuda").abs()
eng = CudaNmfEngine(V, W, H, "f16")
out = []
for which in (0, 1):
    for _ in range(3): eng.contract_only(which, 1.0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): eng.contract_only(which, 1.0)
    e1.record(); torch.cuda.synchronize()
    out.append(e0.elapsed_time(e1) / 10 * 1e3)
eng.check_health()
print("W %%7.1f us  H %%7.1f us" %% tuple(out))
''' % (ROOT, ROOT)
for variant in sys.argv[1].split(","):
    for pf in sys.argv[2].split(","):
        for knock in (sys.argv[3].split(",") if len(sys.argv) > 3 else ["0"]):
            env = dict(os.environ, NMFB200_TC_VARIANT=variant, NMFB200_TC_PF=pf, NMFB200_TC_KNOCK=knock)
            r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=120)
            print(f"variant {variant} pf {pf} knock {knock}: {r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr.strip()[-300:]}", flush=True)
