"""Parity of every precision mode on the heavy-tailed (lognormal) fixtures of tests/golden/reference_r2.npz, in units of
the test tolerance (rtol 1e-3, atol 1e-5 max)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-nmf_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
from torchnmf_b200 import NMF
Z = np.load(os.path.join(ROOT, "tests", "golden", "reference_r2.npz"))
def case(name): return {k.split("/", 1)[1]: Z[k] for k in Z.files if k.startswith(name + "/")}
for name in sys.argv[1:] or ["nmf_heavy_b1", "nmf_heavy_b0"]:
    c = case(name)
    N, C, R, beta = int(c["N"]), int(c["C"]), int(c["R"]), float(c["beta"])
    heavy = "heavy" in name
    torch.manual_seed(0)
    V = (torch.exp(2.0 * torch.randn(N, C)) if heavy else torch.rand(N, C)).bfloat16().float()
    if "floor" in c and float(c["floor"]) > 0: V = V.clamp_min(float(c["floor"]))
    torch.manual_seed(1)
    W0 = torch.randn(C, R).abs(); H0 = torch.randn(N, R).abs()
    for prec in ("f32", "f16", "f16_split"):
        for iters in (1, 5, int(c["max_iter"])):
            m = NMF(W=W0, H=H0).cuda()
            m.fit(V.cuda(), beta, float("-inf"), iters, precision=prec)
            if iters == int(c["max_iter"]):
                errs = []
                for got, want, mx in ((m.W.data.cpu()[::int(c["w_step"])], torch.from_numpy(c["W_sub"]), float(c["w_absmax"])),
                                      (m.H.data.cpu()[::int(c["h_step"])], torch.from_numpy(c["H_sub"]), float(c["h_absmax"]))):
                    errs.append(float(((got - want).abs() / (1e-3 * want.abs() + 1e-5 * mx)).max()))
                print(f"{name} {prec} [{m.last_fit_precision}] {iters} it: W {errs[0]:.2f} H {errs[1]:.2f} x tol", flush=True)
            else:
                ref = NMF(W=W0, H=H0).cuda(); ref.fit(V.cuda(), beta, float("-inf"), iters, precision="f32")
                e = [float(((a - b).abs() / (1e-3 * b.abs() + 1e-5 * b.abs().max())).max()) for a, b in ((m.W.data, ref.W.data), (m.H.data, ref.H.data))]
                print(f"{name} {prec} {iters} it vs f32 engine: W {e[0]:.2f} H {e[1]:.2f} x tol", flush=True)
