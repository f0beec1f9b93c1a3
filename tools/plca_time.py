"""PLCA family (plca.py:193-625): EM iterations per second, the engine vs the reference on the same B200 (`.cuda()`).  One JSON
line per model; the error figure compares both after the same (small) number of iterations, in units of the 1e-3 tolerance."""
import json, os, sys, time, warnings
warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-nmf_b200")]
import torch
import torchnmf_b200 as ours
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
import torchnmf.plca as rp


def tolerr(x, y):
    return float(((x - y).abs() / (1e-3 * y.abs() + 1e-5 * y.abs().max())).max())


CASES = [
    ("PLCA", dict(V=(8192, 2048), W=(2048, 64), H=(8192, 64)), 50, 10),
    ("SIPLCA", dict(V=(1, 513, 2048), W=(513, 16, 32), H=(1, 16, 2017)), 30, 5),
    ("SIPLCA2", dict(V=(1, 16, 64, 512), W=(16, 8, 4, 16), H=(1, 8, 61, 497)), 30, 5),
]
for name, sh, it_ours, it_ref in CASES:
    torch.manual_seed(0)
    V = (torch.rand(*sh["V"]).bfloat16().float() * 3).cuda()
    torch.manual_seed(1)
    W0, H0 = torch.randn(*sh["W"]).abs(), torch.randn(*sh["H"]).abs()
    Z0 = torch.rand(sh["W"][1]) + 0.1

    def run(mod, iters, **kw):
        m = getattr(mod, name)(W=W0.clone(), H=H0.clone(), Z=Z0.clone()).cuda()
        m.fit(V, float("-inf"), 1, **kw)
        m = getattr(mod, name)(W=W0.clone(), H=H0.clone(), Z=Z0.clone()).cuda()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        m.fit(V, float("-inf"), iters, **kw)
        torch.cuda.synchronize()
        return m, iters / (time.perf_counter() - t0)
    a, ra = run(ours, it_ours)
    c, rc = run(rp, it_ref)
    b, _ = run(ours, it_ref)
    print(json.dumps({"model": name, "V": list(sh["V"]), "R": sh["W"][1], "kernel": list(sh["W"][2:]),
                      "engine_it_s": round(ra, 1), "reference_cuda_it_s": round(rc, 2), "speedup": round(ra / rc, 1),
                      "w_err_over_tol_vs_reference_cuda": round(tolerr(b.W.data, c.W.data), 3), "iters_compared": it_ref}),
          flush=True)
