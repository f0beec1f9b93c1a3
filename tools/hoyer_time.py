"""Sparseness-constrained fit (`sparse_fit`, nmf.py:411-599) and the projection alone: the engine vs the reference on the same
B200 (`.cuda()`, its TorchScript `_proj_func` looped over the components from Python).  One JSON line per case."""
import json, os, sys, time, warnings
warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-nmf_b200")]
import torch
from torchnmf_b200 import NMF, engine
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
import torchnmf.nmf as rn


def tolerr(x, y):
    return float(((x - y).abs() / (1e-3 * y.abs() + 1e-5 * y.abs().max())).max())


def timed(fn, reps):
    fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


# ---- the projection alone: every component of a factor in one launch vs the reference's loop -----------------------------
for rows, R, sp in ((4096, 64, 0.5), (65536, 64, 0.7)):
    torch.manual_seed(0)
    X = (torch.randn(rows, R).abs() + 1e-3).cuda()
    L1 = rows ** 0.5 * (1 - sp) + sp
    norms = (X * X).sum(0).sqrt()
    k1, k2 = L1 * norms, norms * norms
    out = X.clone()

    def ours():
        out.copy_(X)
        engine.hoyer_project_(out, 1, k1, k2)
    t_ours = timed(ours, 5)
    ref_out = X.clone()

    def theirs():
        for j in range(R):
            ref_out[:, j] = rn._proj_func(X[:, j], float(k1[j]), float(k2[j]))
    t_ref = timed(theirs, 1)
    print(json.dumps({"what": "projection of every component", "shape": [rows, R], "sparseness": sp,
                      "engine_ms": round(t_ours * 1e3, 3), "reference_cuda_ms": round(t_ref * 1e3, 1),
                      "speedup": round(t_ref / t_ours, 1), "err_over_tol_vs_reference_cuda": round(tolerr(out, ref_out), 3)}),
          flush=True)

# ---- sparse_fit -----------------------------------------------------------------------------------------------------------
for N, C, R, beta, sW, sH, it_ours, it_ref in ((4096, 1024, 64, 2, 0.5, 0.4, 30, 3), (65536, 4096, 64, 2, 0.5, None, 20, 2)):
    torch.manual_seed(0)
    V = torch.rand(N, C).bfloat16().float().cuda()
    torch.manual_seed(1)
    W0, H0 = torch.randn(C, R).abs(), torch.randn(N, R).abs()

    def run(m, iters):
        m.sparse_fit(V, beta, 1, False, sW, sH)
        m.W.data.copy_(W0); m.H.data.copy_(H0)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        m.sparse_fit(V, beta, iters, False, sW, sH)
        torch.cuda.synchronize()
        return iters / (time.perf_counter() - t0)
    a = NMF(W=W0, H=H0).cuda(); ra = run(a, it_ours)
    c = rn.NMF(W=W0.clone(), H=H0.clone()).cuda(); rc = run(c, it_ref)
    # same number of iterations for the error figure
    b = NMF(W=W0, H=H0).cuda(); b.sparse_fit(V, beta, it_ref, False, sW, sH)
    print(json.dumps({"what": "sparse_fit", "V": [N, C], "R": R, "beta": beta, "sW": sW, "sH": sH,
                      "engine_it_s": round(ra, 2), "engine_precision": a.last_fit_precision,
                      "reference_cuda_it_s": round(rc, 3), "speedup": round(ra / rc, 1),
                      "w_err_over_tol_vs_reference_cuda": round(tolerr(b.W.data, c.W.data), 3),
                      "iters_compared": it_ref}), flush=True)
