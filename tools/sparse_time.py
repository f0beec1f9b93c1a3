"""Sparse-target NMF: the library's sparse kernels vs the densified path vs the reference's sparse path on the same B200."""
import json, os, sys, time, warnings
warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-nmf_b200")]
import torch
from torchnmf_b200 import NMF
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
import torchnmf.nmf as rn
N, C, R, dens = 65536, 4096, 64, 0.01
torch.manual_seed(0)
nnz = int(N * C * dens)
idx = torch.stack([torch.randint(0, N, (nnz,)), torch.randint(0, C, (nnz,))])
V = torch.sparse_coo_tensor(idx, torch.rand(nnz) + 0.05, (N, C)).coalesce().cuda()
torch.manual_seed(1)
W0, H0 = torch.randn(C, R).abs(), torch.randn(N, R).abs()
for beta in (1, 2):
    def run(m, iters, **kw):
        m.fit(V, beta, float("-inf"), 2, **kw)
        m.W.data.copy_(W0); m.H.data.copy_(H0)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        m.fit(V, beta, float("-inf"), iters, **kw)
        torch.cuda.synchronize()
        return iters / (time.perf_counter() - t0)
    a = NMF(W=W0, H=H0).cuda(); ra = run(a, 50)
    b = NMF(W=W0, H=H0).cuda(); b._sparse_kernels = False; rb = run(b, 50)
    c = rn.NMF(W=W0.clone(), H=H0.clone()).cuda(); rc = run(c, 10)
    def tolerr(x, y): return float(((x - y).abs() / (1e-3 * y.abs() + 1e-5 * y.abs().max())).max())
    print(json.dumps({"V": [N, C], "nnz": int(V._nnz()), "R": R, "beta": beta, "sparse_kernels_it_s": round(ra, 1),
                      "densified_it_s": round(rb, 1), "densified_precision": b.last_fit_precision,
                      "reference_cuda_sparse_it_s": round(rc, 2),
                      "sparse_vs_densified_w_err_over_tol": round(tolerr(a.W.data, b.W.data), 3)}), flush=True)
