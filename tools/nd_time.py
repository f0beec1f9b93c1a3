"""NMF2D / NMF3D: engine vs the unmodified reference on the same B200 (.cuda()) and on the host cores, same inputs.
There is no BASELINE.json config for these models; the workloads are the reference's docstring examples scaled up."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-nmf_b200")]
import torch
from torchnmf_b200 import NMF2D, NMF3D
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
import torchnmf.nmf as rn

WORK = [("nmf2d", NMF2D, rn.NMF2D, (1, 32, 128, 2048), 8, (4, 16)),
        ("nmf3d", NMF3D, rn.NMF3D, (1, 3, 64, 64, 100), 8, (5, 5, 20))]      # nmf.py:914-922
for name, cls, rcls, vs, R, K in WORK:
    for beta in (1.0, 2.0):
        torch.manual_seed(0)
        V = torch.rand(*vs)
        m0 = cls(vs, R, K)
        W0, H0 = m0.W.data.clone(), m0.H.data.clone()
        flops = 4 * 2.0 * vs[0] * vs[1] * R * m0.W[0, 0].numel() * m0.H[0, 0].numel() * (1.0 if beta == 1 else 1.5)
        def run(mod, Vx, iters, **kw):
            mod.fit(Vx, beta, float("-inf"), 2, **kw)
            mod.W.data.copy_(W0); mod.H.data.copy_(H0)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            mod.fit(Vx, beta, float("-inf"), iters, **kw)
            torch.cuda.synchronize()
            return iters / (time.perf_counter() - t0)
        eng = cls(W=W0.clone(), H=H0.clone()).cuda()
        r_eng = run(eng, V.cuda(), 20)
        rates = {}
        for tf32 in (True, False):                     # cuDNN's default convolution arithmetic is TF32
            torch.backends.cudnn.allow_tf32 = tf32
            ref = rcls(W=W0.clone(), H=H0.clone()).cuda()
            rates[tf32] = run(ref, V.cuda(), 20)
        torch.backends.cudnn.allow_tf32 = True
        cpu = rcls(W=W0.clone(), H=H0.clone())
        cpu.fit(V, beta, float("-inf"), 2)
        cpu.W.data.copy_(W0); cpu.H.data.copy_(H0)
        t0 = time.perf_counter(); cpu.fit(V, beta, float("-inf"), 20); r_cpu = 20 / (time.perf_counter() - t0)
        def tolerr(a, b):
            return float(((a - b).abs() / (1e-3 * b.abs() + 1e-5 * b.abs().max())).max())
        err = tolerr(eng.W.data.cpu(), cpu.W.data)                     # 20 iterations each, against the reference's CPU fit
        err_ref = tolerr(ref.W.data.cpu(), cpu.W.data)                 # the reference's own fp32 GPU fit against its CPU fit
        print(json.dumps({"model": name, "V": vs, "R": R, "kernel_size": K, "beta": beta, "engine_it_s": round(r_eng, 2),
                          "reference_cuda_tf32_it_s": round(rates[True], 2), "reference_cuda_fp32_it_s": round(rates[False], 2),
                          "reference_cpu_it_s": round(r_cpu, 3), "engine_tflops": round(flops * r_eng / 1e12, 2),
                          "engine_w_err_over_tol_vs_reference_cpu": round(err, 3),
                          "reference_cuda_fp32_w_err_over_tol_vs_reference_cpu": round(err_ref, 3),
                          "precision": eng.last_fit_precision}), flush=True)
