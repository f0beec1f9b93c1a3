"""Time the fused contraction kernels for several pipeline variants (tuning aid); one subprocess per variant."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-nmf_b200")]

def child(prec, v, reps):
    import torch
    from torchnmf_b200.engine import CudaNmfEngine
    N, C, R = 65536, 4096, 64
    torch.manual_seed(0)
    V = torch.rand(N, C, device="cuda").bfloat16().float()
    W = torch.randn(C, R, device="cuda").abs(); H = torch.randn(N, R, device="cuda").abs()
    eng = CudaNmfEngine(V, W, H, prec)
    out = []
    for which in (0, 1):
        for _ in range(3): eng.contract_only(which, 1.0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): eng.contract_only(which, 1.0)
        e1.record(); torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) / reps * 1e3)
    eng.close()
    print(f"{prec} variant {v}: W-contract {out[0]:.1f} us  H-contract {out[1]:.1f} us", flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 3:
        child(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]))
    else:
        todo = [("f16_split", v) for v in (0, 1, 2)] + [("f16", v) for v in (0, 1, 2, 3)]
        if len(sys.argv) > 1:
            todo = [(a.split(":")[0], int(a.split(":")[1])) for a in sys.argv[1].split(",")]
        for prec, v in todo:
            env = dict(os.environ, NMFB200_TC_VARIANT=str(v))
            r = subprocess.run([sys.executable, __file__, prec, str(v), "20"], env=env, capture_output=True, text=True, timeout=120)
            tail = (r.stdout + r.stderr).strip().splitlines()
            keep = [l for l in tail if "variant" in l or "nmf_b200" in l or "Error" in l or "error" in l]
            print("\n".join(keep[-4:]) if keep else f"{prec} variant {v}: rc={r.returncode} (no output)", flush=True)
