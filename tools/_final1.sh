# Round-2 evidence run (1 GPU): sanitizer logs, stress runs, ncu captures, launch lists, bench lines.  Everything lands in
# gpurun_out/ and is summarised into profiles/ by tools/collect_profiles.py afterwards.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
# --- sanitizers on every ring configuration of the fused kernel + the NMFD kernels (small shapes) ---
cat > /tmp/small_nmfd.py <<'PY'
import os, sys
sys.path[:0] = [os.getcwd(), os.path.join(os.getcwd(), "pytorch-nmf_b200")]
import torch
from torchnmf_b200 import NMFD
torch.manual_seed(0)
V = torch.rand(2, 130, 700).cuda()
m = NMFD((2, 130, 700), 5, 37).cuda()
print("nmfd", m.fit(V, 1, float("-inf"), 3), m.last_fit_precision, float(m.W.data.sum()))
PY
timeout 400 compute-sanitizer --tool racecheck --racecheck-report all python tools/small_fit.py > gpurun_out/r2_racecheck_nmf.log 2>&1; tail -4 gpurun_out/r2_racecheck_nmf.log
timeout 300 compute-sanitizer --tool racecheck python /tmp/small_nmfd.py > gpurun_out/r2_racecheck_nmfd.log 2>&1; tail -3 gpurun_out/r2_racecheck_nmfd.log
timeout 200 compute-sanitizer --tool synccheck python tools/small_fit.py > gpurun_out/r2_synccheck_nmf.log 2>&1; tail -2 gpurun_out/r2_synccheck_nmf.log
timeout 200 compute-sanitizer --tool synccheck python /tmp/small_nmfd.py > gpurun_out/r2_synccheck_nmfd.log 2>&1; tail -2 gpurun_out/r2_synccheck_nmfd.log
timeout 200 compute-sanitizer --tool memcheck python /tmp/small_nmfd.py > gpurun_out/r2_memcheck_nmfd.log 2>&1; tail -2 gpurun_out/r2_memcheck_nmfd.log
# --- 2000-launch stress of the NV = 3 ring configurations (fresh process each) ---
( python tools/tc_stress.py f16_split 8192 4096 2000; python tools/tc_stress.py f16 8192 4096 2000 ) > gpurun_out/r2_stress.log 2>&1
python - >> gpurun_out/r2_stress.log 2>&1 <<'PY'
import os, sys, time
sys.path[:0] = [os.getcwd(), os.path.join(os.getcwd(), "pytorch-nmf_b200")]
import torch
from torchnmf_b200.engine import CudaNmfEngine
torch.manual_seed(0)
N, C, R = 16384, 4096, 128
V = torch.rand(N, C, device="cuda").bfloat16().float(); W = torch.randn(C, R, device="cuda").abs(); H = torch.randn(N, R, device="cuda").abs()
for prec in ("f16", "f16_split"):
    eng = CudaNmfEngine(V, W, H, prec); t0 = time.time()
    for i in range(2000):
        eng.contract_only(i & 1, 1.0)
        if i % 200 == 199: eng.check_health()
    eng.check_health(); eng.close()
    print(f"OK   {prec} R=128 N={N} C={C} 2000 launches {time.time()-t0:.2f}s", flush=True)
PY
cat gpurun_out/r2_stress.log
