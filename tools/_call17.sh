cd $GRAFT_REPO_ROOT
mkdir -p /tmp; cat > /tmp/nmfd_prof.py <<'PY'
import os, sys
sys.path[:0] = [os.getcwd(), os.path.join(os.getcwd(), "pytorch-nmf_b200")]
import torch
from torchnmf_b200.engine import CudaNmfdEngine
torch.manual_seed(0)
V = torch.rand(1, 1025, 8192).bfloat16().float().cuda()
W = torch.randn(1025, 16, 128).abs().cuda(); H = torch.randn(1, 16, 8065).abs().cuda()
eng = CudaNmfdEngine(V, W, H, "f16")
for _ in range(3):
    eng.update_w(1, 1.0, 0.0, 0.0); eng.update_h(1, 1.0, 0.0, 0.0)
print(eng.loss(1))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_nmfd.csv python /tmp/nmfd_prof.py > /dev/null 2>&1
python3 - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/launches_nmfd.csv')))
hdr=[r for r in rows if r and r[0]=='ID'][0]; ix={h:i for i,h in enumerate(hdr)}
out=[]
for r in rows:
    if len(r)==len(hdr) and r[0].isdigit() and 'nmfb200' in r[ix['Kernel Name']]:
        n=r[ix['Kernel Name']].split('(')[0].split('::')[-1][:38]
        out.append((n, r[ix['Grid Size']], float(r[ix['Metric Value']])/1000))
for o in out[-26:]: print("%-40s %-16s %8.1f us"%o)
PY
timeout 300 python -m pytest tests/test_gpu_r2.py tests/test_gpu_parity.py -m gpu -q -x -k "nmfd or NMFD" 2>&1 | tail -3
