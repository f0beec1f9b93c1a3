cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
cat > /tmp/small_nmfd.py <<PY
import os, sys
sys.path[:0] = [os.getcwd(), os.path.join(os.getcwd(), "pytorch-nmf_b200")]
import torch
from torchnmf_b200 import NMFD
torch.manual_seed(0)
V = torch.rand(2, 130, 700).cuda()
m = NMFD((2, 130, 700), 5, 37).cuda()
print("nmfd", m.fit(V, 1, float("-inf"), 3), m.last_fit_precision, float(m.W.data.sum()))
PY
timeout 300 compute-sanitizer --tool racecheck python /tmp/small_nmfd.py > gpurun_out/r2_racecheck_nmfd.log 2>&1 || true
[ -f /tmp/small_nmfd.py ] || true
tail -2 gpurun_out/r2_racecheck_nmfd.log
ncu --set full --clock-control none --import-source on -k regex:tc_contract -s 4 -c 2 -o gpurun_out/prof_r2_tc_contract_f16 -f python tools/profile_target.py f16 > gpurun_out/ncu_r2_tc.log 2>&1; tail -1 gpurun_out/ncu_r2_tc.log
ncu --set full --clock-control none --import-source on -k regex:tcnmfd -s 3 -c 4 -o gpurun_out/prof_r2_nmfd -f python tools/nmfd_prof.py > gpurun_out/ncu_r2_nmfd.log 2>&1; tail -1 gpurun_out/ncu_r2_nmfd.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2_launches_cfg2_f16.csv python tools/profile_target.py f16 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2_launches_cfg3_nmfd.csv python tools/nmfd_prof.py > /dev/null 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_cfg2_f16.json 2> gpurun_out/r2_bench_cfg2_f16.err; cut -c1-300 gpurun_out/r2_bench_cfg2_f16.json
python bench.py --steps 3 --warmup 3 --precision f16_split --no-extras > gpurun_out/r2_bench_cfg2_f16_split.json 2>/dev/null; cut -c1-200 gpurun_out/r2_bench_cfg2_f16_split.json
python bench.py --config cfg3 --steps 3 --warmup 3 > gpurun_out/r2_bench_cfg3.json 2>/dev/null; cut -c1-200 gpurun_out/r2_bench_cfg3.json
python bench.py --config cfg4s --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_cfg4s.json 2>/dev/null; cut -c1-200 gpurun_out/r2_bench_cfg4s.json
for b in 0 0.5 1 1.5 2; do python bench.py --config cfg5 --beta $b --steps 3 --warmup 3 > gpurun_out/r2_bench_cfg5_beta$b.json 2>/dev/null; cut -c1-160 gpurun_out/r2_bench_cfg5_beta$b.json; done
