cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for i in 1 2; do
NMFB200_LIB=$PWD/pytorch-nmf_b200/lib/prev/libnmf_b200.so NMFB200_LIB_COMPAT=1 python tools/tc_time.py f16 2>&1 | tail -1 | cut -c40-
python tools/tc_time.py f16 2>&1 | tail -1 | cut -c1-200
done
python bench.py --steps 5 --warmup 3 > gpurun_out/r2b_bench_cfg2_f16.json 2>/dev/null; cut -c1-220 gpurun_out/r2b_bench_cfg2_f16.json
