# round-2 closing evidence run (one gpurun call): tests, smoke, bench lines, timing tools, launch lists, one ncu capture
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r2_final_pytest_gpu.txt; tail -1 gpurun_out/r2_final_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.txt 2>&1; tail -2 gpurun_out/r2_final_smoke.txt
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_final_bench_cfg2.json 2> gpurun_out/r2_final_bench_cfg2.err; cut -c1-120 gpurun_out/r2_final_bench_cfg2.json
python bench.py --config cfg3 --steps 3 --warmup 3 > gpurun_out/r2_final_bench_cfg3.json 2> /dev/null; cut -c1-120 gpurun_out/r2_final_bench_cfg3.json
python bench.py --config cfg5 --beta 2 --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_final_bench_cfg5_beta2.json 2> /dev/null; cut -c1-120 gpurun_out/r2_final_bench_cfg5_beta2.json
timeout 90 python tools/plca_time.py > gpurun_out/r2_plca_time.txt 2>&1; tail -3 gpurun_out/r2_plca_time.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2_final_launches_cfg2_f16.csv python tools/profile_target.py f16 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:prep_w -s 2 -c 1 -o gpurun_out/prof_prep_w_final -f python tools/nmfd_prof.py > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/prof_prep_w_final.ncu-rep > gpurun_out/r2_ncu_prep_w.txt 2>&1; head -12 gpurun_out/r2_ncu_prep_w.txt
