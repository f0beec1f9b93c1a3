cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:tc_contract -s 4 -c 2 -o gpurun_out/prof_r2a_tc_f16 -f python tools/profile_target.py f16 > gpurun_out/ncu_r2a.log 2>&1; tail -3 gpurun_out/ncu_r2a.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r2a.csv python tools/profile_target.py f16 > /dev/null 2>&1; tail -30 gpurun_out/launches_r2a.csv | cut -c1-200
python -m pytest tests/test_gpu_r2.py tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -5
