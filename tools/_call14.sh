cd $GRAFT_REPO_ROOT
( python tools/tc_time.py f16 ) 2>&1 | grep -E "lib=|rror"
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r2c.csv python tools/profile_target.py f16 > /dev/null 2>&1; grep -E "apply_finish|tc_contract" gpurun_out/launches_r2c.csv | tail -4 | awk -F'","' '{print substr($5,1,60), $(NF)}'
python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
