cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( python tools/tc_time.py f16; NMFB200_FUSED_TAIL=0 python tools/tc_time.py f16; python tools/tc_time.py f16 ) 2>&1 | grep -E "lib=|rror"
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r2b.csv python tools/profile_target.py f16 > /dev/null 2>&1; grep -E "apply_finish|tc_contract" gpurun_out/launches_r2b.csv | tail -6 | awk -F'","' '{print substr($5,1,60), $(NF)}'
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 300 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; tail -c 3000 gpurun_out/bench_r2a.json; tail -3 gpurun_out/bench_r2a.err
timeout 240 compute-sanitizer --tool synccheck python tools/small_fit.py > gpurun_out/synccheck.log 2>&1; tail -12 gpurun_out/synccheck.log
