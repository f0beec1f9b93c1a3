cd $GRAFT_REPO_ROOT
ncu --set full --clock-control none --import-source on -s 2 -c 1 -o gpurun_out/prof_ubench_ratio -f tools/ubench/ratio > gpurun_out/ncu_ub.log 2>&1; tail -2 gpurun_out/ncu_ub.log
