cd $GRAFT_REPO_ROOT
TR=$PWD/pytorch-nmf_b200/lib/trace/libnmf_b200.so
NMFB200_TC_KNOCK=58 NMFB200_LIB=$TR ncu --set full --clock-control none --import-source on -k regex:tc_contract -s 3 -c 1 -o gpurun_out/prof_r2b_knock58 -f python tools/profile_target.py f16 > gpurun_out/ncu_r2b.log 2>&1; tail -2 gpurun_out/ncu_r2b.log
