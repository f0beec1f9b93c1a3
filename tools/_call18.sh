cd $GRAFT_REPO_ROOT
ncu --set full --clock-control none --import-source on -k regex:recon2 -s 2 -c 1 -o gpurun_out/prof_r2_recon2 -f python /tmp/nmfd_prof.py > gpurun_out/ncu_recon2.log 2>&1; tail -2 gpurun_out/ncu_recon2.log
