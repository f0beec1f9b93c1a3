cd $GRAFT_REPO_ROOT
TR=$PWD/pytorch-nmf_b200/lib/trace/libnmf_b200.so
( NMFB200_LIB=$TR python tools/tc_knock.py f16 0,24,56,58,26,0 ) 2>&1 | tail -6
