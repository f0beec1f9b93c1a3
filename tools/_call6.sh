cd $GRAFT_REPO_ROOT
TR=$PWD/pytorch-nmf_b200/lib/trace/libnmf_b200.so
export NMFB200_TC_NRW=2 NMFB200_LIB=$TR
for park in 0 1 20 50 100 200 0; do echo "park=$park"; NMFB200_TC_PARK=$park python tools/tc_knock.py f16 0,24 2>&1 | tail -2; done
