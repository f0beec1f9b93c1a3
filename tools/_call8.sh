cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TR=$PWD/pytorch-nmf_b200/lib/trace/libnmf_b200.so
( NMFB200_TC_PSEP=0 NMFB200_LIB=$TR python tools/tc_trace.py f16 1 ) > gpurun_out/trace_h8a.txt 2>&1; sed -n 1,1p gpurun_out/trace_h8a.txt; sed -n 24,40p gpurun_out/trace_h8a.txt
( NMFB200_TC_PSEP=0 NMFB200_TC_KNOCK=32 NMFB200_LIB=$TR python tools/tc_trace.py f16 1 ) > gpurun_out/trace_h8b.txt 2>&1; sed -n 24,36p gpurun_out/trace_h8b.txt
