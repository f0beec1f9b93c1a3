cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TR=$PWD/pytorch-nmf_b200/lib/trace/libnmf_b200.so
R1=$PWD/pytorch-nmf_b200/lib/r1/libnmf_b200.so
export NMFB200_LIB_COMPAT=1
( NMFB200_TC_NRW=2 python tools/tc_time.py f16; python tools/tc_time.py f16; NMFB200_LIB=$R1 python tools/tc_time.py f16; NMFB200_TC_NRW=2 python tools/tc_time.py f16; python tools/tc_time.py f16 ) 2>&1 | grep -E "lib=|rror"
( NMFB200_TC_NRW=2 NMFB200_LIB=$TR python tools/tc_knock.py f16 0,8,16,24,32,0 ) 2>&1 | tail -6; ( NMFB200_TC_NRW=2 NMFB200_LIB=$TR python tools/tc_trace.py f16 1 ) > gpurun_out/trace_h5.txt 2>&1; sed -n 22,34p gpurun_out/trace_h5.txt
( python tools/tc_time.py f16 131072 8192 128; python tools/tc_time.py f16_split ) 2>&1 | grep -E "lib=|rror"
python -m pytest tests/test_gpu_parity.py tests/test_gpu_r2.py -m gpu -q -x 2>&1 | tail -3
