cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TR=$PWD/pytorch-nmf_b200/lib/trace/libnmf_b200.so
( NMFB200_TC_NRW=2 python tools/tc_time.py f16; python tools/tc_time.py f16; NMFB200_TC_NRW=2 python tools/tc_time.py f16; python tools/tc_time.py f16 ) 2>&1 | grep -E "lib=|rror"
( NMFB200_LIB=$TR python tools/tc_knock.py f16 0,4,8,16,24,32,36,60,0 ) 2>&1 | tail -9
( NMFB200_LIB=$TR python tools/tc_trace.py f16 1 ) > gpurun_out/trace_h4.txt 2>&1; sed -n 20,40p gpurun_out/trace_h4.txt
python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
