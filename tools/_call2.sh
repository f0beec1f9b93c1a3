cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
R1=$PWD/pytorch-nmf_b200/lib/r1/libnmf_b200.so
TR=$PWD/pytorch-nmf_b200/lib/trace/libnmf_b200.so
export NMFB200_LIB_COMPAT=1
( NMFB200_LIB=$R1 python tools/tc_time.py f16; python tools/tc_time.py f16; NMFB200_LIB=$R1 python tools/tc_time.py f16; python tools/tc_time.py f16; NMFB200_FUSED_TAIL=0 python tools/tc_time.py f16 ) 2>&1 | grep -E "lib=|rror"
( NMFB200_LIB=$TR python tools/tc_knock.py f16 0,4,8,16,24,32,36,60,0 ) 2>&1 | tail -9
( NMFB200_LIB=$TR python tools/tc_trace.py f16 1 ) > gpurun_out/trace_h.txt 2>&1; head -45 gpurun_out/trace_h.txt
( python tools/tc_time.py f16 131072 8192 128; NMFB200_LIB=$R1 python tools/tc_time.py f16 131072 8192 128; python tools/tc_time.py f16_split; NMFB200_LIB=$R1 python tools/tc_time.py f16_split ) 2>&1 | grep -E "lib=|rror"
python tools/diag_heavy.py 2>&1 | tail -30
NMFB200_LIB=$R1 python tools/diag_heavy.py 2>&1 | grep "30 it"
timeout 1000 python -m pytest tests -m gpu -q -x > gpurun_out/pytest2.txt 2>&1; tail -5 gpurun_out/pytest2.txt
