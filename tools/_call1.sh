set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv
R1=$PWD/pytorch-nmf_b200/lib/r1/libnmf_b200.so
TR=$PWD/pytorch-nmf_b200/lib/trace/libnmf_b200.so
( NMFB200_LIB=$R1 python tools/tc_time.py f16; python tools/tc_time.py f16; NMFB200_FUSED_TAIL=0 python tools/tc_time.py f16; NMFB200_LIB=$R1 python tools/tc_time.py f16; python tools/tc_time.py f16 ) > gpurun_out/ab1.txt 2>&1
grep -E "lib=|rror" gpurun_out/ab1.txt
( NMFB200_LIB=$TR python tools/tc_knock.py f16 0,1,2,4,8,16,24,32,36,60,0 ) > gpurun_out/knock1.txt 2>&1
cat gpurun_out/knock1.txt | tail -14
timeout 1000 python -m pytest tests -m gpu -q > gpurun_out/pytest1.txt 2>&1; tail -40 gpurun_out/pytest1.txt
( python tools/tc_time.py f16 131072 8192 128; NMFB200_LIB=$R1 python tools/tc_time.py f16 131072 8192 128; python tools/tc_time.py f16_split; NMFB200_LIB=$R1 python tools/tc_time.py f16_split ) > gpurun_out/ab2.txt 2>&1
grep -E "lib=|rror" gpurun_out/ab2.txt
