cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
./tools/ubench/pipes > gpurun_out/r2_ubench_pipes.txt 2>&1 || true
./tools/ubench/ratio > gpurun_out/r2_ubench_ratio.txt 2>&1 || true
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke.txt 2>&1; tail -3 gpurun_out/r2_smoke.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.txt 2>&1; tail -5 gpurun_out/r2_pytest_gpu.txt
