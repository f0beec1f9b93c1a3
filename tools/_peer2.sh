cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -5
for pe in 1 0; do
NMFB200_PEER=$pe timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_cfg2_2gpu_peer$pe.json 2> gpurun_out/r2_bench_2gpu_peer$pe.err; tail -1 gpurun_out/r2_bench_cfg2_2gpu_peer$pe.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('peer=$pe', round(d['value'],1), d['config'].get('w_update'), d.get('sharded_check'), (d.get('north_star_cfg4') or {}).get('shard_iters_per_s'))"
done
