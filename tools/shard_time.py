"""Per-iteration time of the row-sharded update under torchrun: local (no exchange) vs peer-memory vs NCCL W update."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-nmf_b200")]
import torch, torch.distributed as dist
from torchnmf_b200.engine import CudaNmfEngine, ShardedEngine, release_workspaces
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
N, C, R = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (65536, 4096, 64)
torch.manual_seed(rank)
V = torch.rand(N, C, device="cuda").bfloat16().float()
torch.manual_seed(100)
W = torch.randn(C, R, device="cuda").abs(); H = torch.randn(N, R, device="cuda").abs()
loc = CudaNmfEngine(V, W, H, "f16")
def t(fn, n=300):
    for _ in range(10): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    x = torch.tensor([e0.elapsed_time(e1) / n * 1e3], device="cuda")
    dist.all_reduce(x, op=dist.ReduceOp.MAX)
    return float(x)
def local(): loc.update_w(1.0, 1.0, 0.0, 0.0); loc.update_h(1.0, 1.0, 0.0, 0.0)
res = {"local": t(local)}
for name, env in (("peer", "1"), ("nccl", "0")):
    os.environ["NMFB200_PEER"] = env
    sh = ShardedEngine(loc, dist.group.WORLD)
    def both(): sh.update_w(1.0, 1.0, 0.0, 0.0); sh.update_h(1.0, 1.0, 0.0, 0.0)
    res[name] = t(both)
    res[name + "_w_only"] = t(lambda: sh.update_w(1.0, 1.0, 0.0, 0.0))
    assert sh.w_update_path == name, sh.w_update_path
res["local_w_only"] = t(lambda: loc.update_w(1.0, 1.0, 0.0, 0.0))
loc.check_health()
if rank == 0:
    print(f"world {world} {N}x{C} R={R}: " + "  ".join(f"{k} {v:.1f}" for k, v in res.items()) + "  (us per call, sustained, max over ranks)", flush=True)
loc.close(); release_workspaces(); dist.destroy_process_group()
