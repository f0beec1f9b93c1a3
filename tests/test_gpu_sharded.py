"""2-GPU NCCL test of the row-sharded fit (skipped unless >= 2 CUDA devices): sharded == single-GPU within rtol."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _inputs():
    torch.manual_seed(0)
    N, C, R = 2048, 768, 64
    V = torch.rand(N, C).bfloat16().float()
    torch.manual_seed(1)
    return V, torch.randn(C, R).abs(), torch.randn(N, R).abs()


def _worker(rank, world, port, precision, out_dir, transport="peer"):
    os.environ["NMFB200_PEER"] = "1" if transport == "peer" else "0"
    for p in (ROOT, os.path.join(ROOT, "pytorch-nmf_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from torchnmf_b200 import NMF
    V, W0, H0 = _inputs()
    bounds = [0, 1152, V.shape[0]]          # uneven shards
    lo, hi = bounds[rank], bounds[rank + 1]
    m = NMF(W=W0, H=H0[lo:hi]).cuda()
    n = m.fit(V[lo:hi].cuda(), 1, 1e-5, 40, precision=precision, group=dist.group.WORLD)
    out = {"W": m.W.data.cpu(), "H": m.H.data.cpu(), "n": n, "path": m.last_w_update_path}
    # a second fit on the cached engine context (the peer connection is reused) must repeat the first bit for bit
    m2 = NMF(W=W0, H=H0[lo:hi]).cuda()
    n2 = m2.fit(V[lo:hi].cuda(), 1, 1e-5, 40, precision=precision, group=dist.group.WORLD)
    out["repeat_ok"] = bool(n2 == n and torch.equal(m2.W.data.cpu(), out["W"]) and torch.equal(m2.H.data.cpu(), out["H"]))
    torch.save(out, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier(); dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("transport", ["peer", "nccl"])
@pytest.mark.parametrize("precision", ["f32", "f16_split", "auto"])
def test_two_gpu_sharded_fit_matches_single(tmp_path, precision, transport):
    """W update over peer memory (the fused P2P sum + ratio stage, include/nmf_b200.h: nmfb200_nmf_update_w_peer) and over the
    NCCL all-reduce (NMFB200_PEER=0): both must match the single-GPU fit; the fp32 kernels have no peer path."""
    from torchnmf_b200 import NMF
    mp.spawn(_worker, args=(2, _free_port(), precision, str(tmp_path), transport), nprocs=2, join=True)
    V, W0, H0 = _inputs()
    ref = NMF(W=W0, H=H0).cuda()
    n_ref = ref.fit(V.cuda(), 1, 1e-5, 40, precision=precision)
    parts = [torch.load(os.path.join(tmp_path, f"r{r}.pt")) for r in range(2)]
    assert parts[0]["n"] == parts[1]["n"] == n_ref
    want = "peer" if (transport == "peer" and precision != "f32") else "nccl"
    assert parts[0]["path"] == parts[1]["path"] == want
    assert parts[0]["repeat_ok"] and parts[1]["repeat_ok"]
    assert torch.equal(parts[0]["W"], parts[1]["W"])                     # replicas stay bit-identical
    H = torch.cat([p["H"] for p in parts])
    rtol = 1e-4 if precision == "f32" else 1e-3      # fp16 operands: the north-star tolerance
    assert torch.allclose(parts[0]["W"], ref.W.data.cpu(), rtol=rtol, atol=1e-6)
    assert torch.allclose(H, ref.H.data.cpu(), rtol=rtol, atol=1e-6)
