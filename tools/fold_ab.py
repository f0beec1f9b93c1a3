"""A/B on one box: fit() at BASELINE configs[1] with every 10th iteration's loss folded into the next W update's contraction
(nmfb200_nmf_loss_prefetch_w) versus a loss pass of its own.  Arms alternate; CUDA-event time per 200-iteration fit."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-nmf_b200")]
import torch
from torchnmf_b200 import NMF, engine, _capi

N, C, R, ITERS = 65536, 4096, 64, 200
torch.manual_seed(0)
V = torch.rand(N, C, device="cuda").bfloat16().float()
torch.manual_seed(1)
W0, H0 = torch.randn(C, R).abs().cuda(), torch.randn(N, R).abs().cuda()
m = NMF(W=W0, H=H0).cuda()
res = {True: [], False: []}
launches = {}
for rep in range(5):
    for fold in (True, False):
        engine.LOSS_FOLD = fold
        m.W.data.copy_(W0); m.H.data.copy_(H0)
        n0 = _capi.launch_count()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        m.fit(V, 1, float("-inf"), ITERS)
        b.record(); torch.cuda.synchronize()
        launches[fold] = _capi.launch_count() - n0
        if rep >= 2:                                   # two warm-up rounds (clocks settle at the power cap)
            res[fold].append(a.elapsed_time(b))
    if rep == 0:
        Wf = m.W.data.clone()
out = {"workload": "NMF 65536x4096 R=64 KL, 200 iterations per fit, precision " + m.last_fit_precision}
for fold in (True, False):
    ms = sorted(res[fold])[len(res[fold]) // 2]
    out["folded" if fold else "own_pass"] = {"ms_per_fit": [round(x, 2) for x in res[fold]], "median_it_s": round(ITERS / ms * 1e3, 1),
                                             "launches_per_fit": launches[fold]}
out["speedup"] = round(out["folded"]["median_it_s"] / out["own_pass"]["median_it_s"], 4)
print(json.dumps(out))
