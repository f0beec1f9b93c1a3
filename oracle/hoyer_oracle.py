"""CPU oracle for the sparseness-constrained path of the reference (Hoyer 2004): `_proj_func`, `sparse_fit`, `SparsityProj`.

TEST INFRASTRUCTURE ONLY (same rules as oracle/mu_oracle.py: imported by tests/ only, never by the product).

Closed-form fp32 restatement (plain torch CPU ops, no autograd, no TorchScript) of

    torchnmf/nmf.py:21-49     _proj_func        (projection onto {v >= 0, |v|_1 = k1, |v|_2^2 = k2})
    torchnmf/nmf.py:134-159   _get_norm, _renorm
    torchnmf/nmf.py:411-599   BaseComponent.sparse_fit  (dense targets)
    torchnmf/trainer.py:124-190 SparsityProj.step       (the projected-gradient step around a closure)

Parity status: PINNED by tests/golden/reference_hoyer.npz (`python oracle/make_golden.py --hoyer`, the real torchnmf 0.3.5)
through tests/test_hoyer.py.
"""
import torch

from . import mu_oracle as orc

EPS = orc.EPS


def proj_func(s, k1, k2):
    """nmf.py:21-49.  Returns a new tensor of s's shape."""
    shape = s.shape
    s = s.reshape(-1)
    N = s.numel()
    v = s + (k1 - s.sum()) / N                                   # :28
    zero = torch.zeros(N, dtype=torch.bool)
    while True:
        m = k1 / (N - int(zero.sum()))                           # :32
        w = torch.where(~zero, v - m, v)                         # :33
        a = w @ w
        b = 2 * (w @ v)
        c = v @ v - k2
        alphap = (-b + (b * b - 4 * a * c).clamp_min(0).sqrt()) * 0.5 / a      # :37
        v = v + float(alphap) * w                                # :38
        neg = v < 0
        if not bool(neg.any()):                                  # :41-42
            break
        zero |= neg                                              # :44
        v = v.clamp_min(0)
        v = v + (k1 - v.sum()) / (N - int(zero.sum()))           # :46
        v = v.clamp_min(0)                                       # :47
    return v.view(shape)


def get_norm(x, axis=1):
    """nmf.py:134-139: L2 norm over every axis but `axis`."""
    dims = [d for d in range(x.dim()) if d != axis]
    return (x * x).sum(dims).sqrt()


def project_slices(x, dim, k1, k2):
    """x[..., j, ...] <- proj_func(x[..., j, ...], k1[j], k2[j]) for every j along `dim` (the Python loops of
    nmf.py:464-465, :519-522, trainer.py:176-181); what nmfb200_hoyer_project does in one launch."""
    out = x.clone()
    for j in range(x.shape[dim]):
        sl = (slice(None),) * dim + (j,)
        out[sl] = proj_func(x[sl], float(k1[j]), float(k2[j]))
    return out


def renorm_h(W, H):
    """nmf.py:142-159 with unit_norm='H' (the only form sparse_fit uses, :581)."""
    n = get_norm(H)
    hs = (slice(None),) + (None,) * (H.dim() - 2)
    ws = (slice(None),) + (None,) * (W.dim() - 2)
    return W * n[ws], H / n[hs]


def dloss_dwh(V, WH, beta):
    """d beta_div(WH, V) / d WH, with the eps placement of metrics.py:22, :39, :56-57, :91-96."""
    if beta == 2:
        return WH - V
    if beta == 1:
        return 1.0 - V / (WH + EPS)
    x = WH + EPS
    if beta == 0:
        return 1.0 / x - (V + EPS) / (x * x)
    t = V + EPS if beta < 0 else V
    return x.pow(beta - 1) - t * x.pow(beta - 2)


_KINDS = {
    "nmf": (orc.nmf_reconstruct, lambda G, H, W: G.t() @ H, lambda G, W, H: G @ W,
            orc.nmf_update_w, orc.nmf_update_h),
    "nmfd": (orc.nmfd_reconstruct, lambda G, H, W: orc.nmfd_grad_w(G, H, W.shape[2]),
             lambda G, W, H: orc.nmfd_grad_h(G, W, H.shape[2]), orc.nmfd_update_w, orc.nmfd_update_h),
    "nmfnd": (orc.nmfnd_reconstruct, lambda G, H, W: orc.nmfnd_grad_w(G, H, tuple(W.shape[2:])),
              lambda G, W, H: orc.nmfnd_grad_h(G, W, tuple(H.shape[2:])), orc.nmfnd_update_w, orc.nmfnd_update_h),
}


def sparse_fit(V, W, H, beta=2, max_iter=200, sW=None, sH=None, trainable_W=True, trainable_H=True, kind="nmf"):
    """nmf.py:411-599 for a dense V.  Returns (W, H, n_iter, trace) with trace = [(stepsize_W, stepsize_H)] per iteration."""
    recon, grad_w, grad_h, upd_w, upd_h = _KINDS[kind]
    assert bool(torch.all(V >= 0)), "Target should be non-negative."
    if float(V.min()) == 0 and beta <= 0:
        raise ValueError("When beta <= 0 and V contains zeros, the training process may diverge. "
                         "Please add small values to V, or use a positive beta value.")
    W, H = W.clone(), H.clone()
    R = W.shape[1]
    L1a = L1s = None
    if sW is not None and trainable_W:                                          # :459-467
        dim = W[:, 0].numel()
        L1a = dim ** 0.5 * (1 - sW) + sW
        W = project_slices(W, 1, [L1a] * R, [1.0] * R)
    if sH is not None and trainable_H:                                          # :469-477
        dim = H[:, 0].numel()
        L1s = dim ** 0.5 * (1 - sH) + sH
        H = project_slices(H, 1, [L1s] * R, [1.0] * R)
    gamma = orc.gamma_of(beta)                                                  # :479-484
    step_w = step_h = 1.0
    trace = []
    n_iter = -1
    for n_iter in range(max_iter):
        if trainable_W:
            if L1a is None:
                W = upd_w(V, W, H, beta, gamma, 0.0, 0.0)                       # :503-511
            else:
                WH = recon(H, W)
                loss = orc.beta_div(WH, V, beta)
                g = grad_w(dloss_dwh(V, WH, beta), H, W)
                for _ in range(10):                                             # :519-535
                    Wn = W - step_w * g
                    norms = get_norm(Wn)
                    Wn = project_slices(Wn, 1, L1a * norms, norms ** 2)
                    if orc.beta_div(recon(H, Wn), V, beta) <= loss:
                        break
                    step_w *= 0.5
                step_w *= 1.2                                                   # :537
                W = Wn
        if trainable_H:
            if L1s is None:
                H = upd_h(V, W, H, beta, gamma, 0.0, 0.0)                       # :549-557
            else:
                WH = recon(H, W)
                loss = orc.beta_div(WH, V, beta)
                g = grad_h(dloss_dwh(V, WH, beta), W, H)
                for _ in range(10):                                             # :567-585
                    Hn = H - step_h * g
                    norms = get_norm(Hn)
                    Hn = project_slices(Hn, 1, L1s * norms, norms ** 2)
                    if orc.beta_div(recon(Hn, W), V, beta) <= loss:
                        break
                    step_h *= 0.5
                step_h *= 1.2
                H = Hn
            W, H = renorm_h(W, H)                                               # :588
        trace.append((step_w, step_h))
    return W, H, n_iter + 1, trace


def sparsity_proj_step(params, grads, loss_fn, sparsity, lr, dim=1, max_iter=10):
    """trainer.py:150-190 for one parameter group: `params` are updated in place (list of tensors), `grads` their gradients
    at entry, `loss_fn()` re-evaluates the loss on the current params.  Returns (loss, new lr)."""
    init_loss = loss_fn()
    loss = None
    for _ in range(max_iter):
        for p, g in zip(params, grads):
            norms = get_norm(p, dim)                                            # :173 (norms of p BEFORE the step)
            p.add_(g, alpha=-lr)
            N = p.numel() // p.shape[dim]
            L1 = N ** 0.5 * (1 - sparsity) + sparsity
            p.copy_(project_slices(p, dim, L1 * norms, norms ** 2))
        loss = loss_fn()
        if loss <= init_loss:
            break
        for p, g in zip(params, grads):
            p.add_(g, alpha=lr)                                                 # :186-187
        lr *= 0.5
    lr *= 1.2
    return loss, lr
