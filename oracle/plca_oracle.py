"""CPU oracle for the EM fit of the PLCA family (PLCA, SIPLCA, SIPLCA2, SIPLCA3).

TEST INFRASTRUCTURE ONLY (same rule as oracle/mu_oracle.py: imported by tests/, smoke() and bench.py's CPU legs, never by
the product package).

A closed-form fp32 restatement -- plain torch CPU ops, no autograd, no F.conv* -- of

    torchnmf/plca.py:193-304  BaseComponent.fit   (normalisation of V, one backward pass of V / (WZH + eps) through the
                                                   reconstruction, simultaneous Z / W / H updates, Dirichlet priors,
                                                   loss every 10th iteration, stop rule, returns (n_iter, norm))
    torchnmf/plca.py:371-373  PLCA.reconstruct    H @ (W * Z)^T
    torchnmf/plca.py:453-455, :534-537, :621-625  SIPLCA*.reconstruct: conv with the flipped kernel W * Z, full padding

The gradients the reference takes by autograd are, with P = V / (WZH + eps) and Hz = H * Z (rank axis):
    W.grad = "wgrad"(P, Hz)        H.grad = Z * "dgrad"(P, W)        Z.grad[r] = sum over everything else of H * dgrad(P, W)
where wgrad / dgrad are the contractions of oracle/mu_oracle.py (matrix products for PLCA, shifted sums for SIPLCA*).

Parity status: PINNED by tests/golden/reference_next.npz (PLCA) and tests/golden/reference_plca.npz (SIPLCA*), both written
by oracle/make_golden.py from the real torchnmf 0.3.5; tests/test_plca.py checks this file against them.
"""
import math

import torch

from . import mu_oracle as mu

EPS = mu.EPS


def get_norm(x):
    # plca.py:24-31
    if x.ndim > 1:
        return x.sum([d for d in range(x.dim()) if d != 1], keepdim=True)
    return x.sum()


def _over(z, x):
    return z[(slice(None),) + (None,) * (x.dim() - 2)]


def reconstruct(H, W, Z):
    if W.dim() == 2:
        return (H * Z) @ W.t()                                            # plca.py:371-373
    return mu.nmfnd_reconstruct(H * _over(Z, H), W)                       # plca.py:453-455 etc., as shifted products


def gradients(P, H, W, Z):
    """(W.grad, H.grad, Z.grad) of <WZH, P> (plca.py:251-253)."""
    Hz = H * _over(Z, H)
    if W.dim() == 2:
        dW, dHz = P.t() @ Hz, P @ W
    else:
        dW = mu.nmfnd_grad_w(P, Hz, tuple(W.shape[2:]))
        dHz = mu.nmfnd_grad_h(P, W, tuple(H.shape[2:]))
    return dW, dHz * _over(Z, H), get_norm(H * dHz).reshape(-1)


def _loss(WZH, Vn, norm):
    # plca.py:245-246: sqrt(2 kl_div(WZH * norm, V * norm)); metrics.py:22
    x, t = WZH * norm, Vn * norm
    d = float(t.reshape(-1) @ (torch.log(t + EPS) - torch.log(x + EPS)).reshape(-1) - t.sum() + x.sum())
    return math.sqrt(2.0 * d) if d >= 0 else float("nan")


def _prior(x, alpha, renorm):
    # plca.py:259-261 / :272-275 / :285-288
    if isinstance(alpha, torch.Tensor) or alpha != 1:
        x = x + (alpha - 1)
        x = torch.where(x > EPS, x, torch.full_like(x, EPS))             # F.threshold(x, eps, eps)
        if renorm:
            x = x / get_norm(x)
    return x


def fit(V, W, H, Z, tol=1e-4, max_iter=200, W_alpha=1., H_alpha=1., Z_alpha=1.,
        trainable_W=True, trainable_H=True, trainable_Z=True):
    """Returns (W, H, Z, n_iter, norm).  W, H, Z are the (already normalised) module parameters, plca.py:110-144."""
    W, H, Z = W.clone(), H.clone(), Z.clone()
    norm = V.sum()
    Vn = V.contiguous() / norm                                            # plca.py:241-242
    loss_init = previous = _loss(reconstruct(H, W, Z), Vn, norm)
    n_iter = -1
    for n_iter in range(max_iter):
        P = Vn / (reconstruct(H, W, Z) + EPS)                             # plca.py:252-253
        dW, dH, dZ = gradients(P, H, W, Z)
        Z_prior = None
        if trainable_Z:                                                   # plca.py:256-262
            Z = Z * dZ.clamp_min(0)
            Z_prior = Z.clone()
            Z = _prior(Z, Z_alpha, False)
            Z = Z / Z.sum()
        if trainable_W:                                                   # plca.py:264-275
            W = W * dW.clamp_min(0)
            if Z_prior is None:
                div = get_norm(W)
                Z_prior = div.squeeze()
            else:
                div = _over(Z_prior, W)
            W = _prior(W / div, W_alpha, True)
        if trainable_H:                                                   # plca.py:277-288
            H = H * dH.clamp_min(0)
            div = get_norm(H) if Z_prior is None else _over(Z_prior, H)
            H = _prior(H / div, H_alpha, True)
        if n_iter % 10 == 9:                                              # plca.py:290-302
            loss = _loss(reconstruct(H, W, Z), Vn, norm)
            if (previous - loss) / loss_init < tol:
                break
            previous = loss
    return W, H, Z, n_iter, norm
