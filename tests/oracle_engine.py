"""Test-only engines that implement the engine interface of torchnmf_b200.engine with the CPU oracle.

They exist so the HOST logic (fit loop, validation, stop rule, sharded all-reduce protocol) can be
exercised on a machine with no GPU (`-m "not gpu"`, gloo).  They are never imported by the product.
"""
import torch

from oracle import mu_oracle as orc


class _OracleEngine:
    device = torch.device("cpu")
    precision = "oracle"

    def __init__(self, V, W, H):
        self.V, self.W, self.H = V, W, H      # W, H: the Parameter storages, updated in place

    def close(self):
        pass

    def sync(self):
        pass

    def minmax(self):
        return float(self.V.min()), float(self.V.max())

    def loss_tensor(self, beta):
        return orc.beta_div(self._recon(self.H, self.W), self.V, beta).double().reshape(1)

    def loss(self, beta):
        return float(self.loss_tensor(beta))


class OracleNmfEngine(_OracleEngine):
    kind = "nmf"
    _recon = staticmethod(orc.nmf_reconstruct)

    def update_w(self, beta, gamma, l1, l2):
        self.W.copy_(orc.nmf_update_w(self.V, self.W, self.H, beta, gamma, l1, l2))

    def update_h(self, beta, gamma, l1, l2):
        self.H.copy_(orc.nmf_update_h(self.V, self.W, self.H, beta, gamma, l1, l2))

    # sharded pieces: same buffer layout as nmfb200_nmf_w_partial (include/nmf_b200.h)
    def w_partial(self, beta):
        num, den = orc.nmf_w_contractions(self.V, self.W, self.H, beta)
        return torch.cat([num.reshape(-1), den.reshape(-1)]).contiguous()

    def w_apply(self, reduced, beta, gamma, l1, l2):
        C, R = self.W.shape
        num = reduced[:C * R].view(C, R)
        den = reduced[C * R:].view(1, R) if beta == 1 else reduced[C * R:].view(C, R)
        self.W.copy_(orc._ratio_update(self.W, num, den, gamma, l1, l2, beta == 1))


class OracleNmfdEngine(_OracleEngine):
    kind = "nmfd"
    _recon = staticmethod(orc.nmfd_reconstruct)

    def update_w(self, beta, gamma, l1, l2):
        self.W.copy_(orc.nmfd_update_w(self.V, self.W, self.H, beta, gamma, l1, l2))

    def update_h(self, beta, gamma, l1, l2):
        self.H.copy_(orc.nmfd_update_h(self.V, self.W, self.H, beta, gamma, l1, l2))
