"""The reference's own CPU tests that touch no fit kernel, run unchanged in spirit against this package:
tests/test_metrics.py:6-19 (value ranges of beta_div / sparseness) and tests/test_trainer.py:10-53 (BetaMu on a composed
nn.Sequential of three NMF layers, SparsityProj on one factor -- non-negativity after every step).  BetaMu takes its generic
autograd path here (a composed graph is not a single leaf); SparsityProj's projection is the library's, so on this GPU-less
side the oracle's restatement stands in for it (tests/test_hoyer.py pins both to the reference's outputs)."""
import pytest
import torch
from torch import nn

from oracle import hoyer_oracle as hoy
from torchnmf_b200 import NMF, BetaMu, SparsityProj
from torchnmf_b200 import trainer as _trainer
from torchnmf_b200.metrics import beta_div, sparseness


@pytest.mark.parametrize("beta", [-1, 0, 0.5, 1, 1.5, 2, 3])
@pytest.mark.parametrize("x, y", [(torch.zeros(100), torch.rand(100)), (torch.rand(100), torch.rand(100)),
                                  (torch.rand(100), torch.zeros(100)), (torch.zeros(100), torch.zeros(100))])
def test_beta_value_range(beta, x, y):
    loss = beta_div(x, y, beta)
    assert not torch.any(torch.isnan(loss)), loss.item()
    assert not torch.any(loss < 0), loss.item()


@pytest.mark.parametrize("x", [torch.rand(100)])
def test_sparseness_value_range(x):
    loss = sparseness(x)
    assert not torch.any(torch.isnan(loss)), loss.item()
    assert 0.0 <= float(loss) <= 1.0


@pytest.mark.parametrize("beta", [-1, 0, 0.5, 1, 1.5, 2, 3])
@pytest.mark.parametrize("l1_reg, l2_reg, orthogonal", [(0, 0, 0), (1e-3, 0, 1e-2), (0, 1e-3, 0), (1e-3, 1e-3, 1e-2)])
def test_beta_trainer(beta, l1_reg, l2_reg, orthogonal):
    torch.manual_seed(0)
    m = nn.Sequential(NMF((100, 16), rank=8), NMF(W=(32, 16)), NMF(W=(50, 32)))
    target = torch.rand(100, 50)
    trainer = BetaMu(m.parameters(), beta, l1_reg, l2_reg, orthogonal)

    def closure():
        trainer.zero_grad()
        return target, m(None)

    for _ in range(10):
        trainer.step(closure)
        assert set(trainer.last_step_paths) == {"autograd"}
        for p in m.parameters():
            assert torch.all(p >= 0.)


@pytest.mark.parametrize("attr", ["W", "H"])
def test_sparse_trainer(attr, monkeypatch):
    monkeypatch.setattr(_trainer, "_project_slices_", lambda p, dim, k1, k2: p.copy_(hoy.project_slices(p, dim, k1, k2)))
    torch.manual_seed(0)
    m = NMF((100, 50))
    target = torch.rand(100, 50)
    trainer = SparsityProj([getattr(m, attr)], 0.2)

    def closure():
        trainer.zero_grad()
        return beta_div(m(None), target)

    for _ in range(10):
        trainer.step(closure)
        assert torch.all(getattr(m, attr) >= 0.)


# tests/test_nmf.py:120-136 (`test_sparse_fit`): every beta branch, with and without a constraint, runs its max_iter iterations and
# stays finite.  Host logic of `sparse_fit` with the oracle in place of the library (the GPU path: tests/test_hoyer.py).
@pytest.mark.parametrize("beta", [-1, 0, 0.5, 1, 1.5, 2, 2.5])
@pytest.mark.parametrize("sW, sH", [(None,) * 2, (0.3, None), (None, 0.3)])
def test_sparse_fit(beta, sW, sH):
    from oracle_engine import OracleNmfEngine
    torch.manual_seed(0)
    max_iter = 20
    V = torch.rand(100, 50) + 1e-3
    m = NMF(V.shape, 8)
    m._engine_factory = OracleNmfEngine
    n_iter = m.sparse_fit(V, beta, max_iter, beta == 2, sW, sH)
    assert n_iter == max_iter
    assert not torch.any(torch.isnan(m.W))
    assert not torch.any(torch.isnan(m.H))
    assert torch.all(m.W >= 0) and torch.all(m.H >= 0)


def test_public_surface_covers_the_references():
    """Every public name of torchnmf 0.3.5 (its `__all__` lists, the optimizers of trainer.py, the functions of metrics.py and
    utils.py, the module-level helpers trainer.py imports from nmf.py) exists here under the same module path."""
    import inspect
    import torchnmf_b200 as pkg
    want = {
        "nmf": ["BaseComponent", "NMF", "NMFD", "NMF2D", "NMF3D", "_proj_func", "_get_norm", "_renorm"],       # nmf.py:16-18, :21, :134, :142
        "plca": ["PLCA", "SIPLCA", "SIPLCA2", "SIPLCA3", "BaseComponent"],                                      # plca.py:13-15
        "trainer": ["BetaMu", "SparsityProj"],                                                                   # trainer.py:7, :124
        "metrics": ["kl_div", "euclidean", "is_div", "beta_div", "sparseness"],                                 # metrics.py:6-115
        "utils": ["normalize", "renorm_"],                                                                       # utils.py:5-13
        "constants": ["eps"],                                                                                    # constants.py:3
    }
    for mod, names in want.items():
        m = getattr(pkg, mod)
        for n in names:
            assert hasattr(m, n), f"{mod}.{n}"
    for cls, methods in ((pkg.nmf.BaseComponent, ["fit", "sparse_fit", "forward", "reconstruct", "extra_repr"]),
                         (pkg.plca.BaseComponent, ["fit", "forward", "reconstruct"])):
        for meth in methods:
            assert callable(getattr(cls, meth)), f"{cls.__name__}.{meth}"
    # positional signatures of the fit entry points are the reference's (nmf.py:298-306, :412-419; plca.py:193-201)
    assert list(inspect.signature(pkg.NMF.fit).parameters)[:8] == ["self", "V", "beta", "tol", "max_iter", "verbose", "alpha", "l1_ratio"]
    assert list(inspect.signature(pkg.NMF.sparse_fit).parameters)[:7] == ["self", "V", "beta", "max_iter", "verbose", "sW", "sH"]
    assert list(inspect.signature(pkg.PLCA.fit).parameters)[:8] == ["self", "V", "tol", "max_iter", "verbose", "W_alpha", "H_alpha", "Z_alpha"]
    assert float(pkg.constants.eps) == float(torch.finfo(torch.float32).eps)
