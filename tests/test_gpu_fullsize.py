"""Full-size checks at the BASELINE.json config shapes (cfg 2, 3, 4).  The CPU oracle is too slow at these sizes, so the
checks are size-independent properties and comparisons with plain fp32 torch on the same GPU."""
import numpy as np
import pytest
import torch

from recommendation_b200 import functional as F_, losses, synth
from recommendation_b200.graph import CSRGraph

pytestmark = pytest.mark.gpu


def test_cfg2_ssl_layer_and_infonce_at_full_size(cuda):
    """NCL at the Amazon-book shape: B = 4096 against all U = 52,643 users / I = 91,599 items, d = 64."""
    U, I, E, d, K = synth.CONFIGS["cfg2"]
    g = torch.Generator(device=cuda).manual_seed(2)
    ctx = torch.randn(U + I, d, device=cuda, generator=g)
    ini = (ctx * 0.5 + torch.randn(U + I, d, device=cuda, generator=g)).requires_grad_(True)
    bu = torch.randint(0, U, (4096,), device=cuda, generator=g)
    bi = torch.randint(0, I, (4096,), device=cuda, generator=g)
    ncl = losses.NCLLosses(U, I, 0.1, 1e-6, 1.5, 8e-8, 4096)
    loss = ncl.ssl_layer_loss(ctx, ini, bu, bi)
    loss.backward()

    def side(c, z, idx):  # ncl.py:358-367 in eager fp32
        cn, zn = torch.nn.functional.normalize(c[idx]), torch.nn.functional.normalize(z)
        return (torch.logsumexp(cn @ zn.T / 0.1, 1) - (cn * zn[idx]).sum(1) / 0.1).sum()
    ini2 = ini.detach().clone().requires_grad_(True)
    want = 1e-6 * (side(ctx[:U], ini2[:U], bu) + 1.5 * side(ctx[U:], ini2[U:], bi))
    want.backward()
    np.testing.assert_allclose(loss.item(), want.item(), rtol=2e-2)
    err = (ini.grad - ini2.grad).norm() / ini2.grad.norm()
    assert err < 1e-2, f"relative gradient error {err:.3e}"
    # in-batch InfoNCE, B = 4096
    v1 = torch.randn(4096, d, device=cuda, generator=g); v2 = v1 * 0.7 + torch.randn(4096, d, device=cuda, generator=g)
    got = losses.InfoNCE(v1, v2, 0.2)
    s = torch.nn.functional.normalize(v1) @ torch.nn.functional.normalize(v2).T / 0.2
    np.testing.assert_allclose(got.item(), (-torch.diag(torch.log_softmax(s, 1)).mean()).item(), rtol=2e-2)


def test_cfg3_directau_selfcf_shapes_at_full_size(cuda):
    """DirectAU on the Yelp2018 shape: 2-layer d = 128 propagation (3.1 M non-zeros) + alignment / uniformity at B = 2048."""
    inter, d, K = synth.config_graph("cfg3")
    U, I = inter.n_users, inter.n_items
    csr = CSRGraph.from_pairs(torch.from_numpy(inter.users).to(cuda), torch.from_numpy(inter.items).to(cuda), U, I, norm="sym")
    x = (torch.randn(U + I, d, device=cuda) * 0.1).requires_grad_(True)
    final = F_.propagate(csr, x, K, mode="mean")
    g = torch.Generator(device=cuda).manual_seed(3)
    bu = torch.randint(0, U, (2048,), device=cuda, generator=g); bp = torch.randint(0, I, (2048,), device=cuda, generator=g)
    u_emb, p_emb = F_.gather_rows(final[:U], bu), F_.gather_rows(final[U:], bp)
    dau = losses.DirectAULosses(0.7)
    loss = dau.calculate_loss(u_emb, p_emb)
    loss.backward()
    # eager fp32 reference on the same propagated rows (directau.py:240-251)
    ud, pd_ = u_emb.detach().double(), p_emb.detach().double()
    un, pn = torch.nn.functional.normalize(ud, dim=-1), torch.nn.functional.normalize(pd_, dim=-1)
    unif = lambda z: (torch.pdist(z, p=2).pow(2).mul(-2).exp().mean() + 1e-8).log()
    want = (un - pn).pow(2).sum(1).mean() + 0.7 * (unif(un) + unif(pn)) / 2
    np.testing.assert_allclose(loss.item(), want.item(), rtol=2e-2, atol=2e-3)
    assert torch.isfinite(x.grad).all() and x.grad.abs().sum() > 0
    # linearity of the 2-layer mean propagation at full size
    y = torch.randn_like(x)
    a, b, ab = (F_.propagate(csr, t, K, mode="mean") for t in (x.detach(), y, x.detach() + 3 * y))
    torch.testing.assert_close(ab, a + 3 * b, rtol=1e-4, atol=1e-5)


def test_cfg4_social_operators_at_full_size(cuda):
    """MHCN / DiffNet shape: 250k users x 125k items, 5 M interactions + 1 M directed social edges; non-symmetric
    operators, so the backward runs on the transposed CSR: check <y, A x> = <A^T y, x> and row-normalisation."""
    U, I, E, d, K = synth.CONFIGS["cfg4"]
    inter = synth.power_law_bipartite(U, I, E, seed=1004)
    rng = np.random.default_rng(4)
    s_src = torch.from_numpy(rng.integers(0, U, 1_000_000)).to(cuda)
    s_dst = torch.from_numpy(rng.integers(0, U, 1_000_000)).to(cuda)
    S = CSRGraph.from_coo(s_src, s_dst, None, U, U, norm="row")                       # diffnet.py:1070-1078 weights
    R = CSRGraph.from_coo(torch.from_numpy(inter.users).to(cuda), torch.from_numpy(inter.items).to(cuda), None, U, I, norm="row")
    rs = torch.zeros(U, device=cuda).index_add_(0, torch.repeat_interleave(torch.arange(U, device=cuda), (S.row_ptr[1:] - S.row_ptr[:-1]).long()), S.vals)
    has = (S.row_ptr[1:] - S.row_ptr[:-1]) > 0
    torch.testing.assert_close(rs[has], torch.ones_like(rs[has]), rtol=1e-5, atol=1e-5)   # rows of D^-1 A sum to one
    for op, n_in in ((S, U), (R, I)):
        x = torch.randn(n_in, d, device=cuda, requires_grad=True)
        y = torch.randn(op.n_rows, d, device=cuda)
        ax = F_.spmm(op, x)
        (ax * y).sum().backward()                                                      # = <A^T y, x> through the transposed CSR
        aty = torch.empty(n_in, d, device=cuda)
        F_.spmm_raw(op.transpose(), y, y=aty)
        torch.testing.assert_close(x.grad, aty, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close((ax.detach() * y).sum().double(), (aty * x.detach()).sum().double(), rtol=1e-4, atol=1e-2)
