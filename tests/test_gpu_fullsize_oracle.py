"""Oracle parity at BASELINE.json's own sizes (cfg 1, 2, 3): the CUDA path against oracle/ on identical inputs, at the
north-star tolerances (rtol 1e-3 for fp32 paths, 2e-2 for the bf16 logits of the InfoNCE family).  The CPU side takes a few
seconds per config on the GPU box's host cores (a cfg1 step is ~2 s).

Element-wise rtol needs a floor for entries that are sums with cancellation: atol = 1e-3 x the RMS of the expected tensor for
fp32 paths (stated at each assert), i.e. an entry may be off by 1e-3 of itself or by 1e-3 of the typical magnitude.
"""
import numpy as np
import pytest
import torch

from oracle import lightgcn_ref, losses_ref
from recommendation_b200 import functional as F_, losses, synth
from recommendation_b200.graph import CSRGraph
from recommendation_b200.lightgcn import LightGCN, bpr_step_loss, build_edge_index

pytestmark = pytest.mark.gpu
RTOL = 1e-3


def _close(got: torch.Tensor, want: torch.Tensor, what: str, rtol: float = RTOL):
    want = want.detach().cpu().double().numpy()
    got = got.detach().cpu().double().numpy()
    rms = float(np.sqrt((want ** 2).mean()))
    np.testing.assert_allclose(got, want, rtol=rtol, atol=rtol * rms, err_msg=what)


def _torch_sparse_adj(inter):
    """selfcf.py:297-306 + 240-255 + 219-225 on the CPU: scipy bipartite adjacency, D^-1/2 A D^-1/2, torch COO tensor."""
    import scipy.sparse as sp

    U, I = inter.n_users, inter.n_items
    n = U + I
    tmp = sp.csr_matrix((np.ones(inter.n_edges, np.float32), (inter.users, inter.items + U)), shape=(n, n))
    adj = tmp + tmp.T
    rowsum = np.asarray(adj.sum(1)).ravel()
    with np.errstate(divide="ignore"):
        dinv = np.power(rowsum, -0.5)
    dinv[np.isinf(dinv)] = 0.0
    norm = sp.diags(dinv) @ adj @ sp.diags(dinv)
    coo = norm.tocoo()
    idx = torch.stack([torch.from_numpy(coo.row).long(), torch.from_numpy(coo.col).long()])
    return torch.sparse_coo_tensor(idx, torch.from_numpy(coo.data).float(), coo.shape).coalesce()


def test_cfg1_lightgcn_step_against_oracle_at_full_size(cuda):
    """lightgcn.py:21-27 + 85-118 on the Gowalla-shaped graph (29,858 x 40,981, 1.03 M edges, 3 layers, d = 64):
    propagated embeddings, loss and both table gradients of one full-batch step, GPU path vs oracle/lightgcn_ref (fp32 CPU)."""
    inter, d, K = synth.config_graph("cfg1")
    U, I, E = inter.n_users, inter.n_items, inter.n_edges
    pu, pi = torch.from_numpy(inter.users), torch.from_numpy(inter.items)
    torch.manual_seed(11)
    model = LightGCN(U, I, d, K).to(cuda)
    uw = model.user_embedding.weight.detach().cpu().clone().requires_grad_(True)
    iw = model.item_embedding.weight.detach().cpu().clone().requires_grad_(True)
    ei = build_edge_index(pu, pi, U)
    neg = torch.from_numpy(np.random.default_rng(11).integers(0, I, E))
    ue_w, ie_w = lightgcn_ref.lightgcn_forward(uw, iw, ei, K)
    # the loss of lightgcn.py:95-118 evaluated in float64 on the fp32 embeddings: in fp32 on the CPU, norm(2).pow(2) over the
    # 1 M gathered rows stagnates (addends of ~1e-5 against a running sum of ~1e3) and comes out 0.5 % low
    want = losses_ref.bpr_lightgcn(ue_w.double(), ie_w.double(), pu, pi, neg, 1e-4)
    want.backward()
    ue, ie = model(ei.to(cuda))
    _close(ue, ue_w, "propagated user embeddings")
    _close(ie, ie_w, "propagated item embeddings")
    got = bpr_step_loss(ue, ie, pu.to(cuda), pi.to(cuda), neg.to(cuda), 1e-4)
    np.testing.assert_allclose(got.item(), want.item(), rtol=1e-4)
    got.backward()
    _close(model.user_embedding.weight.grad, uw.grad, "dL/d user table")
    _close(model.item_embedding.weight.grad, iw.grad, "dL/d item table")


def test_cfg2_ncl_encoder_and_ssl_layer_loss_against_oracle_at_full_size(cuda):
    """ncl.py:415-422 (3-layer mean propagation via torch.sparse.mm) and ncl.py:358-367 (B = 4096 against ALL 52,643 users /
    91,599 items) on the Amazon-book-shaped graph (2.98 M edges): GPU path vs the CPU torch.sparse path / oracle/losses_ref."""
    inter, d, K = synth.config_graph("cfg2")
    U, I = inter.n_users, inter.n_items
    adj = _torch_sparse_adj(inter)
    torch.manual_seed(12)
    x0 = (torch.randn(U + I, d) * 0.1).requires_grad_(True)
    layers = [x0]
    for _ in range(K):
        layers.append(torch.sparse.mm(adj, layers[-1]))                      # ncl.py:419
    mean_w = torch.stack(layers, dim=1).mean(dim=1)                           # ncl.py:420-421
    g = torch.Generator().manual_seed(12)
    bu, bi = torch.randint(0, U, (4096,), generator=g), torch.randint(0, I, (4096,), generator=g)
    ssl_w = losses_ref.ssl_layer_loss(layers[2], layers[0], bu, bi, U, 0.1, 1e-6, 1.5)   # hyper_layers = 1: context = E(2)
    (mean_w.pow(2).sum() * 1e-3 + ssl_w).backward()

    csr = CSRGraph.from_pairs(torch.from_numpy(inter.users).to(cuda), torch.from_numpy(inter.items).to(cuda), U, I, norm="sym")
    x0g = x0.detach().to(cuda).requires_grad_(True)
    mean_g, tail = F_.propagate(csr, x0g, K, mode="mean", return_layers=True)
    layers_g = [x0g] + tail                                                   # return_layers gives E(1)..E(K)
    _close(mean_g, mean_w, "mean of the propagated layers")
    for k in range(1, K + 1):
        _close(layers_g[k], layers[k], f"layer {k}")
    ncl = losses.NCLLosses(U, I, 0.1, 1e-6, 1.5, 8e-8, 4096)
    ssl_g = ncl.ssl_layer_loss(layers_g[2], layers_g[0], bu.to(cuda), bi.to(cuda))
    np.testing.assert_allclose(ssl_g.item(), ssl_w.item(), rtol=2e-2)         # bf16 logits
    (mean_g.pow(2).sum() * 1e-3 + ssl_g).backward()
    # gradient: fp32 propagation part + bf16-logit contrastive part -> the bf16 tolerance, relative to the gradient's scale
    err = (x0g.grad.cpu() - x0.grad).norm() / x0.grad.norm()
    assert err < 2e-2, f"relative gradient error {err:.3e}"
    # the propagation part alone at the fp32 tolerance
    x1 = x0.detach().clone().requires_grad_(True); x1g = x0.detach().to(cuda).requires_grad_(True)
    w = torch.randn(U + I, d, generator=g)
    cur = x1; acc = x1
    for _ in range(K):
        cur = torch.sparse.mm(adj, cur); acc = acc + cur
    ((acc / (K + 1)) * w).sum().backward()
    (F_.propagate(csr, x1g, K, mode="mean") * w.to(cuda)).sum().backward()
    _close(x1g.grad, x1.grad, "transpose-backward of the 3-layer mean propagation")


def test_cfg3_directau_against_oracle_at_full_size(cuda):
    """directau.py:286-296 (2-layer d = 128 mean propagation) + directau.py:240-251 (alignment / uniformity, B = 2048) on the
    Yelp2018-shaped graph (31,668 x 38,048, 1.56 M edges): GPU path vs CPU torch.sparse + oracle/losses_ref."""
    inter, d, K = synth.config_graph("cfg3")
    U, I = inter.n_users, inter.n_items
    adj = _torch_sparse_adj(inter)
    torch.manual_seed(13)
    x0 = (torch.randn(U + I, d) * 0.1).requires_grad_(True)
    cur, acc = x0, x0
    for _ in range(K):
        cur = torch.sparse.mm(adj, cur); acc = acc + cur
    final_w = acc / (K + 1)
    g = torch.Generator().manual_seed(13)
    bu, bp = torch.randint(0, U, (2048,), generator=g), torch.randint(0, I, (2048,), generator=g)
    want = losses_ref.directau_loss(final_w[:U][bu], final_w[U:][bp], 0.7)
    want.backward()

    csr = CSRGraph.from_pairs(torch.from_numpy(inter.users).to(cuda), torch.from_numpy(inter.items).to(cuda), U, I, norm="sym")
    x0g = x0.detach().to(cuda).requires_grad_(True)
    final_g = F_.propagate(csr, x0g, K, mode="mean")
    _close(final_g, final_w, "2-layer mean propagation, d = 128")
    u_emb, p_emb = F_.gather_rows(final_g[:U], bu.to(cuda)), F_.gather_rows(final_g[U:], bp.to(cuda))
    got = losses.DirectAULosses(0.7).calculate_loss(u_emb, p_emb)
    np.testing.assert_allclose(got.item(), want.item(), rtol=2e-2, atol=2e-3)   # bf16 Gram matrix in the uniformity term
    got.backward()
    err = (x0g.grad.cpu() - x0.grad).norm() / x0.grad.norm()
    assert err < 2e-2, f"relative gradient error {err:.3e}"
