"""GPU parity of the tcgen05 InfoNCE / DirectAU kernels against the fp64 oracle.  Tolerance: the north star's
rtol 2e-2 for bf16 logits (plus a 2e-2 absolute floor for logits near zero, where a relative bound is meaningless);
losses and gradients rtol 2e-2."""
import numpy as np
import pytest
import torch

from oracle import losses_ref
from recommendation_b200 import functional as F_

pytestmark = pytest.mark.gpu

LOGIT_ATOL = 2e-2
LOGIT_RTOL = 2e-2


def _unit(x):
    return x / x.norm(dim=1, keepdim=True).clamp_min(1e-12)


@pytest.mark.parametrize("m,n,d", [(128, 256, 64), (100, 300, 64), (513, 1000, 128), (4096, 4096, 64), (257, 70000, 64),
                                   (64, 129, 32), (300, 515, 256), (1, 1, 16), (300, 700, 320), (257, 1000, 512), (130, 600, 1024)])
@pytest.mark.parametrize("cos", [True, False])
def test_lse_and_pos_vs_fp64(cuda, m, n, d, cos):
    g = torch.Generator().manual_seed(m * 7 + n)
    scale = 1.0 if cos else 0.3 * (64.0 / max(d, 64)) ** 0.5   # keep un-normalised logits O(1) whatever the width
    q = torch.randn(m, d, generator=g) * scale
    k = torch.randn(n, d, generator=g) * scale
    tau = 0.2 if cos else 0.5
    pos_idx = torch.randint(0, n, (m,), generator=g)
    row, col, pos = F_.infonce_stats_raw(q.to(cuda), k.to(cuda), tau, cos=cos, pos_idx=pos_idx.to(cuda), want_col=True)
    qd, kd = q.double(), k.double()
    if cos:
        qd, kd = _unit(qd), _unit(kd)
    s = qd @ kd.T / tau
    np.testing.assert_allclose(row.cpu().numpy(), torch.logsumexp(s, 1).numpy(), atol=LOGIT_ATOL, rtol=LOGIT_RTOL)
    np.testing.assert_allclose(col.cpu().numpy(), torch.logsumexp(s, 0).numpy(), atol=LOGIT_ATOL, rtol=LOGIT_RTOL)
    np.testing.assert_allclose(pos.cpu().numpy(), s[torch.arange(m), pos_idx].numpy(), atol=LOGIT_ATOL, rtol=LOGIT_RTOL)


# ------------------------------------------------------------------------------------------ backward + scalar losses
def _grad_close(got, want, rtol=2e-2):
    """bf16-logit tolerance (north star: rtol 2e-2) measured against the gradient's own scale: individual entries of a
    softmax gradient cancel to ~0, where an elementwise relative bound is meaningless."""
    got, want = got.detach().cpu().double().numpy(), np.asarray(want, dtype=np.float64)
    scale = np.abs(want).max() + 1e-30
    np.testing.assert_allclose(got, want, rtol=rtol, atol=rtol * scale)
    # and in aggregate much tighter than the per-entry bound
    assert np.linalg.norm(got - want) <= 1e-2 * np.linalg.norm(want) + 1e-12


def _leaf(a, cuda):
    return torch.as_tensor(np.asarray(a), dtype=torch.float32).to(cuda).requires_grad_(True)


def test_infonce_matches_ncl_fixture(cuda, golden):
    z = golden("ncl_losses")
    v1, v2 = _leaf(z["v1"], cuda), _leaf(z["v2"], cuda)
    loss = F_.info_nce(v1, v2, 0.2)
    np.testing.assert_allclose(loss.item(), z["nce"], rtol=2e-2)
    loss.backward()
    _grad_close(v1.grad, z["g_nce_1"]); _grad_close(v2.grad, z["g_nce_2"])
    v1, v2 = _leaf(z["v1"], cuda), _leaf(z["v2"], cuda)
    loss = F_.info_nce(v1 * 0.3, v2 * 0.3, 0.5, b_cos=False)
    np.testing.assert_allclose(loss.item(), z["nce_nocos"], rtol=2e-2)
    loss.backward()
    _grad_close(v1.grad, z["g_nce_nocos_1"]); _grad_close(v2.grad, z["g_nce_nocos_2"])


def test_ssl_layer_loss_matches_ncl_fixture(cuda, golden):
    z = golden("ncl_losses")
    nU = int(z["n_users"])
    ctx, ini = _leaf(z["ctx"], cuda), _leaf(z["ini"], cuda)
    bu, bp = torch.as_tensor(z["bu"]).to(cuda), torch.as_tensor(z["bp"]).to(cuda)
    tau, reg, alpha = float(z["ssl_temp"]), float(z["ssl_reg"]), float(z["alpha"])
    lu = F_.ssl_layer_side(ctx[:nU][bu], ini[:nU], bu, tau)
    li = F_.ssl_layer_side(ctx[nU:][bp], ini[nU:], bp, tau)
    loss = reg * (lu + alpha * li)
    np.testing.assert_allclose(loss.item(), z["ssl"], rtol=2e-2)
    loss.backward()
    _grad_close(ctx.grad, z["g_ssl_ctx"]); _grad_close(ini.grad, z["g_ssl_ini"])


def test_batch_softmax_and_symmetric_match_fixtures(cuda, golden):
    z = golden("ssl4rec_losses")
    a, b = _leaf(z["a"], cuda), _leaf(z["b"], cuda)
    loss = F_.batch_softmax(a, b, 0.2)
    np.testing.assert_allclose(loss.item(), z["batch_softmax"], rtol=2e-2)
    loss.backward()
    _grad_close(a.grad, z["g_bs_a"]); _grad_close(b.grad, z["g_bs_b"])
    z = golden("gcl_losses")
    z1, z2 = _leaf(z["z1"], cuda), _leaf(z["z2"], cuda)
    loss = F_.info_nce_symmetric(z1, z2, 0.2)
    np.testing.assert_allclose(loss.item(), z["info_nce"], rtol=2e-2)
    loss.backward()
    _grad_close(z1.grad, z["g_z1"]); _grad_close(z2.grad, z["g_z2"])


def test_directau_matches_fixture(cuda, golden):
    z = golden("directau_losses")
    gamma = float(z["gamma"])
    xu, xp = _leaf(z["xu"], cuda), _leaf(z["xp"], cuda)
    t3 = F_.directau_terms(xu, xp)
    np.testing.assert_allclose(t3[0].item(), z["align"], rtol=1e-4)
    np.testing.assert_allclose(t3[1].item(), z["unif"], rtol=2e-2)
    loss = t3[0] + gamma * (t3[1] + t3[2]) / 2
    np.testing.assert_allclose(loss.item(), z["calc"], rtol=2e-2)
    loss.backward()
    _grad_close(xu.grad, z["g_calc_u"]); _grad_close(xp.grad, z["g_calc_p"])
    xu2 = _leaf(z["xu"], cuda)
    F_.directau_terms(xu2, xu2.detach())[1].backward()
    _grad_close(xu2.grad, z["g_unif"])


@pytest.mark.parametrize("m,n,d", [(128, 128, 64), (100, 300, 64), (513, 1000, 128), (2048, 2048, 64), (300, 20000, 64),
                                   (64, 129, 32), (300, 515, 256), (257, 1000, 192), (3, 5, 16),
                                   (300, 700, 320), (513, 1000, 512), (260, 300, 1024)])   # d > 256: P materialised + library GEMMs
@pytest.mark.parametrize("mode", ["row", "sym"])
def test_infonce_backward_vs_fp64(cuda, m, n, d, mode):
    """Gradients of sum_i c_i (row_lse_i - pos_i) [+ column term] w.r.t. both operands against fp64 autograd."""
    if mode == "sym" and m != n:
        n = m
    g = torch.Generator().manual_seed(m * 13 + n + d)
    q = torch.randn(m, d, generator=g); k = torch.randn(n, d, generator=g)
    tau = 0.2
    pos_idx = torch.randint(0, n, (m,), generator=g)
    cw = torch.rand(m, generator=g) + 0.5
    qd, kd = q.double().requires_grad_(True), k.double().requires_grad_(True)
    s = _unit(qd) @ _unit(kd).T / tau
    want = (cw.double() * (torch.logsumexp(s, 1) - s[torch.arange(m), pos_idx])).sum()
    if mode == "sym":
        want = want + (torch.logsumexp(s, 0) - s.diagonal()).mean()
    want.backward()
    qc, kc = q.to(cuda).requires_grad_(True), k.to(cuda).requires_grad_(True)
    row, col, pos = F_.infonce_stats(qc, kc, tau, cos=True, pos_idx=pos_idx.to(cuda), want_col=(mode == "sym"))
    got = (cw.to(cuda) * (row - pos)).sum()
    if mode == "sym":
        diag = F_.infonce_stats(qc, kc, tau, cos=True)[2]
        got = got + (col - diag).mean()
    np.testing.assert_allclose(got.item(), want.item(), rtol=2e-2)
    got.backward()
    _grad_close(qc.grad, qd.grad); _grad_close(kc.grad, kd.grad)


@pytest.mark.parametrize("b,d", [(48, 16), (2048, 128), (1000, 64), (130, 256), (700, 512)])
def test_directau_vs_fp64(cuda, b, d):
    g = torch.Generator().manual_seed(b + d)
    x = torch.randn(b, d, generator=g); y = x * 0.5 + torch.randn(b, d, generator=g)
    xd, yd = x.double().requires_grad_(True), y.double().requires_grad_(True)
    want = losses_ref.directau_loss(xd, yd, 1.5)
    want.backward()
    xc, yc = x.to(cuda).requires_grad_(True), y.to(cuda).requires_grad_(True)
    t3 = F_.directau_terms(xc, yc)
    got = t3[0] + 1.5 * (t3[1] + t3[2]) / 2
    np.testing.assert_allclose(got.item(), want.item(), rtol=2e-2, atol=2e-2)
    got.backward()
    _grad_close(xc.grad, xd.grad); _grad_close(yc.grad, yd.grad)
