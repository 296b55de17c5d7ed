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
                                   (64, 129, 32), (300, 515, 256), (1, 1, 16)])
@pytest.mark.parametrize("cos", [True, False])
def test_lse_and_pos_vs_fp64(cuda, m, n, d, cos):
    g = torch.Generator().manual_seed(m * 7 + n)
    q = torch.randn(m, d, generator=g) * (1.0 if cos else 0.3)
    k = torch.randn(n, d, generator=g) * (1.0 if cos else 0.3)
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
