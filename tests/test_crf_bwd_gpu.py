"""GPU: CRF log-likelihood gradient kernel vs forward-backward in float64 (oracle/crf.py)."""
import numpy as np
import pytest
import torch

from chinesener_b200 import ops
from oracle import crf

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,L,K", [(16, 24, 10), (64, 128, 10), (9, 31, 7), (5, 1, 4), (40, 17, 13), (3, 9, 1),
                                   (19000, 8, 10), (5000, 12, 10), (130, 40, 20), (4, 50, 32)])
@pytest.mark.parametrize("wide", [False, True])
def test_crf_backward(B, L, K, wide):
    rng = np.random.default_rng(B + L + K)
    x = rng.normal(size=(B, L, K)).astype(np.float32) * 2
    tr = rng.normal(size=(K, K)).astype(np.float32)
    if wide and K > 2:
        tr[0, 1] = -1e4                                  # forces the exact path
    lens = rng.integers(1, L + 1, size=B).astype(np.int32)
    if B > 2:
        lens[1] = 1
        lens[2] = 0
    tags = rng.integers(0, K, size=(B, L)).astype(np.int32)
    if wide and K > 2:                                   # keep the gold path off the forbidden edge
        tags[tags == 0] = 2
    dx_ref, dtr_ref = crf.crf_marginal_grads(x, tags, lens, tr) if B <= 200 else (None, None)
    xd, td, ld, trd = (torch.from_numpy(a).cuda() for a in (x, tags, lens, tr))
    ll, logz, alpha = ops.crf_loglik_fwd(xd, td, ld, trd, want_alpha=True)
    scale = -1.0 / B
    d_logits, d_trans = ops.crf_loglik_bwd(xd, td, ld, trd, alpha, logz, None, scale)
    if dx_ref is not None:
        np.testing.assert_allclose(d_logits.cpu().numpy(), scale * dx_ref, rtol=2e-3, atol=2e-5)
        np.testing.assert_allclose(d_trans.cpu().numpy(), scale * dtr_ref, rtol=2e-3, atol=2e-4)
    else:
        # size-independent properties: each valid step's d_logits row sums to 0, padded steps are 0
        dl = d_logits.cpu().numpy()
        valid = np.arange(L)[None, :] < lens[:, None]
        assert np.abs(dl.sum(-1)).max() < 1e-5
        assert (dl[~valid] == 0).all()
        # sum of d_trans = -(scale) * 0 net transitions:  sum_ij (count - E[count]) = 0
        assert abs(d_trans.sum().item()) < 1e-2
