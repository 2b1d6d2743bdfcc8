"""GPU: SoftLexicon gather-and-pool forward / backward vs the CPU oracle."""
import pytest
import torch

from chinesener_b200 import ops, synthetic
from oracle import nn as onn

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("realistic", [True, False])
@pytest.mark.parametrize("B,L,E", [(4, 32, 50), (2, 17, 64), (3, 8, 100)])
def test_pool_forward(B, L, E, realistic):
    V = 5000
    g = torch.Generator().manual_seed(B * L + E)
    table = torch.randn(V, E, generator=g)
    ids, w = synthetic.softlexicon_features(B, L, V, seed=E, realistic=realistic)
    out = ops.softlexicon_pool(table.cuda(), ids.view(B, L, 40).cuda(), w.view(B, L, 40).cuda())
    ref = onn.softlexicon_pool(table.double(), ids.view(B, L, 40), w.view(B, L, 40).double())
    torch.testing.assert_close(out.cpu().double(), ref, rtol=1e-5, atol=1e-5)


def test_pool_into_concat_buffer_and_backward():
    B, L, E, V = 3, 12, 50, 2000
    g = torch.Generator().manual_seed(3)
    table = torch.randn(V, E, generator=g)
    ids, w = synthetic.softlexicon_features(B, L, V, seed=1)
    ids3, w3 = ids.view(B, L, 40), w.view(B, L, 40)
    buf = torch.full((B, L, 4 * E + 50), 7.0, device="cuda")
    ops.softlexicon_pool(table.cuda(), ids3.cuda(), w3.cuda(), out=buf)
    ref = onn.softlexicon_pool(table.double(), ids3, w3.double())
    torch.testing.assert_close(buf[..., :4 * E].cpu().double(), ref, rtol=1e-5, atol=1e-5)
    assert (buf[..., 4 * E:] == 7.0).all()
    # backward: autograd of the oracle
    t = table.double().clone().requires_grad_(True)
    d_out = torch.randn(B, L, 4 * E, generator=g).double()
    (onn.softlexicon_pool(t, ids3, w3.double()) * d_out).sum().backward()
    d_table = torch.zeros(V, E, device="cuda")
    ops.softlexicon_pool_bwd(d_table, ids3.cuda(), w3.cuda(), d_out.float().cuda())
    torch.testing.assert_close(d_table.cpu().double(), t.grad, rtol=1e-4, atol=1e-5)
