"""GPU: the kernels at BASELINE.json's roofline sizes (B = 262 144 sequences, L = 128, K = 10; the [704 370, 50]
SoftLexicon table) through size-independent properties — the oracle finishes only on a sample at these sizes:
  * Viterbi: the returned best score equals the score of the returned path (emissions + transitions re-summed independently),
    the path's log-likelihood under the forward kernel is <= 0 (no path scores above log Z), tags are zero beyond seq_len and
    inside [0, K), two runs are identical, and a random sample of sequences is bit-exact against the numpy oracle;
  * forward-alpha: the probability-domain fast path and the exact log-sum-exp path agree on all 262 144 sequences, the sample
    agrees with the float64 oracle;
  * SoftLexicon pool: linear in the table, rows of zero weight are never read, a sample of tokens equals the gather-multiply-sum
    definition (model/bilstm_crf_softlexicon.py:37-44).
"""
import numpy as np
import pytest
import torch

from chinesener_b200 import ops, synthetic
from oracle import crf

pytestmark = pytest.mark.gpu

B, L, K = 262144, 128, 10


def _problem():
    g = torch.Generator(device="cuda").manual_seed(1234)
    x = torch.randn(B, L, K, device="cuda", generator=g)
    tr = torch.randn(K, K, device="cuda", generator=g) * 0.5
    lens = torch.randint(0, L + 1, (B,), device="cuda", generator=g, dtype=torch.int32)
    lens[:1000] = L
    lens[1000:1010] = 1
    return x, tr, lens


def test_viterbi_full_size_properties():
    x, tr, lens = _problem()
    tags, best = ops.crf_viterbi(x, lens, tr, return_score=True)
    tags2, best2 = ops.crf_viterbi(x, lens, tr, return_score=True)
    assert torch.equal(tags, tags2) and torch.equal(best, best2)                      # deterministic
    assert int(tags.min()) >= 0 and int(tags.max()) < K
    pos = torch.arange(L, device="cuda")[None, :]
    eff = lens.clamp(min=1).long()                                                   # len <= 0 decodes like len 1 (TF)
    valid = pos < eff[:, None]
    assert bool((tags.masked_select(~valid) == 0).all())                              # zero beyond seq_len (tools/layer.py:147)
    # score of the returned path, re-summed with torch ops (fp32, different association: tolerance, not bits)
    t64 = tags.long()
    emit = x.gather(2, t64.unsqueeze(-1)).squeeze(-1).masked_fill(~valid, 0.0).sum(1)
    pair = tr[t64[:, :-1], t64[:, 1:]].masked_fill(~valid[:, 1:], 0.0).sum(1)
    path = emit + pair
    assert torch.allclose(path, best, rtol=1e-4, atol=1e-3), float((path - best).abs().max())
    # no path scores above log Z: the decoded path's log-likelihood is <= 0 (and finite)
    ll, logz, _ = ops.crf_loglik_fwd(x, tags, lens.clamp(min=1), tr)
    assert bool(torch.isfinite(ll).all()) and float(ll.max()) <= 1e-3
    # a random sample against the oracle, bit for bit
    idx = torch.randperm(B, device="cuda")[:384]
    idx[:4] = torch.tensor([0, 1000, 1005, B - 1], device="cuda")
    ref_tags, ref_best = crf.crf_decode(x[idx].cpu().numpy(), tr.cpu().numpy(), lens[idx].cpu().numpy(), dtype=np.float32)
    np.testing.assert_array_equal(tags[idx].cpu().numpy(), ref_tags)
    np.testing.assert_array_equal(best[idx].cpu().numpy(), ref_best.astype(np.float32))


def test_forward_alpha_full_size_paths_agree():
    x, tr, lens = _problem()
    g = torch.Generator(device="cuda").manual_seed(7)
    tags = torch.randint(0, K, (B, L), device="cuda", dtype=torch.int32, generator=g)
    ll_fast, logz_fast, _ = ops.crf_loglik_fwd(x, tags, lens, tr)
    ll_exact, logz_exact, _ = ops.crf_loglik_fwd(x, tags, lens, tr, exact=True)
    assert torch.allclose(logz_fast, logz_exact, rtol=1e-5, atol=2e-3), float((logz_fast - logz_exact).abs().max())
    assert torch.allclose(ll_fast, ll_exact, rtol=1e-5, atol=2e-3)
    assert float(ll_fast.max()) <= 1e-3                                               # a log-probability
    assert bool((logz_fast[lens <= 0] == 0).all())                                    # len <= 0 -> log Z = 0 (TF)
    idx = torch.randperm(B, device="cuda")[:256]
    ref = crf.crf_log_likelihood(x[idx].cpu().numpy(), tags[idx].cpu().numpy(), lens[idx].cpu().numpy(), tr.cpu().numpy())
    np.testing.assert_allclose(ll_fast[idx].cpu().numpy(), ref, rtol=1e-4, atol=1e-3)


def test_softlexicon_pool_config4_table_properties():
    V, E, n_tok = 704370, 50, 2048 * 128
    g = torch.Generator(device="cuda").manual_seed(3)
    table = torch.nn.functional.normalize(torch.randn(V, E, device="cuda", generator=g), dim=1).contiguous()
    ids, w = synthetic.softlexicon_features_device(n_tok, V, realistic=True, seed=5)
    out = ops.softlexicon_pool(table, ids, w, 4, 10)
    assert out.shape == (n_tok, 4 * E)
    out2 = ops.softlexicon_pool(table * 2.0, ids, w, 4, 10)
    assert torch.allclose(out2, 2.0 * out, rtol=1e-6, atol=1e-6)                      # linear in the table
    poisoned = table.clone()
    poisoned[V - 1] = float("nan")                                                    # <PAD> rows carry weight 0: never read
    assert torch.equal(ops.softlexicon_pool(poisoned, ids, w, 4, 10), out)
    idx = torch.randperm(n_tok, device="cuda")[:4096]
    ref = (table[ids[idx].long()] * w[idx][..., None]).view(-1, 4, 10, E).sum(2).reshape(-1, 4 * E)
    torch.testing.assert_close(out[idx], ref, rtol=1e-5, atol=1e-6)
    wsum = w.view(n_tok, 40).sum(1)
    assert torch.allclose(wsum, torch.ones_like(wsum), atol=1e-5)                     # the builder's normalisation
