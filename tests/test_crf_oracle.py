"""CPU: the CRF oracle (oracle/crf.py) against brute-force path enumeration and its own invariants."""
import numpy as np
import pytest

from oracle import crf


@pytest.mark.parametrize("K,L", [(2, 1), (3, 4), (4, 5), (5, 3)])
def test_oracle_matches_brute_force(K, L):
    rng = np.random.default_rng(K * 100 + L)
    B = 6
    x = rng.normal(size=(B, L, K)) * 2
    tr = rng.normal(size=(K, K))
    lens = rng.integers(1, L + 1, size=B)
    tags = rng.integers(0, K, size=(B, L))
    ll = crf.crf_log_likelihood(x, tags, lens, tr, dtype=np.float64)
    dec, best = crf.crf_decode(x, tr, lens, dtype=np.float64)
    for b in range(B):
        logz, m, paths = crf.brute_force(x[b], tr, int(lens[b]))
        score = crf.crf_sequence_score(x[b:b + 1], tags[b:b + 1], lens[b:b + 1], tr)[0]
        assert abs(ll[b] - (score - logz)) < 1e-9
        assert abs(best[b] - m) < 1e-9
        assert tuple(int(v) for v in dec[b, :lens[b]]) in paths
        assert (dec[b, lens[b]:] == 0).all()


def test_decode_edge_cases_and_ties():
    K, L = 4, 6
    x = np.zeros((3, L, K), dtype=np.float32)           # all ties -> lowest index everywhere
    tr = np.zeros((K, K), dtype=np.float32)
    dec, best = crf.crf_decode(x, tr, np.array([6, 0, 1]))
    assert (dec == 0).all() and (best == 0).all()
    x[1, 0, 2] = 1.0                                     # len 0 behaves like len 1: tags[0] = argmax x[0]
    dec, _ = crf.crf_decode(x, tr, np.array([6, 0, 1]))
    assert dec[1, 0] == 2 and (dec[1, 1:] == 0).all()


def test_loglik_empty_sequence_is_zero():
    rng = np.random.default_rng(3)
    x = rng.normal(size=(2, 5, 3))
    tr = rng.normal(size=(3, 3))
    ll = crf.crf_log_likelihood(x, np.zeros((2, 5), dtype=int), np.array([0, -2]), tr)
    assert np.allclose(ll, 0)


def test_marginal_grads_match_finite_differences():
    rng = np.random.default_rng(5)
    B, L, K = 3, 5, 4
    x = rng.normal(size=(B, L, K))
    tr = rng.normal(size=(K, K))
    lens = np.array([5, 2, 4])
    tags = rng.integers(0, K, size=(B, L))
    dx, dtr = crf.crf_marginal_grads(x, tags, lens, tr)
    f = lambda xx, tt: crf.crf_log_likelihood(xx, tags, lens, tt).sum()
    eps = 1e-6
    for idx in [(0, 0, 1), (1, 1, 2), (2, 3, 0), (1, 4, 3)]:
        xp = x.copy(); xp[idx] += eps
        xm = x.copy(); xm[idx] -= eps
        assert abs((f(xp, tr) - f(xm, tr)) / (2 * eps) - dx[idx]) < 1e-6
    for idx in [(0, 0), (1, 3), (2, 2)]:
        tp = tr.copy(); tp[idx] += eps
        tm = tr.copy(); tm[idx] -= eps
        assert abs((f(x, tp) - f(x, tm)) / (2 * eps) - dtr[idx]) < 1e-6
