"""CPU: the plain-C restatement of the CRF oracle (oracle/crf_c.c) against the numpy restatement it follows
(oracle/crf.py, itself cross-checked by brute-force enumeration in tests/test_crf_oracle.py): Viterbi tags and scores bit
for bit in float32 — ties, seq_len 0 / 1 / L, L == 1, K == 1 — and the float64 log-likelihood to 1e-12.  The C form is what
bench.py uses to re-run every row of the roofline-sized launches."""
import numpy as np
import pytest

from oracle import crf, native


@pytest.fixture(scope="module", autouse=True)
def _built():
    native.build()
    assert native.available()


def _case(B, T, K, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((B, T, K)).astype(np.float32)
    tr = (rng.standard_normal((K, K)) * 0.5).astype(np.float32)
    lens = rng.integers(0, T + 1, size=B).astype(np.int32)
    lens[0] = T
    if B > 2:
        lens[1], lens[2] = 1, 0
    y = rng.integers(0, K, size=(B, T)).astype(np.int32)
    return x, tr, lens, y


@pytest.mark.parametrize("B,T,K", [(64, 128, 10), (37, 150, 7), (5, 1, 4), (9, 17, 1), (130, 33, 13), (9, 50, 20), (4, 40, 32), (33, 10, 16)])
def test_decode_bit_exact_and_loglik(B, T, K):
    x, tr, lens, y = _case(B, T, K, seed=B + T + K)
    if B > 5:
        x[5] = np.round(x[5])                                   # many exact ties
    tags, best = crf.crf_decode(x, tr, lens, dtype=np.float32)
    tags_c, best_c = native.crf_decode(x, tr, lens)
    np.testing.assert_array_equal(tags_c, tags)
    np.testing.assert_array_equal(best_c, best.astype(np.float32))
    ll = crf.crf_log_likelihood(x, y, lens, tr, dtype=np.float64)
    np.testing.assert_allclose(native.crf_log_likelihood(x, y, lens, tr), ll, rtol=1e-12, atol=1e-12)


def test_ties_resolve_to_lowest_index_and_lengths_are_clipped():
    rng = np.random.default_rng(7)
    x = rng.integers(-1, 2, size=(400, 30, 10)).astype(np.float32)
    tr = rng.integers(-1, 2, size=(10, 10)).astype(np.float32)
    lens = rng.integers(-2, 35, size=400).astype(np.int32)      # below 0 and above L: clipped as crf_decode clips
    tags, best = crf.crf_decode(x, tr, lens, dtype=np.float32)
    tags_c, best_c = native.crf_decode(x, tr, lens)
    np.testing.assert_array_equal(tags_c, tags)
    np.testing.assert_array_equal(best_c, best)
    y = rng.integers(0, 10, size=(400, 30)).astype(np.int32)
    np.testing.assert_allclose(native.crf_log_likelihood(x, y, lens, tr), crf.crf_log_likelihood(x, y, lens, tr), rtol=1e-12, atol=1e-12)


def test_against_brute_force_enumeration():
    rng = np.random.default_rng(3)
    K, T = 3, 5
    x = rng.standard_normal((6, T, K)).astype(np.float32)
    tr = rng.standard_normal((K, K)).astype(np.float32)
    lens = np.array([5, 4, 3, 2, 1, 5], np.int32)
    tags, best = native.crf_decode(x, tr, lens)
    y = rng.integers(0, K, size=(6, T)).astype(np.int32)
    ll = native.crf_log_likelihood(x, y, lens, tr)
    for b in range(6):
        n = int(lens[b])
        logz, m, paths = crf.brute_force(x[b], tr, n)
        assert tuple(tags[b, :n]) in paths and abs(float(best[b]) - m) < 1e-5
        score = sum(float(x[b, t, y[b, t]]) for t in range(n)) + sum(float(tr[y[b, t - 1], y[b, t]]) for t in range(1, n))
        assert abs(ll[b] - (score - logz)) < 1e-9


def test_bad_arguments_are_rejected():
    x, tr, lens, y = _case(4, 6, 3, seed=1)
    y[2, 3] = 7                                                 # tag outside [0, K)
    with pytest.raises(RuntimeError):
        native.crf_log_likelihood(x, y, lens, tr)
    with pytest.raises(RuntimeError):
        native.crf_decode(np.zeros((2, 3, 65), np.float32), np.zeros((65, 65), np.float32), np.array([3, 3], np.int32))


def test_bench_roofline_checker_flags_a_wrong_tag_and_a_wrong_loglik(monkeypatch):
    """bench.crf_sample_check is the untimed checker of the roofline-sized CRF launches: fed the oracle's own outputs it must
    pass on every row, one flipped tag / one shifted log-likelihood must fail it; without the C library it falls back to a
    row sample through the numpy restatement."""
    import sys
    import os
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    B, T, K = 3000, 16, 10
    x, tr, lens, y = _case(B, T, K, seed=5)
    tags, _ = crf.crf_decode(x, tr, lens, dtype=np.float32)
    ll = crf.crf_log_likelihood(x, y, lens, tr).astype(np.float32)
    tx, ttr, tl, ty = torch.from_numpy(x), torch.from_numpy(tr), torch.from_numpy(lens), torch.from_numpy(y)
    tp, tll = torch.from_numpy(tags.astype(np.int32)), torch.from_numpy(ll)
    ok = bench.crf_sample_check(tx, ttr, tl, ty, tll, tp)
    assert ok["rows_checked"] == B and ok["viterbi_bit_exact"] and ok["loglik_within_tolerance"]
    bad_p, bad_l = tp.clone(), tll.clone()
    bad_p[0, 0] = (bad_p[0, 0] + 1) % K
    bad_l[1] += 0.5
    bad = bench.crf_sample_check(tx, ttr, tl, ty, bad_l, bad_p)
    assert not bad["viterbi_bit_exact"] and bad["viterbi_rows_differing"] == 1 and not bad["loglik_within_tolerance"]
    monkeypatch.setattr(native, "available", lambda: False)
    sampled = bench.crf_sample_check(tx, ttr, tl, ty, tll, tp, n=512)
    assert sampled["rows_checked"] == 512 and sampled["viterbi_bit_exact"] and sampled["loglik_within_tolerance"]
