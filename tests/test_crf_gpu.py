"""GPU parity: CUDA CRF kernels (through the C-ABI) vs the numpy oracle.

Viterbi tag indices must be BIT-EXACT (integer output); log-likelihood within 1e-4 relative
(+1e-4 absolute) of the float64 oracle.
"""
import numpy as np
import pytest
import torch

from chinesener_b200 import ops
from oracle import crf

pytestmark = pytest.mark.gpu


def _case(B, L, K, seed, ragged=True, scale=2.0):
    rng = np.random.default_rng(seed)
    x = (rng.normal(size=(B, L, K)) * scale).astype(np.float32)
    tr = (rng.normal(size=(K, K))).astype(np.float32)
    lens = rng.integers(1, L + 1, size=B).astype(np.int32) if ragged else np.full(B, L, np.int32)
    tags = rng.integers(0, K, size=(B, L)).astype(np.int32)
    return x, tr, lens, tags


@pytest.mark.parametrize("B,L,K", [(64, 128, 10), (8, 64, 10), (37, 150, 7), (5, 1, 4), (3, 17, 1), (130, 33, 13),
                                   (9, 50, 20), (4, 40, 32), (300, 21, 3), (70, 150, 10), (33, 9, 16), (6, 256, 10)])
def test_viterbi_bit_exact(B, L, K):
    x, tr, lens, _ = _case(B, L, K, seed=B * 1000 + L * 10 + K)
    lens[0] = L
    if B > 2:
        lens[1] = 1
        lens[2] = 0                                      # TF quirk: decodes like len 1
    ref_tags, ref_best = crf.crf_decode(x, tr, lens, dtype=np.float32)
    tags, best = ops.crf_viterbi(torch.from_numpy(x).cuda(), torch.from_numpy(lens).cuda(),
                                 torch.from_numpy(tr).cuda(), return_score=True)
    assert tags.dtype == torch.int32
    np.testing.assert_array_equal(tags.cpu().numpy(), ref_tags)
    np.testing.assert_array_equal(best.cpu().numpy(), ref_best.astype(np.float32))


def test_viterbi_ties_resolve_to_lowest_index():
    B, L, K = 40, 30, 10
    rng = np.random.default_rng(7)
    x = rng.integers(-1, 2, size=(B, L, K)).astype(np.float32)     # many exact ties
    tr = rng.integers(-1, 2, size=(K, K)).astype(np.float32)
    lens = rng.integers(1, L + 1, size=B).astype(np.int32)
    ref_tags, _ = crf.crf_decode(x, tr, lens, dtype=np.float32)
    tags = ops.crf_viterbi(torch.from_numpy(x).cuda(), torch.from_numpy(lens).cuda(), torch.from_numpy(tr).cuda())
    np.testing.assert_array_equal(tags.cpu().numpy(), ref_tags)


def test_viterbi_mid_batch_uses_thread_per_sequence_kernel():
    B, L, K = 5000, 40, 10                              # above the lane-per-tag threshold, NT=32 CTAs
    x, tr, lens, _ = _case(B, L, K, seed=17)
    ref_tags, _ = crf.crf_decode(x, tr, lens, dtype=np.float32)
    tags = ops.crf_viterbi(torch.from_numpy(x).cuda(), torch.from_numpy(lens).cuda(), torch.from_numpy(tr).cuda())
    np.testing.assert_array_equal(tags.cpu().numpy(), ref_tags)


def test_viterbi_large_batch_uses_128_thread_ctas():
    B, L, K = 148 * 64 + 77, 128, 10                    # > big-batch threshold, ragged tail CTA
    x, tr, lens, _ = _case(B, L, K, seed=11)
    ref_tags, _ = crf.crf_decode(x, tr, lens, dtype=np.float32)
    tags = ops.crf_viterbi(torch.from_numpy(x).cuda(), torch.from_numpy(lens).cuda(), torch.from_numpy(tr).cuda())
    np.testing.assert_array_equal(tags.cpu().numpy(), ref_tags)


@pytest.mark.parametrize("B,L,K", [(64, 128, 10), (8, 64, 10), (37, 150, 7), (5, 1, 4), (3, 17, 1), (130, 33, 13),
                                   (9, 50, 20), (4, 40, 32), (19000, 24, 10), (6, 256, 10), (33, 10, 16), (5000, 16, 10)])
@pytest.mark.parametrize("exact", [False, True])
def test_loglik_forward(B, L, K, exact):
    x, tr, lens, tags = _case(B, L, K, seed=B + L + K)
    if B > 2:
        lens[1] = 1
        lens[2] = 0
    ref = crf.crf_log_likelihood(x, tags, lens, tr, dtype=np.float64)
    ref_logz = crf.crf_log_norm(x.astype(np.float64), lens, tr.astype(np.float64))
    ll, logz, _ = ops.crf_loglik_fwd(torch.from_numpy(x).cuda(), torch.from_numpy(tags).cuda(),
                                     torch.from_numpy(lens).cuda(), torch.from_numpy(tr).cuda(), exact=exact)
    np.testing.assert_allclose(ll.cpu().numpy(), ref, rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(logz.cpu().numpy(), ref_logz, rtol=1e-4, atol=1e-4)


def test_loglik_wide_transitions_fall_back_to_exact_path():
    B, L, K = 16, 40, 10
    x, tr, lens, tags = _case(B, L, K, seed=3)
    tr[2, 5] = -1e4                                      # hard "forbidden" transition
    tr[7, 1] = -np.inf
    tags[:, :] = 0
    ref = crf.crf_log_likelihood(x, tags, lens, tr, dtype=np.float64)
    ll, _, _ = ops.crf_loglik_fwd(torch.from_numpy(x).cuda(), torch.from_numpy(tags).cuda(),
                                  torch.from_numpy(lens).cuda(), torch.from_numpy(tr).cuda())
    np.testing.assert_allclose(ll.cpu().numpy(), ref, rtol=1e-4, atol=1e-4)


def test_loglik_alpha_workspace():
    B, L, K = 20, 31, 10
    x, tr, lens, tags = _case(B, L, K, seed=9)
    _, alphas = crf.crf_log_norm(x.astype(np.float64), lens, tr.astype(np.float64), return_alphas=True)
    _, _, a = ops.crf_loglik_fwd(torch.from_numpy(x).cuda(), torch.from_numpy(tags).cuda(),
                                 torch.from_numpy(lens).cuda(), torch.from_numpy(tr).cuda(), want_alpha=True)
    a = a.cpu().numpy()
    for b in range(B):
        n = int(lens[b])
        np.testing.assert_allclose(a[b, :n], alphas[b, :n], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("variant", [0, 1, 2])
@pytest.mark.parametrize("B,L,K", [(148 * 64 + 77, 128, 10), (9600, 37, 7), (9500, 50, 16), (9601, 21, 12), (9490, 9, 3),
                                   (9500, 12, 5), (9533, 16, 9), (9480, 20, 8), (9479, 8, 11), (9600, 256, 10), (9500, 6, 2)])
def test_viterbi_large_batch_kernels_bit_exact(monkeypatch, variant, B, L, K):
    """variant 0 = the pipe-balanced TMA kernel (max tree + first-equal index, all backpointers in shared memory; taken
    when L*K % 4 == 0, otherwise the call falls through to variant 2), variant 2 = the occupancy-first kernel (low
    backpointer nibbles parked in the tags_out slab), variant 1 = the all-on-chip kernel (the fallback for K > 16): all
    must reproduce the oracle's tags and scores, ragged lengths and a partial tail CTA included."""
    monkeypatch.setenv("NER_CRF_VIT_VARIANT", str(variant))
    x, tr, lens, _ = _case(B, L, K, seed=variant * 100 + K)
    lens[0], lens[1], lens[2] = L, 1, 0
    x[5] = np.round(x[5])                                # a row with many exact ties
    ref_tags, ref_best = crf.crf_decode(x, tr, lens, dtype=np.float32)
    tags, best = ops.crf_viterbi(torch.from_numpy(x).cuda(), torch.from_numpy(lens).cuda(),
                                 torch.from_numpy(tr).cuda(), return_score=True)
    np.testing.assert_array_equal(tags.cpu().numpy(), ref_tags)
    np.testing.assert_array_equal(best.cpu().numpy(), ref_best.astype(np.float32))


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("K", [10, 7, 13])
def test_loglik_large_batch_configurations(monkeypatch, variant, K):
    monkeypatch.setenv("NER_CRF_FWD_VARIANT", str(variant))
    B, L = 19000, 40
    x, tr, lens, tags = _case(B, L, K, seed=variant + K)
    ref = crf.crf_log_likelihood(x, tags, lens, tr, dtype=np.float64)
    ll, _, _ = ops.crf_loglik_fwd(torch.from_numpy(x).cuda(), torch.from_numpy(tags).cuda(),
                                  torch.from_numpy(lens).cuda(), torch.from_numpy(tr).cuda())
    np.testing.assert_allclose(ll.cpu().numpy(), ref, rtol=1e-4, atol=1e-4)
