"""numpy restatement of tf.contrib.crf (TensorFlow 1.14) as the reference calls it.

Reference call sites: tools/layer.py:122-127 (crf_log_likelihood) and tools/layer.py:140-142
(crf_decode).  Semantics per SURVEY.md Appendix A.1.  All functions take a `dtype`
(np.float32 reproduces the reference's fp32 arithmetic order; np.float64 is the "truth" used
for tolerance tests).
"""
import itertools

import numpy as np


def _logsumexp(x, axis):
    # tf.reduce_logsumexp: max-subtract with a finite-max guard.
    m = np.max(x, axis=axis, keepdims=True)
    m = np.where(np.isfinite(m), m, 0)
    with np.errstate(divide="ignore"):
        out = np.log(np.sum(np.exp(x - m), axis=axis, keepdims=True)) + m
    return np.squeeze(out, axis=axis)


def crf_sequence_score(inputs, tag_indices, sequence_lengths, transition_params):
    """crf_unary_score + crf_binary_score (masked by t < len and t+1 < len)."""
    B, T, K = inputs.shape
    lens = np.asarray(sequence_lengths)
    mask = (np.arange(T)[None, :] < lens[:, None]).astype(inputs.dtype)
    tags = np.asarray(tag_indices)
    unary = np.take_along_axis(inputs, tags[:, :, None], axis=2)[:, :, 0]
    score = np.sum(unary * mask, axis=1)
    if T > 1:
        binary = transition_params[tags[:, :-1], tags[:, 1:]]
        score = score + np.sum(binary * mask[:, 1:], axis=1)
    return score.astype(inputs.dtype)


def crf_log_norm(inputs, sequence_lengths, transition_params, return_alphas=False):
    """Forward-alpha recursion; state frozen past the end; 0 where len <= 0."""
    B, T, K = inputs.shape
    lens = np.asarray(sequence_lengths)
    alpha = inputs[:, 0, :].copy()
    alphas = np.zeros_like(inputs)
    alphas[:, 0] = alpha
    for t in range(1, T):
        scores = alpha[:, :, None] + transition_params[None, :, :]
        new = inputs[:, t, :] + _logsumexp(scores, axis=1)
        valid = (t < lens)[:, None]
        alpha = np.where(valid, new, alpha).astype(inputs.dtype)
        alphas[:, t] = alpha
    log_norm = _logsumexp(alpha, axis=1)
    log_norm = np.where(lens <= 0, 0, log_norm).astype(inputs.dtype)
    return (log_norm, alphas) if return_alphas else log_norm


def crf_log_likelihood(inputs, tag_indices, sequence_lengths, transition_params, dtype=np.float64):
    inputs = np.asarray(inputs, dtype=dtype)
    trans = np.asarray(transition_params, dtype=dtype)
    score = crf_sequence_score(inputs, tag_indices, sequence_lengths, trans)
    score = np.where(np.asarray(sequence_lengths) <= 0, 0, score) if inputs.shape[1] == 1 else score
    return (score - crf_log_norm(inputs, sequence_lengths, trans)).astype(dtype)


def crf_decode(potentials, transition_params, sequence_length, dtype=np.float32):
    """Viterbi: (tags [B,T] int32 zero past len, best_score [B]).  First max on ties (np.argmax)."""
    x = np.asarray(potentials, dtype=dtype)
    trans = np.asarray(transition_params, dtype=dtype)
    B, T, K = x.shape
    lens = np.clip(np.asarray(sequence_length), 0, T)
    if T == 1:
        return np.argmax(x[:, 0, :], axis=1).astype(np.int32)[:, None], np.max(x[:, 0, :], axis=1)
    s = x[:, 0, :].copy()
    bps = np.zeros((B, T, K), dtype=np.int32)
    lm1 = np.maximum(lens - 1, 0)
    for t in range(1, T):
        m = s[:, :, None] + trans[None, :, :]            # (s[i] + trans[i,j]) first
        bp = np.argmax(m, axis=1).astype(np.int32)
        new = x[:, t, :] + np.max(m, axis=1)             # then + potentials
        valid = ((t - 1) < lm1)[:, None]                 # dynamic_rnn over inputs[1:], length len-1
        s = np.where(valid, new, s).astype(dtype)
        bps[:, t] = np.where(valid, bp, 0)
    last = np.argmax(s, axis=1).astype(np.int32)
    best = np.max(s, axis=1)
    tags = np.zeros((B, T), dtype=np.int32)
    for b in range(B):
        n = max(int(lens[b]), 1)                         # len 0 decodes like len 1 (reverse_sequence no-op)
        y = int(last[b])
        tags[b, n - 1] = y
        for t in range(n - 1, 0, -1):
            y = int(bps[b, t, y])
            tags[b, t - 1] = y
    return tags, best


def crf_marginal_grads(inputs, tag_indices, sequence_lengths, transition_params):
    """d(sum_b ll[b]) / d inputs and / d trans by forward-backward, float64."""
    x = np.asarray(inputs, dtype=np.float64)
    trans = np.asarray(transition_params, dtype=np.float64)
    B, T, K = x.shape
    lens = np.asarray(sequence_lengths)
    tags = np.asarray(tag_indices)
    dx = np.zeros_like(x)
    dtr = np.zeros_like(trans)
    for b in range(B):
        n = int(min(max(lens[b], 0), T))
        if n <= 0:
            continue
        a = np.zeros((n, K))
        a[0] = x[b, 0]
        for t in range(1, n):
            a[t] = x[b, t] + _logsumexp(a[t - 1][:, None] + trans, axis=0)
        logz = _logsumexp(a[n - 1], axis=0)
        beta = np.zeros((n, K))
        for t in range(n - 2, -1, -1):
            beta[t] = _logsumexp(trans + (x[b, t + 1] + beta[t + 1])[None, :], axis=1)
        for t in range(n):
            p = np.exp(a[t] + beta[t] - logz)
            dx[b, t] = -p
            dx[b, t, tags[b, t]] += 1.0
            if t >= 1:
                pair = np.exp(a[t - 1][:, None] + trans + (x[b, t] + beta[t])[None, :] - logz)
                dtr -= pair
                dtr[tags[b, t - 1], tags[b, t]] += 1.0
    return dx, dtr


def brute_force(inputs, transition_params, length):
    """Enumerate all K^length paths of ONE sequence: (logZ, best_path (first in lexicographic-max
    order is NOT defined -> returns best score and the set of argmax paths)."""
    x = np.asarray(inputs, dtype=np.float64)
    trans = np.asarray(transition_params, dtype=np.float64)
    K = x.shape[1]
    scores = []
    paths = list(itertools.product(range(K), repeat=length))
    for p in paths:
        s = sum(x[t, p[t]] for t in range(length)) + sum(trans[p[t - 1], p[t]] for t in range(1, length))
        scores.append(s)
    scores = np.array(scores)
    m = scores.max()
    logz = m + np.log(np.exp(scores - m).sum())
    best = [paths[i] for i in np.flatnonzero(scores >= m - 1e-12)]
    return logz, m, best
