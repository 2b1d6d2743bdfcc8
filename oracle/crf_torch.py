"""Differentiable (torch, float64) restatement of tf.contrib.crf.crf_log_likelihood — used to get
reference GRADIENTS by autograd, the way the reference gets them from tf.gradients
(tools/train_utils.py:314,383).  Values are checked against oracle/crf.py in tests."""
import torch


def crf_log_likelihood(inputs, tag_indices, sequence_lengths, transition_params):
    B, T, K = inputs.shape
    lens = sequence_lengths.long()
    tags = tag_indices.long()
    mask = (torch.arange(T)[None, :] < lens[:, None]).to(inputs.dtype)
    unary = inputs.gather(2, tags[:, :, None])[:, :, 0]
    score = (unary * mask).sum(1)
    if T > 1:
        binary = transition_params[tags[:, :-1], tags[:, 1:]]
        score = score + (binary * mask[:, 1:]).sum(1)
    alpha = inputs[:, 0, :]
    for t in range(1, T):
        new = inputs[:, t, :] + torch.logsumexp(alpha[:, :, None] + transition_params[None], dim=1)
        alpha = torch.where((t < lens)[:, None], new, alpha)
    log_norm = torch.logsumexp(alpha, dim=1)
    zero = torch.zeros_like(log_norm)
    log_norm = torch.where(lens <= 0, zero, log_norm)
    score = torch.where(lens <= 0, zero, score)
    return score - log_norm
