"""Thin functional wrappers over the C-ABI (one per exported kernel family).

Each function checks devices/dtypes/contiguity, allocates outputs with torch, and forwards raw
pointers + the current CUDA stream to libner_b200.so.  Nothing here computes on the CPU.
"""
import torch

from . import _lib
from ._lib import check, lib, ptr, require_cuda, stream

EPI_F32, EPI_BF16, EPI_GELU_TANH_BF16, EPI_GELU_ERF_BF16, EPI_RELU_BF16, EPI_RES_F32 = range(6)


def _i32(t):
    return t if t.dtype == torch.int32 else t.to(torch.int32)


# --------------------------------------------------------------------------- CRF
def crf_viterbi(logits, seq_len, trans, return_score=False):
    """tf.contrib.crf.crf_decode (reference tools/layer.py:140).  -> tags [B,L] int32 (+ best_score [B])."""
    require_cuda(logits, seq_len, trans)
    assert logits.dtype == torch.float32 and trans.dtype == torch.float32
    B, L, K = logits.shape
    assert trans.shape == (K, K)
    seq_len = _i32(seq_len)
    tags = torch.empty((B, L), dtype=torch.int32, device=logits.device)
    score = torch.empty((B,), dtype=torch.float32, device=logits.device) if return_score else None
    check(lib().ner_crf_viterbi(ptr(logits), ptr(seq_len), ptr(trans), ptr(tags), ptr(score), B, L, K, stream()))
    return (tags, score) if return_score else tags


def crf_loglik_fwd(logits, tags, seq_len, trans, want_alpha=False, exact=False):
    """tf.contrib.crf.crf_log_likelihood forward (reference tools/layer.py:122). -> ll [B], logz [B], alpha|None."""
    require_cuda(logits, tags, seq_len, trans)
    assert logits.dtype == torch.float32 and trans.dtype == torch.float32
    B, L, K = logits.shape
    tags, seq_len = _i32(tags), _i32(seq_len)
    ll = torch.empty((B,), dtype=torch.float32, device=logits.device)
    logz = torch.empty((B,), dtype=torch.float32, device=logits.device)
    alpha = torch.empty((B, L, K), dtype=torch.float32, device=logits.device) if want_alpha else None
    check(lib().ner_crf_loglik_fwd(ptr(logits), ptr(tags), ptr(seq_len), ptr(trans), ptr(ll), ptr(logz), ptr(alpha),
                                   B, L, K, 1 if exact else 0, stream()))
    return ll, logz, alpha


# --------------------------------------------------------------------------- dense (tcgen05)
def gemm_bf16(a, wt, bias=None, residual=None, epilogue=EPI_BF16, tile_n=0, out=None):
    """out[M,N] = epilogue(a[M,K] @ wt[N,K]^T + bias).  a, wt bf16; see ner_gemm_bf16."""
    require_cuda(a, wt, bias, residual, out)
    assert a.dtype == torch.bfloat16 and wt.dtype == torch.bfloat16
    M, K = a.shape
    N, K2 = wt.shape
    assert K == K2
    odt = torch.float32 if epilogue in (EPI_F32, EPI_RES_F32) else torch.bfloat16
    if out is None:
        out = torch.empty((M, N), dtype=odt, device=a.device)
    assert out.dtype == odt and out.shape == (M, N)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N
    if residual is not None:
        assert residual.dtype == torch.float32 and residual.shape == (M, N)
    check(lib().ner_gemm_bf16(ptr(a), ptr(wt), ptr(bias), ptr(residual), ptr(out), M, N, K, epilogue, tile_n, stream()))
    return out
