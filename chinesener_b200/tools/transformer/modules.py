# -*-coding:utf-8 -*-
"""Transformer building blocks with the reference's surface (reference tools/transformer/modules.py),
executed at fp32 accuracy: dense layers run as three split-bf16 tcgen05 products (ops.gemm_split_f32),
LayerNorm / attention in fp32."""
import numpy as np
import torch

from ... import ops, variables

FP32_EPS = float(np.finfo(np.float32).eps)


def _dense_pack(store, name, K, N):
    """(w_hi, w_lo) bf16 [N, K8] packs of the TF kernel `name/kernel` [K, N] (K zero-padded to a multiple of 8)."""
    Kp = (K + 7) // 8 * 8

    def build():
        w = store.vars[f"{name}/kernel"]
        if Kp != K:
            w = torch.nn.functional.pad(w, (0, 0, 0, Kp - K))
        w = w.contiguous()
        hi = ops.pack_weight_bf16(w)                                   # [N, Kp] bf16
        lo = ops.pack_weight_bf16((w - hi.float().t()).contiguous())
        return hi, lo, Kp
    return store.cached(("dense_split_pack", name, K, N), build)


def dense_f32(x2d, units, name, relu=False, residual=None, use_bias=True):
    """tf.layers.dense(x, units, name=name) on a [M, K] f32 activation at fp32 accuracy -> f32 [M, units]."""
    M, K = x2d.shape
    assert units % 32 == 0, "split-bf16 dense path needs units % 32 == 0"
    store = variables.default_store()
    store.get_variable(f"{name}/kernel", (K, units), variables.glorot_uniform)
    b = store.get_variable(f"{name}/bias", (units,), variables.zeros) if use_bias else None
    w_hi, w_lo, Kp = _dense_pack(store, name, K, units)
    a_hi, a_lo = ops.split_bf16(x2d, Kp)
    return ops.gemm_split_f32(a_hi, a_lo, w_hi, w_lo, b, residual=residual, relu=relu)


def embedding_project(embedding2d, d_model, name='embedding/dense'):
    """reference modules.py:11-20 — linear map of the raw char(+bichar) embedding to d_model."""
    if embedding2d.shape[-1] == d_model:
        return embedding2d
    return dense_f32(embedding2d, d_model, name)


def layer_norm(x2d, scope):
    """reference modules.py:40-65 — (x-mean)/sqrt(var+eps)*norm_kernel+norm_bias, eps = fp32 machine eps."""
    d = x2d.shape[-1]
    k = variables.get_variable(f"{scope}/layer_normalization/norm_kernel", (d,), variables.ones)
    b = variables.get_variable(f"{scope}/layer_normalization/norm_bias", (d,), variables.zeros)
    out, _ = ops.layernorm(x2d, k, b, eps=FP32_EPS, want_bf16=False)
    return out


def ffn(x2d, ffn_hidden, dropout_rate, is_training, scope):
    """reference modules.py:68-80 — dense-relu-dense (+dropout when training) + add & norm."""
    d_model = x2d.shape[-1]
    y = dense_f32(x2d, ffn_hidden, f"{scope}/ffn/ffn_inner", relu=True)
    y = dense_f32(y, d_model, f"{scope}/ffn/ffn_outer", residual=x2d)          # x + y fused into the epilogue
    return layer_norm(y, f"{scope}/ffn/add_and_norm")


def sinusoidal_positional_encoding(emb_dim, pos_seq):
    """reference modules.py:177-197 — numpy table, cast to float32 exactly like the reference."""
    inv_freq = np.array([1 / (10000 ** ((i - i % 2) / emb_dim)) for i in range(emb_dim)])
    enc = np.einsum('i,j->ij', np.asarray(pos_seq, dtype=np.float64), inv_freq)
    out = np.where(np.arange(emb_dim)[None, :] % 2 == 1, np.cos(enc), np.sin(enc))
    return out.astype(np.float32)
