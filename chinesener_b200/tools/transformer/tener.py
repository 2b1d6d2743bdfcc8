# -*-coding:utf-8 -*-
"""TENER relative multi-head attention (reference tools/transformer/tener.py:77-119) on the
fp32 attention kernel (csrc/attention_f32.cu)."""
import numpy as np
import torch

from ... import autodiff, ops, variables
from .modules import add_and_norm_train, dense_f32, dense_train, layer_norm, sinusoidal_positional_encoding

_rel_cache = {}


def _rel_table(L, dh, device):
    key = (L, dh, str(device))
    t = _rel_cache.get(key)
    if t is None:
        t = torch.from_numpy(sinusoidal_positional_encoding(dh, np.arange(-L, L))).to(device).contiguous()
        _rel_cache[key] = t
    return t


def relative_multi_head_attention(x2d, seq_len, B, L, num_head, dropout_rate, is_training, scope):
    """key = value = query = x.  K is NOT projected, scores are unscaled (tener.py:94-96, 12-48)."""
    d_model = x2d.shape[-1]
    dh = d_model // num_head
    p = f"{scope}/multi_head_attention"
    store = variables.default_store()
    if is_training:
        return _rel_mha_train(x2d, seq_len, B, L, num_head, dh, dropout_rate, p, store)
    # value / query projections (two tf.layers.dense in the reference; same arithmetic per output column)
    v = dense_f32(x2d, d_model, f"{p}/pre_value_project")
    q = dense_f32(x2d, d_model, f"{p}/pre_query_project")
    u = store.get_variable(f"{p}/tener_relative_attention/content_bias_u", (num_head, dh), variables.xavier)
    vb = store.get_variable(f"{p}/tener_relative_attention/positional_bias_v", (num_head, dh), variables.xavier)
    ctx, _, _ = ops.attention_f32(q, x2d, v, seq_len, B, L, num_head, dh, scale=1.0, bias_u=u, bias_v=vb,
                                  rel_table=_rel_table(L, dh, x2d.device))
    y = dense_f32(ctx, d_model, f"{p}/post_linear_project", residual=x2d)       # query + weighted_val
    return layer_norm(y, f"{p}/add_and_norm")


def _rel_mha_train(x2d, seq_len, B, L, num_head, dh, dropout_rate, p, store):
    """TRAIN mode: same forward kernels; the attention core's backward is ner_attention_f32_bwd (scores recomputed),
    dK flows straight into x because the key is not projected (tener.py:94)."""
    tape = autodiff.current()
    d_model = x2d.shape[-1]
    v = dense_train(x2d, d_model, f"{p}/pre_value_project")
    q = dense_train(x2d, d_model, f"{p}/pre_query_project")
    un, vn = f"{p}/tener_relative_attention/content_bias_u", f"{p}/tener_relative_attention/positional_bias_v"
    u = store.get_variable(un, (num_head, dh), variables.xavier)
    vb = store.get_variable(vn, (num_head, dh), variables.xavier)
    rel = _rel_table(L, dh, x2d.device)
    ctx, _, _ = ops.attention_f32(q, x2d, v, seq_len, B, L, num_head, dh, scale=1.0, bias_u=u, bias_v=vb, rel_table=rel)

    def bwd(g):
        if g is None:
            return
        dq, dk, dv, du, dvb = ops.attention_f32_bwd(q, x2d, v, seq_len, B, L, num_head, dh, g.contiguous(), scale=1.0, bias_u=u,
                                                    bias_v=vb, rel_table=rel)
        store.grad(un).add_(du)
        store.grad(vn).add_(dvb)
        tape.add_grad(q, dq)
        tape.add_grad(v, dv)
        tape.add_grad(x2d, dk)
    tape.record(ctx, bwd)
    y = dense_train(ctx, d_model, f"{p}/post_linear_project")
    return add_and_norm_train(x2d, y, f"{p}/add_and_norm", dropout_rate)
