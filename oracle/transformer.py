"""PyTorch-CPU restatement of the reference's TENER encoder (fully spelled out in the reference:
tools/transformer/{modules,tener,encoder}.py) — eval mode.

Variable names follow the TF scopes the reference builds:
  encoding/self_attention_layer_{i}/multi_head_attention/{pre_value_project,pre_query_project,post_linear_project}/{kernel,bias}
  encoding/self_attention_layer_{i}/multi_head_attention/tener_relative_attention/{content_bias_u,positional_bias_v}
  encoding/self_attention_layer_{i}/multi_head_attention/add_and_norm/layer_normalization/{norm_kernel,norm_bias}
  encoding/self_attention_layer_{i}/ffn/{ffn_inner,ffn_outer}/{kernel,bias}
  encoding/self_attention_layer_{i}/ffn/add_and_norm/layer_normalization/{norm_kernel,norm_bias}
"""
import numpy as np
import torch

FP32_EPS = float(np.finfo(np.float32).eps)
MASK_ADD = float(-2 ** 32 + 1)          # tools/transformer/modules.py:126


def layer_norm_tf(x, kernel, bias):
    """tools/transformer/modules.py:40-65: (x-mean)/sqrt(var+eps) * kernel + bias, eps = fp32 machine eps."""
    mean = x.mean(-1, keepdim=True)
    var = ((x - mean) ** 2).mean(-1, keepdim=True)
    return kernel * ((x - mean) / (var + FP32_EPS) ** 0.5) + bias


def sinusoidal_positional_encoding(emb_dim, pos_seq, dtype=torch.float64):
    """tools/transformer/modules.py:177-197: even index sin, odd index cos, inv_freq_i = 10000^-((i - i%2)/d)."""
    inv_freq = np.array([1 / (10000 ** ((i - i % 2) / emb_dim)) for i in range(emb_dim)])
    enc = np.einsum('i,j->ij', np.asarray(pos_seq, dtype=np.float64), inv_freq)
    out = np.where(np.arange(emb_dim)[None, :] % 2 == 1, np.cos(enc), np.sin(enc))
    # the reference casts the table to float32 (tf.cast(tf.constant(...), tf.float32))
    return torch.from_numpy(out.astype(np.float32)).to(dtype)


def shift(BD):
    """tools/transformer/tener.py:51-74, literally (zero-pad column + reshape trick)."""
    b, n, q, p = BD.shape
    BD = torch.cat([BD, torch.zeros_like(BD[:, :, :, :1])], dim=-1)
    BD = BD.reshape(b, n, p + 1, q)[:, :, :-1]
    BD = BD.reshape(b, n, q, p)
    return BD[:, :, :, q:]


def relative_multi_head_attention(x, mask, w, prefix, num_head):
    """tools/transformer/tener.py:77-119 with key = value = query = x."""
    B, L, d_model = x.shape
    dh = d_model // num_head
    g = lambda n: w[f"{prefix}/{n}"].to(x.dtype)
    new_value = x @ g("pre_value_project/kernel") + g("pre_value_project/bias")
    new_query = x @ g("pre_query_project/kernel") + g("pre_query_project/bias")
    split = lambda t: t.view(B, L, num_head, dh).permute(0, 2, 1, 3)          # [B,n,L,dh]
    key, query, value = split(x), split(new_query), split(new_value)
    pos_emb = sinusoidal_positional_encoding(dh, np.arange(-L, L), x.dtype)     # [2L, dh]
    u = g("tener_relative_attention/content_bias_u")
    v = g("tener_relative_attention/positional_bias_v")
    AC = torch.einsum('bnqd,bnkd->bnqk', query + u[:, None, :], key)
    BD = shift(torch.einsum('bnqd,ld->bnql', query + v[:, None, :], pos_emb))
    weight = AC + BD + (1 - mask.to(x.dtype))[:, None, None, :] * MASK_ADD
    weight = torch.softmax(weight, dim=-1)
    out = (weight @ value).permute(0, 2, 1, 3).reshape(B, L, d_model)
    out = out @ g("post_linear_project/kernel") + g("post_linear_project/bias")
    return layer_norm_tf(x + out, g("add_and_norm/layer_normalization/norm_kernel"),
                         g("add_and_norm/layer_normalization/norm_bias"))


def multi_head_attention(x, mask, w, prefix, num_head):
    """tools/transformer/modules.py:129-175 with key = value = query = x (scaled dot product, projected key)."""
    B, L, d_model = x.shape
    dh = d_model // num_head
    g = lambda n: w[f"{prefix}/{n}"].to(x.dtype)
    proj = lambda n: x @ g(f"{n}/kernel") + g(f"{n}/bias")
    split = lambda t: t.view(B, L, num_head, dh).permute(0, 2, 1, 3)
    key, value, query = split(proj("pre_key_project")), split(proj("pre_value_project")), split(proj("pre_query_project"))
    weight = query @ key.transpose(-1, -2) / (dh ** 0.5) + (1 - mask.to(x.dtype))[:, None, None, :] * MASK_ADD
    out = (torch.softmax(weight, dim=-1) @ value).permute(0, 2, 1, 3).reshape(B, L, d_model)
    out = out @ g("post_linear_project/kernel") + g("post_linear_project/bias")
    return layer_norm_tf(x + out, g("add_and_norm/layer_normalization/norm_kernel"),
                         g("add_and_norm/layer_normalization/norm_bias"))


def transformer_encoder(x, seq_len, w, encode_layers, num_head):
    """tools/transformer/encoder.py:6-19."""
    B, L, _ = x.shape
    mask = (torch.arange(L)[None, :] < seq_len.long()[:, None])
    for i in range(encode_layers):
        p = f"encoding/self_attention_layer_{i}"
        x = multi_head_attention(x, mask, w, f"{p}/multi_head_attention", num_head)
        x = ffn(x, w, f"{p}/ffn")
    return x


def ffn(x, w, prefix):
    g = lambda n: w[f"{prefix}/{n}"].to(x.dtype)
    y = torch.relu(x @ g("ffn_inner/kernel") + g("ffn_inner/bias"))
    y = y @ g("ffn_outer/kernel") + g("ffn_outer/bias")
    return layer_norm_tf(x + y, g("add_and_norm/layer_normalization/norm_kernel"),
                         g("add_and_norm/layer_normalization/norm_bias"))


def tener_encoder(x, seq_len, w, encode_layers, num_head):
    """tools/transformer/encoder.py:21-33."""
    B, L, _ = x.shape
    mask = (torch.arange(L)[None, :] < seq_len.long()[:, None])
    for i in range(encode_layers):
        p = f"encoding/self_attention_layer_{i}"
        x = relative_multi_head_attention(x, mask, w, f"{p}/multi_head_attention", num_head)
        x = ffn(x, w, f"{p}/ffn")
    return x
