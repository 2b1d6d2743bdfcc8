# -*-coding:utf-8 -*-
"""reference tools/transformer/encoder.py:21-33 — tener_encoder."""
from .modules import ffn, multi_head_attention
from .tener import relative_multi_head_attention


def tener_encoder(encoder_input, seq_len, max_seq_len, encode_layers, num_head, dropout_rate, ffn_hidden, is_training):
    if encoder_input.dim() == 2:          # TRAIN passes the [B*L, d] activation itself (the tape keys on tensor identity)
        x, (B, L), d = encoder_input, (seq_len.shape[0], max_seq_len), encoder_input.shape[-1]
    else:
        B, L, d = encoder_input.shape
        x = encoder_input.reshape(B * L, d)
    for i in range(encode_layers):
        scope = f"encoding/self_attention_layer_{i}"
        x = relative_multi_head_attention(x, seq_len, B, L, num_head, dropout_rate, is_training, scope)
        x = ffn(x, ffn_hidden, dropout_rate, is_training, scope)
    return x if encoder_input.dim() == 2 else x.view(B, L, d)


def transformer_encoder(encoder_input, seq_len, max_seq_len, encode_layers, num_head, dropout_rate, ffn_hidden, is_training):
    """reference tools/transformer/encoder.py:6-19 — absolute-position encoder (scaled dot-product attention)."""
    if encoder_input.dim() == 2:
        x, (B, L), d = encoder_input, (seq_len.shape[0], max_seq_len), encoder_input.shape[-1]
    else:
        B, L, d = encoder_input.shape
        x = encoder_input.reshape(B * L, d)
    for i in range(encode_layers):
        scope = f"encoding/self_attention_layer_{i}"
        x = multi_head_attention(x, seq_len, B, L, num_head, dropout_rate, is_training, scope)
        x = ffn(x, ffn_hidden, dropout_rate, is_training, scope)
    return x if encoder_input.dim() == 2 else x.view(B, L, d)
