# -*-coding:utf-8 -*-
"""`transformer_tener_crf_bichar` (reference model/transformer_tener_crf_bichar.py:8-62): [character | bi-character]
embeddings -> projection to d_model -> TENER encoder (relative-position attention) -> label projection -> CRF."""
import torch

from .. import autodiff, ops
from ..tools.layer import dropout
from ..tools.transformer.encoder import tener_encoder
from ..tools.transformer.modules import embedding_project
from . import _blocks as nn


def build_graph(features, labels, params, is_training):
    input_ids, bichar_ids, seq_len = features['token_ids'], features['bichar_ids'], features['seq_len']
    B, L = input_ids.shape
    char_table = nn.device_constant(params, 'embedding')
    bichar_table = nn.device_constant(params, 'bichar_embedding')
    Ec, Eb = char_table.shape[1], bichar_table.shape[1]
    # concat([char_embedding, bichar_embedding], -1): both lookups write into one buffer
    embedding = torch.empty((B * L, Ec + Eb), dtype=torch.float32, device=input_ids.device)
    ops.embedding_lookup(char_table, input_ids, out=embedding)
    ops.embedding_lookup(bichar_table, bichar_ids, out=embedding, col_offset=Ec)
    embedding = embedding_project(embedding, params['d_model'], is_training=is_training)
    embedding = dropout(embedding, params['embedding_dropout'], is_training, seed=1234)

    transformer_output = tener_encoder(encoder_input=embedding if is_training else embedding.view(B, L, -1), seq_len=seq_len,
                                       max_seq_len=params['max_seq_len'], encode_layers=params['encode_layers'],
                                       num_head=params['num_head'], dropout_rate=params['dropout_rate'],
                                       ffn_hidden=params['ffn_hidden'], is_training=is_training)
    transformer_output = dropout(transformer_output, params['fc_dropout'], is_training, seed=1234)
    if is_training:                      # [B*L, d] -> [B, L, d] as a recorded op (the tape keys on tensor identity)
        tape, out2d = autodiff.current(), transformer_output
        transformer_output = out2d.view(B, L, -1)
        tape.record(transformer_output, lambda g: tape.add_grad(out2d, g.reshape(out2d.shape)) if g is not None else None)

    return nn.crf_head(transformer_output, features, params, is_training)


# MSRA-sized encoder (people_daily in the reference: num_head 5, d_model 200); d_model projects the 50-d giga
# character / bi-character vectors up, num_head must divide it
TRAIN_PARAMS = nn.hyper(
    dict(num_head=8, d_model=160, ffn_hidden=320, encode_layers=2, batch_size=16, wramup_ratio=0.1, epochs=100),
    lr=0.001,
    decay_rate=0.95,
    embedding_dropout=0.3,
    fc_dropout=0.4,
    dropout_rate=0.2,         # transformer sub-layer dropout
    early_stop_ratio=2,
)
