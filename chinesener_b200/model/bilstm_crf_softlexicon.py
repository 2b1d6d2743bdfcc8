# -*-coding:utf-8 -*-
"""`bilstm_crf_softlexicon` (reference model/bilstm_crf_softlexicon.py:14-84): [pooled B/M/E/S lexicon embedding | frozen
character embedding] -> BiLSTM(200, tanh) -> label projection -> CRF."""
import torch

from .. import autodiff, ops, variables
from ..tools.layer import dropout
from . import _blocks as nn


def reshape_input(input_, params):
    return input_.reshape(-1, params['max_seq_len'], int(params['word_enhance_dim'] * params['max_lexicon_len']))


def build_graph(features, labels, params, is_training):
    input_ids = features['token_ids']
    # flat [B, L * G * S] features -> [B, L, G * S]
    softlexicon_ids = reshape_input(features['softlexicon_ids'], params)
    softlexicon_weights = reshape_input(features['softlexicon_weights'], params)
    B, L = input_ids.shape
    G, S = params['word_enhance_dim'], params['max_lexicon_len']

    char_table = nn.device_constant(params, 'embedding')
    init = params['word_embedding']
    softword_embedding = variables.get_variable('word_enhance/softlexicon_embedding', tuple(init.shape),
                                                variables.constant(init))
    E, Ec = softword_embedding.shape[1], char_table.shape[1]
    # concat([wh_embedding, embedding], -1) without materialising the pieces: both kernels write
    # straight into the [B, L, G*E + Ec] buffer
    if not is_training:
        embedding = torch.empty((B, L, G * E + Ec), dtype=torch.float32, device=input_ids.device)
        ops.softlexicon_pool(softword_embedding, softlexicon_ids, softlexicon_weights, G, S, out=embedding)
        ops.embedding_lookup(char_table, input_ids, out=embedding, col_offset=G * E)
    else:
        # TRAIN (reference :26-49): dropout on the constant char embedding, dropout on the pooled lexicon
        # embedding, concat; only the lexicon table is a variable, its gradient is the scatter-add of the pool
        char = dropout(ops.embedding_lookup(char_table, input_ids), params['embedding_dropout'], is_training)
        wh = ops.softlexicon_pool(softword_embedding, softlexicon_ids, softlexicon_weights, G, S)
        tape = autodiff.current()
        if tape is not None:
            store = variables.default_store()

            def pool_bwd(g):
                if g is not None:
                    ops.softlexicon_pool_bwd(store.grad('word_enhance/softlexicon_embedding'), softlexicon_ids,
                                             softlexicon_weights, g.contiguous(), G, S)
            tape.record(wh, pool_bwd)
        wh = dropout(wh, params['embedding_dropout'], is_training)
        embedding = torch.cat([wh, char], dim=-1)
        if tape is not None:
            wh_d = wh

            def cat_bwd(g):
                if g is not None:
                    tape.add_grad(wh_d, g[..., :G * E].contiguous())
            tape.record(embedding, cat_bwd)

    return nn.crf_head(nn.recurrent(embedding, features, params, is_training), features, params, is_training)


TRAIN_PARAMS = nn.hyper(
    dict(cell_type='lstm', cell_size=1, hidden_units_list=[200],     # 200 for MSRA (128 for people_daily in the reference)
         keep_prob_list=[0.9], rnn_activation='tanh'),
    lr=0.0015,
    decay_rate=0.95,
    embedding_dropout=0.5,
    early_stop_ratio=1,
)
