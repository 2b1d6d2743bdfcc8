# -*-coding:utf-8 -*-
"""Plugin mirror of reference model/bilstm_crf_softlexicon.py (build_graph :14-64, params :67-84)."""
import torch

from .. import autodiff, ops, variables
from ..config import TRAIN_PARAMS as _BASE
from ..tools.layer import bilstm, crf_decode, crf_layer, dense, dropout
from .bilstm_crf import _const_table


def reshape_input(input_, params):
    return input_.reshape(-1, params['max_seq_len'], int(params['word_enhance_dim'] * params['max_lexicon_len']))


def build_graph(features, labels, params, is_training):
    """
    Giga pretrain character embedding + bilstm + CRF + softlexicon word enhance
    """
    input_ids = features['token_ids']
    label_ids = features['label_ids']
    seq_len = features['seq_len']
    # reshape -> batch, max_seq_len, word_enhance_dim * max_lexicon_len
    softlexicon_ids = reshape_input(features['softlexicon_ids'], params)
    softlexicon_weights = reshape_input(features['softlexicon_weights'], params)
    B, L = input_ids.shape
    G, S = params['word_enhance_dim'], params['max_lexicon_len']

    char_table = _const_table(params, 'embedding')
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

    lstm_output = bilstm(embedding, params['cell_type'], params['rnn_activation'],
                         params['hidden_units_list'], params['keep_prob_list'],
                         params['cell_size'], seq_len, params['dtype'], is_training)

    logits = dense(lstm_output, units=params['label_size'], name='logits', is_training=is_training)

    trans, log_likelihood = crf_layer(logits, label_ids, seq_len, params['label_size'], is_training)
    pred_ids = crf_decode(logits, trans, seq_len, params['idx2tag'], is_training)
    crf_loss = (-log_likelihood).mean()

    return crf_loss, pred_ids


RNN_PARAMS = {
    'cell_type': 'lstm',
    'cell_size': 1,
    'hidden_units_list': [200],  # 128 for people_daily ,200 for msra
    'keep_prob_list': [0.9],
    'rnn_activation': 'tanh',
}

TRAIN_PARAMS = dict(_BASE)
TRAIN_PARAMS.update(RNN_PARAMS)
TRAIN_PARAMS.update({
    'lr': 0.0015,
    'decay_rate': 0.95,  # lr * decay_rate ^ (global_step / train_steps_per_epoch)
    'embedding_dropout': 0.5,
    'early_stop_ratio': 1  # stop after no improvement after 1.5 epochs
})
