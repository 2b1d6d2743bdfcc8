# -*-coding:utf-8 -*-
"""Plugin mirror of reference model/bert_bilstm_crf_softlexicon.py (build_graph :14-67, params :70-86) — one of the
"remaining plugins reusing the same kernels" (SURVEY §8(f) rank 4): BertModel sequence output, the SoftLexicon
gather-and-pool of the B/M/E/S lexicon ids, concat([wh_embedding, bert_embedding]) -> bilstm -> dropout -> dense -> CRF."""
import torch

from .. import autodiff, ops, variables
from ..config import TRAIN_PARAMS as _BASE
from ..tools import layer as _layer
from ..tools.layer import bilstm, crf_decode, crf_layer, dense, dropout, pretrain_bert_embedding


def reshape_input(input_, params):
    return input_.reshape(-1, params['max_seq_len'], int(params['word_enhance_dim'] * params['max_lexicon_len']))


def build_graph(features, labels, params, is_training):
    """
    bert +  bilstm + CRF + softlexicon word enhance
    """
    input_ids = features['token_ids']
    label_ids = features['label_ids']
    input_mask = features['mask']
    segment_ids = features['segment_ids']
    seq_len = features['seq_len']
    B, L = input_ids.shape
    G, S = params['word_enhance_dim'], params['max_lexicon_len']

    # the lexicon features live on the padded [B, L] grid, so the encoder keeps the padded layout here
    pack0, _layer.PACK_SEQUENCES = _layer.PACK_SEQUENCES, False
    try:
        embedding = pretrain_bert_embedding(input_ids, input_mask, segment_ids, params['pretrain_dir'],
                                            params['embedding_dropout'], is_training)
    finally:
        _layer.PACK_SEQUENCES = pack0

    softlexicon_ids = reshape_input(features['softlexicon_ids'], params)
    softlexicon_weights = reshape_input(features['softlexicon_weights'], params)
    init = params['word_embedding']
    softword_embedding = variables.get_variable('word_enhance/softlexicon_embedding', tuple(init.shape), variables.constant(init))
    E = softword_embedding.shape[1]
    wh = ops.softlexicon_pool(softword_embedding, softlexicon_ids, softlexicon_weights, G, S)          # [B, L, G*E]
    tape = autodiff.current() if is_training else None
    if tape is not None:
        store = variables.default_store()

        def pool_bwd(g):
            if g is not None:
                ops.softlexicon_pool_bwd(store.grad('word_enhance/softlexicon_embedding'), softlexicon_ids, softlexicon_weights,
                                         g.contiguous(), G, S)
        tape.record(wh, pool_bwd)
    wh = dropout(wh, params['embedding_dropout'], is_training)
    bert_emb = embedding
    embedding = torch.cat([wh, bert_emb.reshape(B, L, -1)], dim=-1)
    if tape is not None:
        def cat_bwd(g):
            if g is not None:
                tape.add_grad(wh, g[..., :G * E].contiguous())
                tape.add_grad(bert_emb, g[..., G * E:].contiguous().view(bert_emb.shape))
        tape.record(embedding, cat_bwd)

    lstm_output = bilstm(embedding, params['cell_type'], params['rnn_activation'],
                         params['hidden_units_list'], params['keep_prob_list'],
                         params['cell_size'], seq_len, params['dtype'], is_training)
    lstm_output = dropout(lstm_output, params['embedding_dropout'], is_training)

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
    'lr': 5e-6,  # small base learning rate for bert
    'diff_lr_times': {'crf': 500, 'logit': 500, 'lstm': 100, 'word_enhance': 100},  # different lr per-layer
    'embedding_dropout': 0.5,
    'early_stop_ratio': 1  # stop after no improvement after 1.5 epochs
})
