# -*-coding:utf-8 -*-
"""Plugin mirror of reference model/bilstm_crf.py (build_graph :8-44, hyper-params :47-62)."""
import torch

from .. import ops
from ..config import TRAIN_PARAMS as _BASE
from ..tools.layer import bilstm, crf_decode, crf_layer, dense, dropout


def _const_table(params, key):
    """params['embedding'] is a non-trainable numpy constant in the reference; keep one device copy."""
    cache = params.setdefault('_device_consts', {})
    t = cache.get(key)
    if t is None:
        t = torch.as_tensor(params[key], dtype=torch.float32).cuda().contiguous()
        cache[key] = t
    return t


def build_graph(features, labels, params, is_training):
    """
    Use pretrain character embedding + bilstm + crf
    """
    input_ids = features['token_ids']
    label_ids = features['label_ids']
    seq_len = features['seq_len']

    embedding = ops.embedding_lookup(_const_table(params, 'embedding'), input_ids)
    embedding = dropout(embedding, rate=params['embedding_dropout'], is_training=is_training, seed=1234)

    lstm_output = bilstm(embedding, params['cell_type'], params['rnn_activation'],
                         params['hidden_units_list'], params['keep_prob_list'],
                         params['cell_size'], seq_len, params['dtype'], is_training)

    lstm_output = dropout(lstm_output, rate=params['embedding_dropout'], is_training=is_training, seed=1234)

    logits = dense(lstm_output, units=params['label_size'], name='logits', is_training=is_training)

    trans, log_likelihood = crf_layer(logits, label_ids, seq_len, params['label_size'], is_training)
    pred_ids = crf_decode(logits, trans, seq_len, params['idx2tag'], is_training)
    crf_loss = (-log_likelihood).mean()

    return crf_loss, pred_ids


RNN_PARAMS = {
    'cell_type': 'lstm',
    'cell_size': 1,
    'hidden_units_list': [128],
    'keep_prob_list': [1],
    'rnn_activation': 'tanh'
}

TRAIN_PARAMS = dict(_BASE)
TRAIN_PARAMS.update(RNN_PARAMS)
TRAIN_PARAMS.update({
    'lr': 0.005,
    'decay_rate': 0.95,  # lr * decay_rate ^ (global_step / train_steps_per_epoch)
    'embedding_dropout': 0.3,
    'early_stop_ratio': 2  # stop after no improvement after 1.5 epochs
})
