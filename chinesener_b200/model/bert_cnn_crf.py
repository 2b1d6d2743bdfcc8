# -*-coding:utf-8 -*-
"""Plugin mirror of reference model/bert_cnn_crf.py (build_graph :8-36, params :39-52) — SURVEY §8(f) rank 4:
BertModel sequence output -> conv1d(k=4, 10 filters, SAME, relu) -> dense -> CRF."""
from ..config import TRAIN_PARAMS as _BASE
from ..tools import layer as _layer
from ..tools.layer import cnn_layer, crf_decode, crf_layer, dense, pretrain_bert_embedding


def build_graph(features, labels, params, is_training):
    """
    pretrain Bert Model output + cnn + CRF
    """
    input_ids = features['token_ids']
    label_ids = features['label_ids']
    input_mask = features['mask']
    segment_ids = features['segment_ids']
    seq_len = features['seq_len']
    B, L = input_ids.shape
    # the convolution window crosses seq_len into the [PAD] positions, whose BERT outputs the reference does compute:
    # keep the padded layout (no sequence packing) for this plugin
    pack0, _layer.PACK_SEQUENCES = _layer.PACK_SEQUENCES, False
    try:
        embedding = pretrain_bert_embedding(input_ids, input_mask, segment_ids, params['pretrain_dir'],
                                            params['embedding_dropout'], is_training)
    finally:
        _layer.PACK_SEQUENCES = pack0

    cnn_output = cnn_layer(embedding, params['filter_list'], params['kernel_size_list'],
                           params['cnn_activation'], params['cnn_dropout'], is_training)

    logits = dense(cnn_output, units=params['label_size'], name='logits', is_training=is_training)

    trans, log_likelihood = crf_layer(logits, label_ids, seq_len, params['label_size'], is_training)
    pred_ids = crf_decode(logits, trans, seq_len, params['idx2tag'], is_training)
    crf_loss = (-log_likelihood).mean()

    return crf_loss, pred_ids


CNN_PARAMS = {
    'filter_list': [10],
    'kernel_size_list': [4],
    'padding': 'SAME',
    'cnn_activation': 'relu',
    'cnn_dropout': 0.2
}

TRAIN_PARAMS = dict(_BASE)
TRAIN_PARAMS.update(CNN_PARAMS)
TRAIN_PARAMS.update({
    'diff_lr_times': {'crf': 500, 'logit': 500, 'lstm': 100, 'cnn': 100}
})
