# -*-coding:utf-8 -*-
"""Plugin mirror of reference model/bert_crf.py (build_graph :8-28, hyper-params :31-33)."""
from ..config import TRAIN_PARAMS as _BASE
from ..tools.layer import crf_decode, crf_layer, dense, pretrain_bert_embedding


def build_graph(features, labels, params, is_training):
    """
    pretrain Bert model output + CRF Layer
    """
    input_ids = features['token_ids']
    label_ids = features['label_ids']
    input_mask = features['mask']
    segment_ids = features['segment_ids']
    seq_len = features['seq_len']
    embedding = pretrain_bert_embedding(input_ids, input_mask, segment_ids, params['pretrain_dir'],
                                        params['embedding_dropout'], is_training)

    logits = dense(embedding, units=params['label_size'], name='logits', is_training=is_training)

    trans, log_likelihood = crf_layer(logits, label_ids, seq_len, params['label_size'], is_training)
    pred_ids = crf_decode(logits, trans, seq_len, params['idx2tag'], is_training)
    crf_loss = (-log_likelihood).mean()
    return crf_loss, pred_ids


TRAIN_PARAMS = dict(_BASE)
TRAIN_PARAMS.update({
    'diff_lr_times': {'crf': 500,  'logit': 500}
})
