"""`bert_crf` (reference model/bert_crf.py:8-33): BertModel sequence output -> label projection -> CRF."""
from . import _blocks as nn


def build_graph(features, labels, params, is_training):
    return nn.crf_head(nn.bert_sequence(features, params, is_training), features, params, is_training)


TRAIN_PARAMS = nn.hyper(diff_lr_times={'crf': 500, 'logit': 500})
