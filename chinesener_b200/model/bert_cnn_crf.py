"""`bert_cnn_crf` (reference model/bert_cnn_crf.py:8-52): BertModel sequence output ->
conv1d(kernel 4, 10 filters, SAME, relu) -> label projection -> CRF."""
from ..tools.layer import cnn_layer
from . import _blocks as nn


def build_graph(features, labels, params, is_training):
    # the convolution window crosses seq_len into [PAD] positions, whose BERT outputs the reference does compute:
    # padded layout (no sequence packing) for this plugin
    hidden = nn.bert_sequence(features, params, is_training, packed=False)
    hidden = cnn_layer(hidden, params['filter_list'], params['kernel_size_list'], params['cnn_activation'],
                       params['cnn_dropout'], is_training)
    return nn.crf_head(hidden, features, params, is_training)


TRAIN_PARAMS = nn.hyper(
    dict(filter_list=[10], kernel_size_list=[4], padding='SAME', cnn_activation='relu', cnn_dropout=0.2),
    diff_lr_times={'crf': 500, 'logit': 500, 'lstm': 100, 'cnn': 100},
)
