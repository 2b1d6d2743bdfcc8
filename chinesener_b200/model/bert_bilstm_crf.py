"""`bert_bilstm_crf` — the north-star plugin (reference model/bert_bilstm_crf.py:8-48): BertModel sequence output ->
BiLSTM(128, relu) -> label projection -> CRF.  PREDICT also has a one-call executor (fastpath.FUSED_PREDICT) that runs the
same kernels; this graph is the definition it is tested against."""
from . import _blocks as nn


def build_graph(features, labels, params, is_training):
    hidden = nn.bert_sequence(features, params, is_training)
    hidden = nn.recurrent(hidden, features, params, is_training)
    return nn.crf_head(hidden, features, params, is_training)


TRAIN_PARAMS = nn.hyper(
    dict(cell_type='lstm', cell_size=1, hidden_units_list=[128], keep_prob_list=[0.8], rnn_activation='relu'),
    diff_lr_times={'crf': 500, 'logit': 500, 'lstm': 100},     # learning-rate multipliers of the non-BERT variables
)
