"""Shared pieces of the plugin graphs.  Every plugin keeps the reference contract — `build_graph(features, labels, params,
is_training) -> (loss, pred_ids)` plus a module-level `TRAIN_PARAMS` — and composes these blocks; the blocks call the
reference-shaped layer functions of `tools/layer.py`, which hold the kernels."""
import torch

from ..config import TRAIN_PARAMS as BASE_TRAIN_PARAMS
from ..tools import layer as L


def hyper(*groups, **overrides):
    """TRAIN_PARAMS of a plugin: the shared defaults, then each group dict, then keyword overrides."""
    out = dict(BASE_TRAIN_PARAMS)
    for g in groups:
        out.update(g)
    out.update(overrides)
    return out


def device_constant(params, key):
    """A non-trainable numpy table of `params` (pretrained embeddings): one cached device copy per params dict."""
    cache = params.setdefault('_device_consts', {})
    if key not in cache:
        cache[key] = torch.as_tensor(params[key], dtype=torch.float32).cuda().contiguous()
    return cache[key]


def bert_sequence(features, params, is_training, packed=True):
    """BertModel sequence output for the batch (reference tools/layer.py:63-81).  packed=False keeps the padded layout for
    plugins whose next layer reads [PAD] positions."""
    args = (features['token_ids'], features['mask'], features['segment_ids'], params['pretrain_dir'],
            params['embedding_dropout'], is_training)
    if packed:
        return L.pretrain_bert_embedding(*args)
    keep, L.PACK_SEQUENCES = L.PACK_SEQUENCES, False
    try:
        return L.pretrain_bert_embedding(*args)
    finally:
        L.PACK_SEQUENCES = keep


def recurrent(x, features, params, is_training):
    """The bidirectional LSTM block configured by the plugin's RNN hyper-parameters."""
    return L.bilstm(x, params['cell_type'], params['rnn_activation'], params['hidden_units_list'], params['keep_prob_list'],
                    params['cell_size'], features['seq_len'], params['dtype'], is_training)


def crf_head(hidden, features, params, is_training, name='logits'):
    """Label projection + CRF: -> (mean negative log-likelihood, Viterbi tags)."""
    n_tags, lengths = params['label_size'], features['seq_len']
    emissions = L.dense(hidden, units=n_tags, name=name, is_training=is_training)
    transitions, log_lik = L.crf_layer(emissions, features['label_ids'], lengths, n_tags, is_training)
    tags = L.crf_decode(emissions, transitions, lengths, params['idx2tag'], is_training)
    return (-log_lik).mean(), tags
