# -*-coding:utf-8 -*-
"""`bert_bilstm_crf_softlexicon` (reference model/bert_bilstm_crf_softlexicon.py:14-86):
[pooled B/M/E/S lexicon embedding | BertModel sequence output] -> BiLSTM(200, tanh) -> dropout -> label projection -> CRF."""
import torch

from .. import autodiff, ops, variables
from ..tools.layer import dropout
from . import _blocks as nn


def reshape_input(input_, params):
    return input_.reshape(-1, params['max_seq_len'], int(params['word_enhance_dim'] * params['max_lexicon_len']))


def build_graph(features, labels, params, is_training):
    B, L = features['token_ids'].shape
    G, S = params['word_enhance_dim'], params['max_lexicon_len']
    # the lexicon features live on the padded [B, L] grid, so the encoder keeps the padded layout here
    embedding = nn.bert_sequence(features, params, is_training, packed=False)

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

    hidden = dropout(nn.recurrent(embedding, features, params, is_training), params['embedding_dropout'], is_training)
    return nn.crf_head(hidden, features, params, is_training)


TRAIN_PARAMS = nn.hyper(
    dict(cell_type='lstm', cell_size=1, hidden_units_list=[200],     # 200 for MSRA (128 for people_daily in the reference)
         keep_prob_list=[0.9], rnn_activation='tanh'),
    lr=5e-6,                                                         # BERT-sized base rate, multiplied per layer group below
    diff_lr_times={'crf': 500, 'logit': 500, 'lstm': 100, 'word_enhance': 100},
    embedding_dropout=0.5,
    early_stop_ratio=1,
)
