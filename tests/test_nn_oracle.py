"""CPU: pin the neural oracle (oracle/nn.py) against independent implementations of the same
published algorithms: HuggingFace BertModel (gelu_new = tanh approximation) and torch.nn.LSTM
(packed sequences = dynamic_rnn masking / reverse_sequence semantics)."""
import pytest
import torch

from oracle import nn as onn


def _tf_bert_weights(cfg, seed):
    g = torch.Generator().manual_seed(seed)
    H, I = cfg['hidden_size'], cfg['intermediate_size']
    rn = lambda *s: torch.randn(*s, generator=g) * 0.05
    w = {'bert/embeddings/word_embeddings': rn(cfg['vocab_size'], H),
         'bert/embeddings/token_type_embeddings': rn(2, H),
         'bert/embeddings/position_embeddings': rn(cfg['max_position_embeddings'], H),
         'bert/embeddings/LayerNorm/gamma': 1 + rn(H), 'bert/embeddings/LayerNorm/beta': rn(H)}
    for l in range(cfg['num_hidden_layers']):
        p = f'bert/encoder/layer_{l}'
        for n in ('query', 'key', 'value'):
            w[f'{p}/attention/self/{n}/kernel'] = rn(H, H)
            w[f'{p}/attention/self/{n}/bias'] = rn(H)
        w[f'{p}/attention/output/dense/kernel'] = rn(H, H)
        w[f'{p}/attention/output/dense/bias'] = rn(H)
        w[f'{p}/attention/output/LayerNorm/gamma'] = 1 + rn(H)
        w[f'{p}/attention/output/LayerNorm/beta'] = rn(H)
        w[f'{p}/intermediate/dense/kernel'] = rn(H, I)
        w[f'{p}/intermediate/dense/bias'] = rn(I)
        w[f'{p}/output/dense/kernel'] = rn(I, H)
        w[f'{p}/output/dense/bias'] = rn(H)
        w[f'{p}/output/LayerNorm/gamma'] = 1 + rn(H)
        w[f'{p}/output/LayerNorm/beta'] = rn(H)
    return w


def test_bert_oracle_matches_huggingface():
    transformers = pytest.importorskip("transformers")
    cfg = dict(vocab_size=100, hidden_size=64, num_hidden_layers=2, num_attention_heads=4, intermediate_size=128,
               max_position_embeddings=40)
    w = _tf_bert_weights(cfg, 0)
    hf_cfg = transformers.BertConfig(vocab_size=100, hidden_size=64, num_hidden_layers=2, num_attention_heads=4,
                                     intermediate_size=128, max_position_embeddings=40, type_vocab_size=2,
                                     hidden_act="gelu_new", layer_norm_eps=1e-12, hidden_dropout_prob=0.0,
                                     attention_probs_dropout_prob=0.0)
    hf = transformers.BertModel(hf_cfg, add_pooling_layer=False).double().eval()
    sd = {}
    m = {'embeddings.word_embeddings.weight': 'bert/embeddings/word_embeddings',
         'embeddings.token_type_embeddings.weight': 'bert/embeddings/token_type_embeddings',
         'embeddings.position_embeddings.weight': 'bert/embeddings/position_embeddings',
         'embeddings.LayerNorm.weight': 'bert/embeddings/LayerNorm/gamma',
         'embeddings.LayerNorm.bias': 'bert/embeddings/LayerNorm/beta'}
    for k, v in m.items():
        sd[k] = w[v].double()
    for l in range(2):
        p, q = f'encoder.layer.{l}', f'bert/encoder/layer_{l}'
        for a, b in (('attention.self.query', 'attention/self/query'), ('attention.self.key', 'attention/self/key'),
                     ('attention.self.value', 'attention/self/value'), ('attention.output.dense', 'attention/output/dense'),
                     ('intermediate.dense', 'intermediate/dense'), ('output.dense', 'output/dense')):
            sd[f'{p}.{a}.weight'] = w[f'{q}/{b}/kernel'].double().t().contiguous()
            sd[f'{p}.{a}.bias'] = w[f'{q}/{b}/bias'].double()
        for a, b in (('attention.output.LayerNorm', 'attention/output/LayerNorm'), ('output.LayerNorm', 'output/LayerNorm')):
            sd[f'{p}.{a}.weight'] = w[f'{q}/{b}/gamma'].double()
            sd[f'{p}.{a}.bias'] = w[f'{q}/{b}/beta'].double()
    missing = hf.load_state_dict(sd, strict=False)
    assert not [k for k in missing.missing_keys if 'position_ids' not in k], missing
    g = torch.Generator().manual_seed(1)
    ids = torch.randint(0, 100, (3, 20), generator=g)
    lens = torch.tensor([20, 7, 13])
    mask = (torch.arange(20)[None] < lens[:, None]).long()
    seg = torch.randint(0, 2, (3, 20), generator=g)
    with torch.no_grad():
        ref = hf(input_ids=ids, attention_mask=mask, token_type_ids=seg).last_hidden_state
    out = onn.bert_encoder(w, ids, mask, seg, num_layers=2, num_heads=4, dtype=torch.float64, gelu_variant="tanh")
    valid = mask.bool()
    # HF masks with finfo.min instead of -10000: identical on valid query rows
    assert (out - ref)[valid].abs().max() < 1e-9


@pytest.mark.parametrize("D,H", [(12, 16), (30, 8)])
def test_lstm_oracle_matches_torch_lstm(D, H):
    g = torch.Generator().manual_seed(D)
    B, L = 5, 11
    x = torch.randn(B, L, D, generator=g, dtype=torch.float64)
    lens = torch.tensor([11, 3, 1, 7, 11])
    w = {}
    lstm = torch.nn.LSTM(D, H, batch_first=True, bidirectional=True).double()
    for d, suf in (('fw', ''), ('bw', '_reverse')):
        k = torch.randn(D + H, 4 * H, generator=g, dtype=torch.float64) * 0.3
        b = torch.randn(4 * H, generator=g, dtype=torch.float64) * 0.1
        w[f'bilstm_layer/bidirectional_rnn/{d}/multi_rnn_cell/cell_0/lstm_cell/kernel'] = k
        w[f'bilstm_layer/bidirectional_rnn/{d}/multi_rnn_cell/cell_0/lstm_cell/bias'] = b
        i, j, f, o = k.split(H, dim=1)
        bi, bj, bf, bo = b.split(H)
        kt = torch.cat([i, f, j, o], dim=1)                 # torch order (i, f, g, o)
        bt = torch.cat([bi, bf + 1.0, bj, bo])              # forget_bias folded into the bias
        getattr(lstm, 'weight_ih_l0' + suf).data = kt[:D].t().contiguous()
        getattr(lstm, 'weight_hh_l0' + suf).data = kt[D:].t().contiguous()
        getattr(lstm, 'bias_ih_l0' + suf).data = bt
        getattr(lstm, 'bias_hh_l0' + suf).data = torch.zeros_like(bt)
    packed = torch.nn.utils.rnn.pack_padded_sequence(x, lens, batch_first=True, enforce_sorted=False)
    with torch.no_grad():
        ref, _ = torch.nn.utils.rnn.pad_packed_sequence(lstm(packed)[0], batch_first=True, total_length=L)
    out = onn.bilstm(x, w, lens, 'tanh', 1.0, torch.float64)
    assert (out - ref).abs().max() < 1e-10


def test_softlexicon_oracle_shapes_and_weights():
    table = torch.arange(20, dtype=torch.float64).view(10, 2)
    ids = torch.tensor([[[1, 2, 9, 9, 3, 9, 9, 9]]])        # G=2, S=4
    w = torch.tensor([[[0.5, 0.25, 0, 0, 0.25, 0, 0, 0]]], dtype=torch.float64)
    out = onn.softlexicon_pool(table, ids, w, G=2, S=4)
    exp = torch.cat([0.5 * table[1] + 0.25 * table[2], 0.25 * table[3]])
    assert torch.allclose(out[0, 0], exp)
