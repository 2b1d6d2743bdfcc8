"""PyTorch-CPU restatement of the neural layers on the reference's bert_bilstm_crf path.

Reference call sites and the third-party semantics they resolve to:
  * bert_encoder      <- tools/layer.py:63-81 -> bert_base.bert.modeling.BertModel (A.3)
  * bilstm            <- tools/layer.py:10-41 -> LSTMCell + bidirectional_dynamic_rnn (A.2)
  * dense             <- tf.layers.dense, e.g. model/bert_bilstm_crf.py:26
  * softlexicon_pool  <- model/bilstm_crf_softlexicon.py:37-44
  * layer_norm_tf     <- tools/transformer/modules.py:40-65

Weights are a dict keyed by the reference's TF variable names (serving_model/*/variables.index).
`dtype` = torch.float64 (truth) or torch.float32 (the reference's arithmetic; also the timed CPU
baseline).  `emulate_bf16=True` rounds GEMM operands / stored activations to bfloat16 at the
points where the CUDA path does (fp32 accumulate), which is what the bf16 configuration of
BASELINE.json config 3 is compared against at tight tolerance.
"""
import math

import torch


def _rb(t, on):
    return t.to(torch.bfloat16).to(t.dtype) if on else t


def gelu(x, variant="tanh"):
    if variant == "erf":
        return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x ** 3)))


def layer_norm(x, gamma, beta, eps):
    mean = x.mean(-1, keepdim=True)
    var = ((x - mean) ** 2).mean(-1, keepdim=True)
    return (x - mean) * torch.rsqrt(var + eps) * gamma + beta


def dense(x, kernel, bias=None):
    y = x @ kernel
    return y if bias is None else y + bias


def bert_encoder(w, input_ids, input_mask, segment_ids, num_layers=12, num_heads=12, dtype=torch.float64,
                 gelu_variant="tanh", emulate_bf16=False, prefix="bert", return_all=False):
    """sequence_output [B,L,H] of BertModel (eval mode: no dropout)."""
    g = lambda name: w[f"{prefix}/{name}"].to(dtype)
    rb = lambda t: _rb(t, emulate_bf16)
    ids = input_ids.long()
    B, L = ids.shape
    seg = torch.zeros_like(ids) if segment_ids is None else segment_ids.long()
    x = g("embeddings/word_embeddings")[ids] + g("embeddings/token_type_embeddings")[seg] \
        + g("embeddings/position_embeddings")[:L][None]
    x = layer_norm(x, g("embeddings/LayerNorm/gamma"), g("embeddings/LayerNorm/beta"), 1e-12)
    H = x.shape[-1]
    dh = H // num_heads
    adder = (1.0 - input_mask.to(dtype))[:, None, None, :] * -10000.0
    outs = []
    for l in range(num_layers):
        p = f"encoder/layer_{l}"
        xb = rb(x)
        q = rb(dense(xb, rb(g(f"{p}/attention/self/query/kernel")), g(f"{p}/attention/self/query/bias")))
        k = rb(dense(xb, rb(g(f"{p}/attention/self/key/kernel")), g(f"{p}/attention/self/key/bias")))
        v = rb(dense(xb, rb(g(f"{p}/attention/self/value/kernel")), g(f"{p}/attention/self/value/bias")))
        sh = lambda t: t.view(B, L, num_heads, dh).permute(0, 2, 1, 3)
        scores = sh(q) @ sh(k).transpose(-1, -2) * (1.0 / math.sqrt(dh)) + adder
        probs = torch.softmax(scores, dim=-1)
        if emulate_bf16:
            # CUDA path: exp(s - max) rounded to bf16 as the P·V operand, row sum kept in fp32
            m = scores.max(-1, keepdim=True).values
            e = torch.exp(scores - m)
            ctx = (rb(e) @ sh(v)) / e.sum(-1, keepdim=True)
        else:
            ctx = probs @ sh(v)
        ctx = rb(ctx.permute(0, 2, 1, 3).reshape(B, L, H))
        a = rb(dense(ctx, rb(g(f"{p}/attention/output/dense/kernel")), g(f"{p}/attention/output/dense/bias")))
        x1 = layer_norm(a + x, g(f"{p}/attention/output/LayerNorm/gamma"), g(f"{p}/attention/output/LayerNorm/beta"), 1e-12)
        i = rb(gelu(dense(rb(x1), rb(g(f"{p}/intermediate/dense/kernel")), g(f"{p}/intermediate/dense/bias")), gelu_variant))
        o = rb(dense(i, rb(g(f"{p}/output/dense/kernel")), g(f"{p}/output/dense/bias")))
        x = layer_norm(o + x1, g(f"{p}/output/LayerNorm/gamma"), g(f"{p}/output/LayerNorm/beta"), 1e-12)
        outs.append(x)
    return outs if return_all else x


def lstm_direction(x, kernel, bias, seq_len, activation="tanh", forget_bias=1.0, reverse=False, emulate_bf16=False):
    """One direction of dynamic_rnn(LSTMCell): kernel [D+H,4H] gate order (i,j,f,o)."""
    B, L, D = x.shape
    H = kernel.shape[1] // 4
    act = torch.relu if activation == "relu" else torch.tanh
    wx, wh = kernel[:D], kernel[D:]
    xproj = _rb(x, emulate_bf16) @ _rb(wx, emulate_bf16) + bias
    h = x.new_zeros(B, H)
    c = x.new_zeros(B, H)
    out = x.new_zeros(B, L, H)
    lens = seq_len.long()
    ar = torch.arange(B)
    for s in range(L):
        active = s < lens
        if not bool(active.any()):
            break
        pos = torch.where(active, (lens - 1 - s) if reverse else torch.full_like(lens, s), torch.zeros_like(lens))
        z = xproj[ar, pos] + h @ wh
        i, j, f, o = z.split(H, dim=1)
        c_new = torch.sigmoid(f + forget_bias) * c + torch.sigmoid(i) * act(j)
        h_new = torch.sigmoid(o) * act(c_new)
        a = active[:, None]
        c = torch.where(a, c_new, c)
        h = torch.where(a, h_new, h)
        idx = ar[active]
        out[idx, pos[active]] = h_new[active]
    return out


def bilstm(x, w, seq_len, activation="tanh", forget_bias=1.0, dtype=torch.float64, emulate_bf16=False,
           prefix="bilstm_layer/bidirectional_rnn"):
    x = x.to(dtype)
    outs = []
    for d, rev in (("fw", False), ("bw", True)):
        k = w[f"{prefix}/{d}/multi_rnn_cell/cell_0/lstm_cell/kernel"].to(dtype)
        b = w[f"{prefix}/{d}/multi_rnn_cell/cell_0/lstm_cell/bias"].to(dtype)
        outs.append(lstm_direction(x, k, b, seq_len, activation, forget_bias, rev, emulate_bf16))
    return torch.cat(outs, dim=-1)


def softlexicon_pool(table, ids, weights, G=4, S=10):
    """[..., G*S] ids/weights -> [..., G*E] (model/bilstm_crf_softlexicon.py:37-44)."""
    E = table.shape[1]
    lead = ids.shape[:-1]
    emb = table[ids.long()] * weights[..., None].to(table.dtype)        # [..., G*S, E]
    emb = emb.view(*lead, G, S, E).sum(-2)
    return emb.reshape(*lead, G * E)
