"""End-to-end CPU restatement of the reference's build_graph plugins (eval mode).

  bert_bilstm_crf  <- model/bert_bilstm_crf.py:8-34
  bert_crf         <- model/bert_crf.py:8-28
  bilstm_crf       <- model/bilstm_crf.py:8-44
  bilstm_crf_softlexicon <- model/bilstm_crf_softlexicon.py:14-64
Each returns dict(logits, loss, pred_ids, ll).
"""
import numpy as np
import torch

from . import crf, nn


def _crf_tail(logits, w, features):
    lg = logits.detach().to(torch.float32).numpy()
    trans = w["crf_layer/transitions"].to(torch.float32).numpy()
    lens = features["seq_len"].numpy()
    ll = crf.crf_log_likelihood(lg, features["label_ids"].numpy(), lens, trans, dtype=np.float64)
    pred, _ = crf.crf_decode(lg, trans, lens, dtype=np.float32)
    return dict(logits=logits, ll=ll, loss=float(np.mean(-ll)), pred_ids=pred)


def bert_bilstm_crf(w, features, params, dtype=torch.float32, emulate_bf16=False, gelu_variant="tanh"):
    seq = nn.bert_encoder(w, features["token_ids"], features["mask"], features["segment_ids"],
                          num_layers=params.get("num_hidden_layers", 12), num_heads=params.get("num_attention_heads", 12),
                          dtype=dtype, gelu_variant=gelu_variant, emulate_bf16=emulate_bf16)
    lstm = nn.bilstm(seq, w, features["seq_len"], params["rnn_activation"], 1.0, dtype, emulate_bf16)
    logits = nn.dense(lstm, w["logits/kernel"].to(dtype), w["logits/bias"].to(dtype))
    return _crf_tail(logits, w, features)


def bert_crf(w, features, params, dtype=torch.float32, emulate_bf16=False, gelu_variant="tanh"):
    seq = nn.bert_encoder(w, features["token_ids"], features["mask"], features["segment_ids"],
                          num_layers=params.get("num_hidden_layers", 12), num_heads=params.get("num_attention_heads", 12),
                          dtype=dtype, gelu_variant=gelu_variant, emulate_bf16=emulate_bf16)
    # CUDA path feeds the bf16 copy of sequence_output to the label projection
    logits = nn.dense(nn._rb(seq, emulate_bf16), w["logits/kernel"].to(dtype), w["logits/bias"].to(dtype))
    return _crf_tail(logits, w, features)


def bilstm_crf(w, features, params, dtype=torch.float32, emulate_bf16=False):
    emb = torch.as_tensor(params["embedding"]).to(dtype)[features["token_ids"].long()]
    lstm = nn.bilstm(emb, w, features["seq_len"], params["rnn_activation"], 1.0, dtype, emulate_bf16)
    logits = nn.dense(lstm, w["logits/kernel"].to(dtype), w["logits/bias"].to(dtype))
    return _crf_tail(logits, w, features)


def bilstm_crf_softlexicon(w, features, params, dtype=torch.float32, emulate_bf16=False):
    B = features["token_ids"].shape[0]
    L = params["max_seq_len"]
    G, S = params["word_enhance_dim"], params["max_lexicon_len"]
    emb = torch.as_tensor(params["embedding"]).to(dtype)[features["token_ids"].long()]
    ids = features["softlexicon_ids"].view(B, L, G * S)
    wts = features["softlexicon_weights"].view(B, L, G * S)
    wh = nn.softlexicon_pool(w["word_enhance/softlexicon_embedding"].to(dtype), ids, wts, G, S)
    x = torch.cat([wh, emb], dim=-1)
    lstm = nn.bilstm(x, w, features["seq_len"], params["rnn_activation"], 1.0, dtype, emulate_bf16)
    logits = nn.dense(lstm, w["logits/kernel"].to(dtype), w["logits/bias"].to(dtype))
    return _crf_tail(logits, w, features)


def bert_bilstm_crf_softlexicon(w, features, params, dtype=torch.float32, emulate_bf16=False, gelu_variant="tanh"):
    """model/bert_bilstm_crf_softlexicon.py:14-67 (eval mode): concat([pooled lexicon, BERT sequence output]) -> bilstm."""
    B = features["token_ids"].shape[0]
    L = params["max_seq_len"]
    G, S = params["word_enhance_dim"], params["max_lexicon_len"]
    seq = nn.bert_encoder(w, features["token_ids"], features["mask"], features["segment_ids"],
                          num_layers=params.get("num_hidden_layers", 12), num_heads=params.get("num_attention_heads", 12),
                          dtype=dtype, gelu_variant=gelu_variant, emulate_bf16=emulate_bf16)
    ids = features["softlexicon_ids"].view(B, L, G * S)
    wts = features["softlexicon_weights"].view(B, L, G * S)
    wh = nn.softlexicon_pool(w["word_enhance/softlexicon_embedding"].to(dtype), ids, wts.to(dtype), G, S)
    lstm = nn.bilstm(torch.cat([wh, seq], dim=-1), w, features["seq_len"], params["rnn_activation"], 1.0, dtype, emulate_bf16)
    logits = nn.dense(lstm, w["logits/kernel"].to(dtype), w["logits/bias"].to(dtype))
    return _crf_tail(logits, w, features)


def transformer_tener_crf_bichar(w, features, params, dtype=torch.float32):
    """model/transformer_tener_crf_bichar.py:8-42 (eval mode)."""
    from . import transformer as tfm
    char = torch.as_tensor(params["embedding"]).to(dtype)[features["token_ids"].long()]
    bichar = torch.as_tensor(params["bichar_embedding"]).to(dtype)[features["bichar_ids"].long()]
    x = torch.cat([char, bichar], dim=-1)
    if x.shape[-1] != params["d_model"]:
        x = x @ w["embedding/dense/kernel"].to(dtype) + w["embedding/dense/bias"].to(dtype)
    x = tfm.tener_encoder(x, features["seq_len"], w, params["encode_layers"], params["num_head"])
    logits = nn.dense(x, w["logits/kernel"].to(dtype), w["logits/bias"].to(dtype))
    return _crf_tail(logits, w, features)


def transformer_crf_bichar(w, features, params, dtype=torch.float32):
    """model/transformer_crf_bichar.py:8-46 (eval mode): projected char+bichar embedding + sinusoidal positions."""
    from . import transformer as tfm
    char = torch.as_tensor(params["embedding"]).to(dtype)[features["token_ids"].long()]
    bichar = torch.as_tensor(params["bichar_embedding"]).to(dtype)[features["bichar_ids"].long()]
    x = torch.cat([char, bichar], dim=-1)
    if x.shape[-1] != params["d_model"]:
        x = x @ w["embedding/dense/kernel"].to(dtype) + w["embedding/dense/bias"].to(dtype)
    L = params["max_seq_len"]
    x = x + tfm.sinusoidal_positional_encoding(params["d_model"], np.arange(L), dtype)[None]
    x = tfm.transformer_encoder(x, features["seq_len"], w, params["encode_layers"], params["num_head"])
    logits = nn.dense(x, w["logits/kernel"].to(dtype), w["logits/bias"].to(dtype))
    return _crf_tail(logits, w, features)


def bert_cnn_crf(w, features, params, dtype=torch.float32, emulate_bf16=False, gelu_variant="tanh"):
    """model/bert_cnn_crf.py:8-36 (eval mode): tf.layers.conv1d(padding='SAME', relu) per kernel size over the BERT output."""
    seq = nn.bert_encoder(w, features["token_ids"], features["mask"], features["segment_ids"],
                          num_layers=params.get("num_hidden_layers", 12), num_heads=params.get("num_attention_heads", 12),
                          dtype=dtype, gelu_variant=gelu_variant, emulate_bf16=emulate_bf16)
    outs = []
    for filters, k in zip(params["filter_list"], params["kernel_size_list"]):
        kern = w[f"cnn_kernel{k}/kernel"].to(dtype)                    # [k, C, F]
        x = torch.nn.functional.pad(seq.transpose(1, 2), ((k - 1) // 2, k // 2))        # TF 'SAME'
        y = torch.nn.functional.conv1d(x, kern.permute(2, 1, 0), w[f"cnn_kernel{k}/bias"].to(dtype)).transpose(1, 2)
        outs.append(torch.relu(y))
    logits = nn.dense(torch.cat(outs, -1), w["logits/kernel"].to(dtype), w["logits/bias"].to(dtype))
    return _crf_tail(logits, w, features)


def bert_bilstm_crf_mtl(w, features, params, dtype=torch.float32, emulate_bf16=False, gelu_variant="tanh"):
    """model/bert_bilstm_crf_mtl.py:8-66 (eval mode): shared BERT, one BiLSTM+logits+CRF tower per task scope,
    loss = sum_t weight_t * sum(-ll_t[task_ids == t]) / batch, pred_ids picked per sentence by task_ids."""
    seq = nn.bert_encoder(w, features["token_ids"], features["mask"], features["segment_ids"],
                          num_layers=params.get("num_hidden_layers", 12), num_heads=params.get("num_attention_heads", 12),
                          dtype=dtype, gelu_variant=gelu_variant, emulate_bf16=emulate_bf16)
    task_ids = features["task_ids"].numpy()
    lens = features["seq_len"].numpy()
    loss, preds, logits_all, prev = 0.0, [], [], None
    for t, task in enumerate(params["task_list"]):
        lstm = nn.bilstm(seq, w, features["seq_len"], params["rnn_activation"], 1.0, dtype, emulate_bf16,
                         prefix=f"{task}/bilstm_layer/bidirectional_rnn")
        feats = torch.cat([prev, lstm], -1) if (t == 1 and params["asymmetry"]) else lstm
        prev = lstm
        logits = nn.dense(feats, w[f"{task}/logits/kernel"].to(dtype), w[f"{task}/logits/bias"].to(dtype))
        lg = logits.detach().to(torch.float32).numpy()
        trans = w[f"{task}/crf_layer/transitions"].to(torch.float32).numpy()
        K = trans.shape[0]
        ll = crf.crf_log_likelihood(lg, np.minimum(features["label_ids"].numpy(), K - 1), lens, trans, dtype=np.float64)
        loss += params["task_weight"][t] * float(np.sum(-ll[task_ids == t]))
        preds.append(crf.crf_decode(lg, trans, lens, dtype=np.float32)[0])
        logits_all.append(logits)
    pred = np.where((task_ids == 0)[:, None], preds[0], preds[1])
    return dict(logits=logits_all, loss=loss / len(task_ids), pred_ids=pred)


def bert_bilstm_crf_adv(w, features, params, dtype=torch.float32, emulate_bf16=False, gelu_variant="tanh"):
    """model/bert_bilstm_crf_adv.py:9-87 (eval mode; seq_len passed to the task-2 bilstm, see the plugin mirror):
    loss = sum_t weight_t * sum(-ll_t[task_ids == t]) / batch + lambda * mean(xent(discriminator(max_t shared), task_ids))."""
    seq = nn.bert_encoder(w, features["token_ids"], features["mask"], features["segment_ids"],
                          num_layers=params.get("num_hidden_layers", 12), num_heads=params.get("num_attention_heads", 12),
                          dtype=dtype, gelu_variant=gelu_variant, emulate_bf16=emulate_bf16)
    task_ids = features["task_ids"].numpy()
    lens = features["seq_len"].numpy()
    act = params["rnn_activation"]
    share = nn.bilstm(seq, w, features["seq_len"], act, 1.0, dtype, emulate_bf16,
                      prefix="task_discriminator/bilstm_layer/bidirectional_rnn")
    pool = share.max(dim=1).values
    dlogits = nn.dense(pool, w["task_discriminator/logits/kernel"].to(dtype), w["task_discriminator/logits/bias"].to(dtype))
    xent = torch.nn.functional.cross_entropy(dlogits, features["task_ids"].long(), reduction="mean")
    loss, preds = 0.0, []
    for t, task in enumerate(params["task_list"]):
        scope = f"task{t + 1}_{task}"
        lstm = nn.bilstm(seq, w, features["seq_len"], act, 1.0, dtype, emulate_bf16, prefix=f"{scope}/bilstm_layer/bidirectional_rnn")
        logits = nn.dense(torch.cat([share, lstm], -1), w[f"{scope}/logits/kernel"].to(dtype), w[f"{scope}/logits/bias"].to(dtype))
        lg = logits.detach().to(torch.float32).numpy()
        trans = w[f"{scope}/crf_layer/transitions"].to(torch.float32).numpy()
        ll = crf.crf_log_likelihood(lg, np.minimum(features["label_ids"].numpy(), trans.shape[0] - 1), lens, trans, dtype=np.float64)
        loss += params["task_weight"][t] * float(np.sum(-ll[task_ids == t]))
        preds.append(crf.crf_decode(lg, trans, lens, dtype=np.float32)[0])
    pred = np.where((task_ids == 0)[:, None], preds[0], preds[1])
    return dict(loss=loss / len(task_ids) + params["lambda"] * float(xent), adv_loss=float(xent), pred_ids=pred, disc_logits=dlogits)
