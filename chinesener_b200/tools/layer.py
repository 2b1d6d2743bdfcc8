"""Layer library with the reference's function surface (reference tools/layer.py), each a thin
wrapper over the C-ABI kernels.  Same names, argument order and meaning; tensors are torch
CUDA tensors, variables live in `chinesener_b200.variables` under the reference's TF names.
"""
import os

import torch

from .. import autodiff
from .. import bert as _bert
from .. import ops, variables

# Remove padding rows from the token-major activations of the BERT plugins (exact for loss and
# pred_ids: the CRF never reads t >= seq_len).  NER_B200_PACK=0 keeps the padded layout.
PACK_SEQUENCES = os.environ.get("NER_B200_PACK", "1") != "0"
# TRAIN mode of BertModel on the packed layout too (ner_bert_encoder_train_fwd_packed / _bwd_packed)
TRAIN_PACK = os.environ.get("NER_B200_TRAIN_PACK", "1") != "0"
# 'bf16': bf16 tcgen05 operands (BASELINE config 3, the benchmark path); 'fp32': fp32-accurate encoder (config 2's
# "fp32": split-bf16 dense + fp32 attention, emission logits within 1e-3 of the reference).  Estimator sets it from
# params['bert_precision'] around build_graph.
BERT_PRECISION = os.environ.get("NER_B200_BERT_PRECISION", "bf16")


class TrainingPathNotBuilt(NotImplementedError):
    pass


def pretrain_bert_embedding(input_ids, input_mask, segment_ids, pretrain_dir, drop_out, is_training):
    """reference tools/layer.py:63-81 — BertModel(...).get_sequence_output() (+ dropout when training).

    Returns sequence_output [B,L,H] f32; its bf16 copy (what the next GEMM consumes) rides along
    as attribute `.bf16`.
    """
    cfg = _bert.load_bert_config(pretrain_dir)
    B, L = input_ids.shape
    if is_training:
        tape = autodiff.current()
        if tape is None:
            raise TrainingPathNotBuilt("pretrain_bert_embedding(is_training=True) needs an autodiff tape (engine.train_step)")
        # TRAIN_PACK: the encoder runs on the real tokens only; its output keeps the [B,L,H] shape, zero at [PAD] (positions
        # no layer after it reads when PACK_SEQUENCES holds — plugins whose next layer does read them switch it off)
        pack = _bert.make_pack(input_mask) if (PACK_SEQUENCES and TRAIN_PACK and not _bert.PER_KERNEL) else None
        emb = _bert.bert_forward_train(input_ids, input_mask, segment_ids, cfg, variables.default_store(), tape, pack=pack)
        return dropout(emb, rate=drop_out, is_training=True, seed=1234)
    if BERT_PRECISION == 'fp32':
        return _bert.bert_forward_f32(input_ids, input_mask, segment_ids, cfg).view(B, L, -1)
    if PACK_SEQUENCES:
        pack = _bert.make_pack(input_mask)
        x32, x16 = _bert.bert_forward(input_ids, input_mask, segment_ids, cfg, pack=pack)
        emb = x32                      # [total_tokens, H]: packed rows, see PackInfo
        emb.bf16, emb.pack = x16, pack
        return emb
    x32, x16 = _bert.bert_forward(input_ids, input_mask, segment_ids, cfg)
    emb = x32.view(B, L, -1)
    emb.bf16 = x16.view(B, L, -1)
    return emb


def dropout(inputs, rate, is_training, seed=1234):
    """tf.layers.dropout(inputs, rate=rate, seed=seed, training=is_training) — identity in eval mode.

    The mask is a counter-based hash of (seed, global_step, call index, element index): the backward
    pass regenerates it instead of storing it.  (TF's own RNG stream cannot be reproduced; the
    reference's seed=1234 only fixes ITS stream.)
    """
    if not is_training or rate <= 0.0:
        return inputs
    store = variables.default_store()
    store.dropout_calls += 1
    s = (int(seed) * 1000003 + store.global_step) * 1009 + store.dropout_calls
    keep = 1.0 - rate
    x = inputs.contiguous()
    y = ops.dropout(x, keep, s)
    tape = autodiff.current()
    if tape is not None and tape.needs_grad(inputs):
        def bwd(g):
            if g is not None:
                tape.add_grad(inputs, ops.dropout(g.contiguous(), keep, s))
        tape.record(y, bwd)
    return y


def _lstm_pack(store, D, H, scope):
    """bf16 [8H, Dp] input-projection pack (fw | bw), fused bias [8H], fp32 recurrent matrices."""
    Dp = (D + 7) // 8 * 8

    def build():
        ks = [store.vars[f"{scope}/{d}/multi_rnn_cell/cell_0/lstm_cell/kernel"] for d in ("fw", "bw")]
        bs = [store.vars[f"{scope}/{d}/multi_rnn_cell/cell_0/lstm_cell/bias"] for d in ("fw", "bw")]
        wx = torch.cat([k[:D] for k in ks], dim=1)                      # [D, 8H]
        if Dp != D:
            wx = torch.nn.functional.pad(wx, (0, 0, 0, Dp - D))         # zero rows for the K padding
        return dict(wx=ops.pack_weight_bf16(wx.contiguous()), bias=torch.cat(bs).contiguous(),
                    wh_fw=ks[0][D:].contiguous(), wh_bw=ks[1][D:].contiguous(), Dp=Dp)
    return store.cached(("lstm_pack", scope, D, H), build)


def bilstm(embedding, cell_type, activation, hidden_units_list, keep_prob_list, cell_size, seq_len, dtype, is_training):
    """reference tools/layer.py:27-41 — bidirectional_dynamic_rnn over LSTMCell; -> [B,L,2H] f32."""
    if is_training:
        return _bilstm_train(embedding, activation, hidden_units_list, keep_prob_list, cell_size, seq_len)
    if cell_type.lower() != 'lstm':
        raise Exception('Only lstm is built on the sm_100a path (reference models all use cell_type=lstm)')
    if cell_size != 1:
        raise Exception('cell_size must be 1 (every reference model uses a single LSTM layer)')
    pack = getattr(embedding, "pack", None)
    if pack is not None:
        B, L, D = pack.B, pack.L, embedding.shape[-1]
    else:
        B, L, D = embedding.shape
    H = hidden_units_list[0]
    store = variables.default_store()
    scope = variables.scoped("bilstm_layer/bidirectional_rnn")
    for d in ("fw", "bw"):
        store.get_variable(f"{scope}/{d}/multi_rnn_cell/cell_0/lstm_cell/kernel", (D + H, 4 * H), variables.glorot_uniform)
        store.get_variable(f"{scope}/{d}/multi_rnn_cell/cell_0/lstm_cell/bias", (4 * H,), variables.zeros)
    pk = _lstm_pack(store, D, H, scope)
    x16 = getattr(embedding, "bf16", None)
    if x16 is not None and pk["Dp"] == D:
        x16 = x16.reshape(-1, D)
    else:
        x16 = ops.cast_pad_bf16(embedding.reshape(-1, D), pk["Dp"])
    xproj = ops.gemm_bf16(x16, pk["wx"], pk["bias"], epilogue=ops.EPI_F32)
    return ops.bilstm_recurrence(xproj, pk["wh_fw"], pk["wh_bw"], seq_len, B, L, H, activation=activation, forget_bias=1.0,
                                 cu_seqlens=pack.cu_seqlens if pack is not None else None)


def _bilstm_train(embedding, activation, hidden_units_list, keep_prob_list, cell_size, seq_len):
    """Training-mode bilstm(): same kernels on the padded layout, saves gates / cell states / carried
    h and records the BPTT closure.  keep_prob < 1 = DropoutWrapper(output_keep_prob, state_keep_prob)
    (reference tools/layer.py:20-23): counter-based masks inside the recurrence kernels."""
    if cell_size != 1:
        raise Exception('cell_size must be 1')
    keep = float(keep_prob_list[0])
    B, L, D = embedding.shape
    H = hidden_units_list[0]
    store = variables.default_store()
    scope = variables.scoped("bilstm_layer/bidirectional_rnn")
    names = {}
    for d in ("fw", "bw"):
        names[d] = (f"{scope}/{d}/multi_rnn_cell/cell_0/lstm_cell/kernel", f"{scope}/{d}/multi_rnn_cell/cell_0/lstm_cell/bias")
        store.get_variable(names[d][0], (D + H, 4 * H), variables.glorot_uniform)
        store.get_variable(names[d][1], (4 * H,), variables.zeros)
    pk = _lstm_pack(store, D, H, scope)
    x2d = embedding.reshape(B * L, D).contiguous()
    x16 = ops.cast_pad_bf16(x2d, pk["Dp"])
    xproj = ops.gemm_bf16(x16, pk["wx"], pk["bias"], epilogue=ops.EPI_F32)
    store.dropout_calls += 1
    seed = (1234 * 1000003 + store.global_step) * 1009 + store.dropout_calls
    out, gates, cst, hst = ops.bilstm_recurrence(xproj, pk["wh_fw"], pk["wh_bw"], seq_len, B, L, H, activation=activation,
                                                 forget_bias=1.0, save_for_backward=True, keep_prob=keep, seed=seed)
    tape = autodiff.current()
    if tape is not None:
        need_dx = tape.needs_grad(embedding)

        def bwd(g):
            if g is None:
                return
            dxp = ops.bilstm_recurrence_bwd(g.contiguous(), gates, cst, pk["wh_fw"], pk["wh_bw"], seq_len, B, L, H,
                                            activation=activation, keep_prob=keep, seed=seed)
            dxp16 = ops.cast_bf16(dxp)
            gks = [store.grad(names[d][0]) for d in ("fw", "bw")]
            for di, d in enumerate(("fw", "bw")):
                ops.colsum_add(dxp[:, di * 4 * H:(di + 1) * 4 * H], store.grad(names[d][1]), 1.0)
            grouped = pk["Dp"] == D and D % 128 == 0 and H % 128 == 0 and (4 * H) % 256 == 0 and all(k.is_contiguous() for k in gks)
            if grouped:
                # dW_x (both directions) and dW_h (both directions) as ONE grouped launch: token-major operands read in place
                # (ner_wgrad_group_bf16) instead of fp32 transposes + three stream-K GEMMs
                hprev16 = torch.zeros((2, B, L, H), dtype=torch.bfloat16, device=out.device)
                hprev16[0, :, 1:] = hst[:, :-1, :H]          # carried (state-dropped) h of the previous forward step
                hprev16[1, :, :-1] = hst[:, 1:, H:]
                probs = []
                for di in range(2):
                    probs.append((x16, dxp16, di * 4 * H, gks[di][:D]))
                    probs.append((hprev16[di].view(B * L, H), dxp16, di * 4 * H, gks[di][D:]))
                ops.wgrad_group(probs, B * L)
            else:
                dwx = ops.wgrad_gemm(x2d, dxp)                                   # [D, 8H] = x^T d_xproj
                for di, d in enumerate(("fw", "bw")):
                    gk = gks[di]
                    dz = dxp[:, di * 4 * H:(di + 1) * 4 * H]
                    gk[:D] += dwx[:, di * 4 * H:(di + 1) * 4 * H]
                    hprev = torch.zeros((B, L, H), dtype=torch.float32, device=out.device)
                    if di == 0:                       # carried (state-dropped) h of the previous forward step
                        hprev[:, 1:] = hst[:, :-1, :H]
                    else:
                        hprev[:, :-1] = hst[:, 1:, H:]
                    gk[D:] += ops.wgrad_gemm(hprev.view(B * L, H), dz)            # dW_h = h_prev^T dz
            if need_dx:
                wx = torch.cat([store.vars[names[d][0]][:D] for d in ("fw", "bw")], dim=1)   # [D, 8H]: K-major for dx
                Dn = (D + 31) // 32 * 32
                wxp = torch.nn.functional.pad(wx, (0, 0, 0, Dn - D)).to(torch.bfloat16).contiguous()
                dx = ops.gemm_bf16(dxp16, wxp, None, epilogue=ops.EPI_F32)[:, :D]
                tape.add_grad(embedding, dx.reshape(B, L, D).contiguous())
        tape.record(out, bwd)
    return out


def cnn_layer(embedding, filter_list, kernel_size_list, activation, drop_out, is_training):
    """reference tools/layer.py:44-60 — tf.layers.conv1d(padding='SAME') per kernel size (+ dropout), concatenated.
    A kernel-k convolution over [B, L, C] is the label-projection kernel (ner_dense_small_n, filters <= 32) applied to the
    k shifted copies of the sequence laid side by side (TF 'SAME': (k-1)//2 zero rows before, k//2 after); the shifts are
    slices of one zero-padded buffer.  TRAIN: the same kernel's backward tap by tap, the shifted gradient slices are added back."""
    if activation not in ('relu', None):
        raise Exception("cnn_layer: only activation='relu' / None is built")
    B, L, C = embedding.shape
    store, tape = variables.default_store(), (autodiff.current() if is_training else None)
    x = embedding.float() if embedding.dtype != torch.float32 else embedding
    outputs = []
    for filters, k in zip(filter_list, kernel_size_list):
        name = f"cnn_kernel{k}"
        w = store.get_variable(f"{name}/kernel", (k, C, filters), variables.glorot_uniform)       # TF conv1d kernel [k, in, out]
        b = store.get_variable(f"{name}/bias", (filters,), variables.zeros)
        pl, pr = (k - 1) // 2, k // 2
        xp = torch.nn.functional.pad(x, (0, 0, pl, pr))
        cols = torch.cat([xp[:, j:j + L] for j in range(k)], dim=-1).reshape(B * L, k * C).contiguous()
        w2d = w.reshape(k * C, filters)
        pre = ops.dense_small_n(cols, w2d, b)
        out = ops.relu(pre) if activation == 'relu' else pre
        if tape is not None:
            need_dx = tape.needs_grad(embedding)

            def bwd(g, xp=xp, w=w, out=out, name=name, k=k, pl=pl):
                if g is None:
                    return
                g = g.reshape(B * L, -1).contiguous()
                if activation == 'relu':
                    g = ops.relu_bwd(out, g)
                dW, db = store.grad(f"{name}/kernel"), store.grad(f"{name}/bias")
                dxp = torch.zeros_like(xp) if need_dx else None
                for j in range(k):            # one tap at a time: [C, filters] weight + its partial fit the kernel's smem
                    xj = xp[:, j:j + L].reshape(B * L, C)
                    dxj = ops.dense_small_n_bwd(xj, w[j], g, dW[j], db if j == 0 else None, want_dx=need_dx)
                    if need_dx:
                        dxp[:, j:j + L] += dxj.view(B, L, C)
                if need_dx:
                    tape.add_grad(embedding, dxp[:, pl:pl + L].contiguous())
            tape.record(out, bwd)
        out = dropout(out, drop_out, is_training, seed=1234)
        out3 = out.view(B, L, filters)
        if tape is not None:                      # the tape keys on tensor identity: record the reshape
            tape.record(out3, lambda g, out=out: tape.add_grad(out, g.reshape(out.shape)) if g is not None else None)
        outputs.append(out3)
    output = outputs[0] if len(outputs) == 1 else torch.cat(outputs, dim=-1)
    if tape is not None and len(outputs) > 1:
        widths = [o.shape[-1] for o in outputs]

        def cat_bwd(g):
            if g is not None:
                off = 0
                for o, wd in zip(outputs, widths):
                    tape.add_grad(o, g[..., off:off + wd].contiguous())
                    off += wd
        tape.record(output, cat_bwd)
    return output


def dense(inputs, units, name='logits', is_training=False):
    """tf.layers.dense(inputs, units, activation=None, use_bias=True, name=name) for units <= 32."""
    F = inputs.shape[-1]
    lead = inputs.shape[:-1]
    name = variables.scoped(name)
    w = variables.get_variable(f"{name}/kernel", (F, units), variables.glorot_uniform)
    b = variables.get_variable(f"{name}/bias", (units,), variables.zeros)
    tape = autodiff.current()
    if is_training and tape is not None:
        store = variables.default_store()
        x2d = inputs.reshape(-1, F).contiguous()
        out = ops.dense_small_n(x2d, w, b).view(*lead, units)
        need_dx = tape.needs_grad(inputs)

        def bwd(g):
            if g is None:
                return
            dx = ops.dense_small_n_bwd(x2d, w, g.reshape(-1, units).contiguous(), store.grad(f"{name}/kernel"),
                                       store.grad(f"{name}/bias"), want_dx=need_dx)
            if need_dx:
                tape.add_grad(inputs, dx.view(*lead, F))
        tape.record(out, bwd)
        return out
    x = getattr(inputs, "bf16", None)
    x = inputs if x is None else x
    pack = getattr(inputs, "pack", None)
    if pack is not None:
        # packed rows -> padded [B, L, units]; padded positions stay 0 (never read by the CRF)
        out = torch.zeros((pack.B * pack.L, units), dtype=torch.float32, device=x.device)
        ops.dense_small_n(x.reshape(-1, F), w, b, row_map=pack.tok_src, out=out)
        return out.view(pack.B, pack.L, units)
    out = ops.dense_small_n(x.reshape(-1, F), w, b)
    return out.view(*lead, units)


def crf_layer(logits, label_ids, seq_len, label_size, is_training):
    """reference tools/layer.py:112-131 -> (trans, log_likelihood [B])."""
    tname = variables.scoped("crf_layer/transitions")
    trans = variables.get_variable(tname, (label_size, label_size), variables.xavier)
    if label_ids is None:
        return trans, None
    tape = autodiff.current()
    if is_training and tape is not None:
        store = variables.default_store()
        lg = logits.contiguous()
        ll, logz, alpha = ops.crf_loglik_fwd(lg, label_ids, seq_len, trans, want_alpha=True)

        def bwd(g):
            # g = d loss / d ll  [B]; the plugins use loss = mean(-ll)  ->  g = -1/B
            B = lg.shape[0]
            d_ll = g if g is not None else torch.full((B,), -1.0 / B, dtype=torch.float32, device=lg.device)
            d_logits, d_trans = ops.crf_loglik_bwd(lg, label_ids, seq_len, trans, alpha, logz, d_ll.contiguous(), 1.0)
            store.grad(tname).add_(d_trans)
            tape.add_grad(logits, d_logits)
        tape.record(ll, bwd)
        return trans, ll
    # EVAL / PREDICT: built lazily — evaluated when the loss is fetched (EVAL), never in PREDICT
    return trans, variables.Deferred(lambda: ops.crf_loglik_fwd(logits, label_ids, seq_len, trans)[0])


def concat(tensors, is_training=False):
    """tf.concat(tensors, axis=-1) with the tape entry that splits the gradient back."""
    out = torch.cat(tensors, dim=-1)
    tape = autodiff.current() if is_training else None
    if tape is not None:
        widths = [t.shape[-1] for t in tensors]

        def bwd(g):
            if g is not None:
                off = 0
                for t, wd in zip(tensors, widths):
                    tape.add_grad(t, g[..., off:off + wd].contiguous())
                    off += wd
        tape.record(out, bwd)
    return out


def reduce_max_flip(x, shrink, is_training):
    """flip_gradient(tf.reduce_max(x, axis=1), shrink) (reference model/bert_bilstm_crf_adv.py:35-37,
    tools/train_utils.py:47-63): max over time forward; backward sends -shrink * g to every position that attains it."""
    x = x.contiguous()
    y = ops.reduce_max_time(x)
    tape = autodiff.current() if is_training else None
    if tape is not None and tape.needs_grad(x):
        def bwd(g):
            if g is not None:
                tape.add_grad(x, ops.reduce_max_time_bwd(x, y, g.contiguous(), torch.zeros_like(x), scale=-float(shrink)))
        tape.record(y, bwd)
    return y


def softmax_cross_entropy_mean(logits, labels, weight, is_training):
    """weight * tf.reduce_mean(tf.nn.sparse_softmax_cross_entropy_with_logits(labels, logits)); a loss root: TRAIN seeds
    d/d logits = weight / B * (softmax - onehot) on the tape."""
    B = logits.shape[0]
    tape = autodiff.current() if is_training else None
    lg = logits.contiguous()
    if tape is None:
        return ops.softmax_xent(lg, labels).mean() * weight
    xent, dz = ops.softmax_xent(lg, labels, scale=float(weight) / B, want_grad=True)
    loss = xent.mean() * weight
    tape.record(loss, lambda g: tape.add_grad(logits, dz))
    return loss


def masked_task_loss(log_likelihoods, masks, weights, batch_size, is_training):
    """sum_t w_t * sum(-ll_t[mask_t]) / batch — the loss of the multi-task plugins (reference
    model/bert_bilstm_crf_mtl.py:42,61,64).  TRAIN: seeds d loss / d ll_t = -w_t mask_t / batch on the tape."""
    coef = [m.to(torch.float32) * (float(w) / batch_size) for m, w in zip(masks, weights)]
    tape = autodiff.current() if is_training else None
    if tape is None:
        return variables.Deferred(lambda: sum((-(ll.value() if isinstance(ll, variables.Deferred) else ll) * c).sum()
                                              for ll, c in zip(log_likelihoods, coef)))
    loss = sum((-ll * c).sum() for ll, c in zip(log_likelihoods, coef))

    def bwd(g):
        for ll, c in zip(log_likelihoods, coef):
            tape.add_grad(ll, -c)
    tape.record(loss, bwd)
    return loss


def crf_decode(logits, trans, seq_len, idx2tag, is_training, mask=None):
    """reference tools/layer.py:134-149 -> pred_ids [B,L] int32, zero beyond seq_len."""
    return ops.crf_viterbi(logits, seq_len, trans)
