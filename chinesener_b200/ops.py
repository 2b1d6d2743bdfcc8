"""Thin functional wrappers over the C-ABI (one per exported kernel family).

Each function checks devices/dtypes/contiguity, allocates outputs with torch, and forwards raw
pointers + the current CUDA stream to libner_b200.so.  Nothing here computes on the CPU.
"""
import torch

from . import _lib
from ._lib import NerB200Error, check, lib, ptr, require_cuda, stream

EPI_F32, EPI_BF16, EPI_GELU_TANH_BF16, EPI_GELU_ERF_BF16, EPI_RELU_BF16, EPI_RES_F32, EPI_RES_RELU_F32 = range(7)
EPI_DIAG_DISCARD = 99
TILE_2CTA_128, TILE_2CTA_256 = 1128, 1256   # CTA-pair (cta_group::2) tiles of ner_gemm_bf16
TILE_SK_128, TILE_SK_256 = 2128, 2256       # stream-K scheduling of 128 x {128,256} tiles
TILE_AUTO_THROUGHPUT = 3000                 # auto, preferring the tile with the best FLOP rate (multi-stream serving)
DEFAULT_TILE = 0                            # what gemm_bf16(tile_n=None) passes; predict_iter(streams>1) switches it


def _i32(t):
    return t if t.dtype == torch.int32 else t.to(torch.int32)


# --------------------------------------------------------------------------- CRF
def crf_viterbi(logits, seq_len, trans, return_score=False):
    """tf.contrib.crf.crf_decode (reference tools/layer.py:140).  -> tags [B,L] int32 (+ best_score [B])."""
    require_cuda(logits, seq_len, trans)
    assert logits.dtype == torch.float32 and trans.dtype == torch.float32
    B, L, K = logits.shape
    assert trans.shape == (K, K)
    seq_len = _i32(seq_len)
    tags = torch.empty((B, L), dtype=torch.int32, device=logits.device)
    score = torch.empty((B,), dtype=torch.float32, device=logits.device) if return_score else None
    check(lib().ner_crf_viterbi(ptr(logits), ptr(seq_len), ptr(trans), ptr(tags), ptr(score), B, L, K, stream()))
    return (tags, score) if return_score else tags


def crf_loglik_fwd(logits, tags, seq_len, trans, want_alpha=False, exact=False):
    """tf.contrib.crf.crf_log_likelihood forward (reference tools/layer.py:122). -> ll [B], logz [B], alpha|None."""
    require_cuda(logits, tags, seq_len, trans)
    assert logits.dtype == torch.float32 and trans.dtype == torch.float32
    B, L, K = logits.shape
    tags, seq_len = _i32(tags), _i32(seq_len)
    ll = torch.empty((B,), dtype=torch.float32, device=logits.device)
    logz = torch.empty((B,), dtype=torch.float32, device=logits.device)
    alpha = torch.empty((B, L, K), dtype=torch.float32, device=logits.device) if want_alpha else None
    check(lib().ner_crf_loglik_fwd(ptr(logits), ptr(tags), ptr(seq_len), ptr(trans), ptr(ll), ptr(logz), ptr(alpha),
                                   B, L, K, 1 if exact else 0, stream()))
    return ll, logz, alpha


# --------------------------------------------------------------------------- dense (tcgen05)
def gemm_bf16(a, wt, bias=None, residual=None, epilogue=EPI_BF16, tile_n=None, out=None):
    """out[M,N] = epilogue(a[M,K] @ wt[N,K]^T + bias).  a, wt bf16; see ner_gemm_bf16.  tile_n None = DEFAULT_TILE."""
    if tile_n is None:
        tile_n = DEFAULT_TILE
    require_cuda(a, wt, bias, residual, out)
    assert a.dtype == torch.bfloat16 and wt.dtype == torch.bfloat16
    M, K = a.shape
    N, K2 = wt.shape
    assert K == K2
    odt = torch.float32 if epilogue in (EPI_F32, EPI_RES_F32, EPI_RES_RELU_F32) else torch.bfloat16
    if out is None:
        out = torch.empty((M, N), dtype=odt, device=a.device)
    assert out.dtype == odt and out.shape == (M, N)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N
    if residual is not None:
        assert residual.dtype == torch.float32 and residual.shape == (M, N)
    hook = _lib._HOOK
    if hook is not None:
        with hook("gemm_bf16", 2.0 * M * N * K):
            check(lib().ner_gemm_bf16(ptr(a), ptr(wt), ptr(bias), ptr(residual), ptr(out), M, N, K, epilogue, tile_n, stream()))
    else:
        check(lib().ner_gemm_bf16(ptr(a), ptr(wt), ptr(bias), ptr(residual), ptr(out), M, N, K, epilogue, tile_n, stream()))
    return out


def pack_weight_bf16(w_kn):
    """TF dense kernel [K,N] f32 -> bf16 [N,K] (B operand layout of gemm_bf16)."""
    require_cuda(w_kn)
    assert w_kn.dtype == torch.float32 and w_kn.dim() == 2
    K, N = w_kn.shape
    out = torch.empty((N, K), dtype=torch.bfloat16, device=w_kn.device)
    check(lib().ner_pack_weight_bf16(ptr(w_kn), ptr(out), K, N, stream()))
    return out


def cast_bf16(x):
    require_cuda(x)
    assert x.dtype == torch.float32
    out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    check(lib().ner_cast_bf16(ptr(x), ptr(out), x.numel(), stream()))
    return out


def dense_small_n(x, w, bias=None, row_map=None, out=None):
    """tf.layers.dense(units=label_size): x [M,F] (f32|bf16) @ w [F,N] f32 + bias -> f32 [M,N], N <= 32.
    row_map [M] i32 scatters input row r to out[row_map[r]] (packed -> padded); `out` then must be given."""
    require_cuda(x, w, bias, row_map, out)
    assert w.dtype == torch.float32 and x.dtype in (torch.float32, torch.bfloat16)
    M, F = x.shape
    F2, N = w.shape
    assert F == F2
    if out is None:
        assert row_map is None
        out = torch.empty((M, N), dtype=torch.float32, device=x.device)
    check(lib().ner_dense_small_n(ptr(x), 1 if x.dtype == torch.bfloat16 else 0, ptr(w), ptr(bias), ptr(out), M, F, N,
                                  ptr(row_map), stream()))
    return out


def seq_pack_plan(mask):
    """prefix mask [B,L] -> (cu_seqlens [B+1] i32, tok_src [B*L] i32) on the device."""
    require_cuda(mask)
    B, L = mask.shape
    mask = _i32(mask)
    cu = torch.empty((B + 1,), dtype=torch.int32, device=mask.device)
    tok_src = torch.empty((B * L,), dtype=torch.int32, device=mask.device)
    check(lib().ner_seq_pack_plan(ptr(mask), ptr(cu), ptr(tok_src), B, L, stream()))
    return cu, tok_src


# --------------------------------------------------------------------------- BERT pieces
def bert_embed_ln(word_emb, type_emb, pos_emb, gamma, beta, ids, seg, eps=1e-12, want_f32=True, want_bf16=True,
                  tok_src=None, n_packed=0):
    """Padded mode: B*L output rows.  Packed mode (tok_src, n_packed): n_packed rows."""
    require_cuda(word_emb, type_emb, pos_emb, gamma, beta, ids, seg, tok_src)
    B, L = ids.shape
    V, H = word_emb.shape
    ids = _i32(ids)
    seg = None if seg is None else _i32(seg)
    rows = n_packed if tok_src is not None else B * L
    of = torch.empty((rows, H), dtype=torch.float32, device=ids.device) if want_f32 else None
    ob = torch.empty((rows, H), dtype=torch.bfloat16, device=ids.device) if want_bf16 else None
    check(lib().ner_bert_embed_ln(ptr(word_emb), ptr(type_emb), ptr(pos_emb), ptr(gamma), ptr(beta), ptr(ids), ptr(seg),
                                  ptr(of), ptr(ob), B, L, H, V, type_emb.shape[0], pos_emb.shape[0], eps, ptr(tok_src),
                                  n_packed, stream()))
    return of, ob


def layernorm(y, gamma, beta, residual=None, eps=1e-12, want_f32=True, want_bf16=True, keep_prob=1.0, seed=0):
    """LN(dropout(y) + residual); keep_prob < 1 fuses BertModel's hidden dropout (training)."""
    require_cuda(y, gamma, beta, residual)
    assert y.dtype in (torch.float32, torch.bfloat16)
    M, H = y.shape
    of = torch.empty((M, H), dtype=torch.float32, device=y.device) if want_f32 else None
    ob = torch.empty((M, H), dtype=torch.bfloat16, device=y.device) if want_bf16 else None
    check(lib().ner_layernorm_dropout(ptr(y), 1 if y.dtype == torch.bfloat16 else 0, ptr(residual), ptr(gamma), ptr(beta),
                                      ptr(of), ptr(ob), M, H, eps, float(keep_prob), int(seed) & 0xFFFFFFFFFFFFFFFF, stream()))
    return of, ob


def bert_attention(qkv, mask, B, L, num_heads, head_dim=64, scale=None, mask_add=-10000.0, cu_seqlens=None, keep_prob=1.0,
                   seed=0):
    """Padded mode: qkv [B*L, 3HD] + mask.  Packed mode: qkv [T, 3HD] + cu_seqlens [B+1] (L = max length).
    keep_prob < 1: attention_probs dropout (training)."""
    require_cuda(qkv, mask, cu_seqlens)
    assert qkv.dtype == torch.bfloat16 and qkv.shape[1] == 3 * num_heads * head_dim
    assert cu_seqlens is not None or qkv.shape[0] == B * L
    mask = None if mask is None else _i32(mask)
    ctx = torch.empty((qkv.shape[0], num_heads * head_dim), dtype=torch.bfloat16, device=qkv.device)
    if scale is None:
        scale = 1.0 / (head_dim ** 0.5)
    check(lib().ner_bert_attention(ptr(qkv), ptr(mask), ptr(ctx), B, L, num_heads, head_dim, scale, mask_add,
                                   ptr(cu_seqlens), int(qkv.shape[0]), float(keep_prob), int(seed) & 0xFFFFFFFFFFFFFFFF, stream()))
    return ctx


# --------------------------------------------------------------------------- BiLSTM
def bilstm_recurrence(xproj, wh_fw, wh_bw, seq_len, B, L, H, activation="tanh", forget_bias=1.0, cu_seqlens=None,
                      save_for_backward=False, keep_prob=1.0, seed=0):
    """-> out [B,L,2H]; with save_for_backward also (gates [B*L,8H], cstate [B,L,2H], hstate [B,L,2H]) for
    bilstm_recurrence_bwd.  keep_prob < 1: DropoutWrapper output/state dropout (training)."""
    require_cuda(xproj, wh_fw, wh_bw, seq_len, cu_seqlens)
    assert xproj.dtype == torch.float32 and xproj.shape[1] == 8 * H
    assert cu_seqlens is not None or xproj.shape[0] == B * L
    assert wh_fw.shape == (H, 4 * H) and wh_bw.shape == (H, 4 * H)
    act = {"tanh": 0, "relu": 1}[activation]
    out = torch.empty((B, L, 2 * H), dtype=torch.float32, device=xproj.device)
    gates = cst = hst = None
    if save_for_backward:
        assert cu_seqlens is None, "training runs on the padded layout"
        gates = torch.zeros((B * L, 8 * H), dtype=torch.float32, device=xproj.device)
        cst = torch.zeros((B, L, 2 * H), dtype=torch.float32, device=xproj.device)
        hst = torch.zeros((B, L, 2 * H), dtype=torch.float32, device=xproj.device)
    check(lib().ner_bilstm_recurrence(ptr(xproj), ptr(wh_fw), ptr(wh_bw), ptr(_i32(seq_len)), ptr(out), B, L, H, act,
                                      forget_bias, ptr(cu_seqlens), ptr(gates), ptr(cst), ptr(hst), float(keep_prob),
                                      int(seed) & 0xFFFFFFFFFFFFFFFF, stream()))
    return (out, gates, cst, hst) if save_for_backward else out


def bilstm_recurrence_bwd(d_out, gates, cstate, wh_fw, wh_bw, seq_len, B, L, H, activation="tanh", keep_prob=1.0, seed=0):
    """-> d_xproj [B*L, 8H] f32 (gradient of the hoisted input projection)."""
    require_cuda(d_out, gates, cstate, wh_fw, wh_bw, seq_len)
    assert d_out.shape == (B, L, 2 * H) and d_out.dtype == torch.float32
    act = {"tanh": 0, "relu": 1}[activation]
    d_xproj = torch.empty((B * L, 8 * H), dtype=torch.float32, device=d_out.device)
    check(lib().ner_bilstm_recurrence_bwd(ptr(d_out), ptr(gates), ptr(cstate), ptr(wh_fw), ptr(wh_bw), ptr(_i32(seq_len)),
                                          ptr(d_xproj), B, L, H, act, float(keep_prob), int(seed) & 0xFFFFFFFFFFFFFFFF,
                                          stream()))
    return d_xproj


# --------------------------------------------------------------------------- SoftLexicon
def softlexicon_pool(table, ids, weights, G=4, S=10, out=None):
    """ids/weights [..., G*S] -> [..., G*E]; `out` may be a wider [n_tok, >=G*E] buffer (concat target)."""
    require_cuda(table, ids, weights, out)
    assert table.dtype == torch.float32 and weights.dtype == torch.float32
    V, E = table.shape
    lead = ids.shape[:-1]
    assert ids.shape[-1] == G * S and weights.shape == ids.shape
    n_tok = ids.numel() // (G * S)
    if out is None:
        out = torch.empty((*lead, G * E), dtype=torch.float32, device=table.device)
    ld = out.shape[-1]
    check(lib().ner_softlexicon_pool_fwd(ptr(table), ptr(_i32(ids)), ptr(weights), ptr(out), n_tok, G, S, E, V, ld,
                                         stream()))
    return out


def embedding_lookup(table, ids, out=None, col_offset=0):
    """tf.nn.embedding_lookup into out[..., col_offset:col_offset+E] (out row-major [n_tok, ld])."""
    require_cuda(table, ids, out)
    assert table.dtype == torch.float32
    V, E = table.shape
    n_tok = ids.numel()
    if out is None:
        out = torch.empty((*ids.shape, E), dtype=torch.float32, device=table.device)
    ld = out.shape[-1]
    assert col_offset + E <= ld
    check(lib().ner_embedding_lookup(ptr(table), ptr(_i32(ids)), out.data_ptr() + 4 * col_offset, n_tok, E, V, ld, stream()))
    return out


def cast_pad_bf16(x2d, Dp):
    """f32 [M,D] -> bf16 [M,Dp] zero padded."""
    require_cuda(x2d)
    assert x2d.dtype == torch.float32 and x2d.dim() == 2
    M, D = x2d.shape
    out = torch.empty((M, Dp), dtype=torch.bfloat16, device=x2d.device)
    check(lib().ner_cast_pad_bf16(ptr(x2d), ptr(out), M, D, Dp, D, stream()))
    return out


def softlexicon_pool_bwd(d_table, ids, weights, d_out, G=4, S=10):
    require_cuda(d_table, ids, weights, d_out)
    V, E = d_table.shape
    n_tok = ids.numel() // (G * S)
    check(lib().ner_softlexicon_pool_bwd(ptr(d_table), ptr(_i32(ids)), ptr(weights), ptr(d_out), n_tok, G, S, E, V,
                                         stream()))
    return d_table


def crf_loglik_bwd(logits, tags, seq_len, trans, alpha, logz, d_ll=None, scale=1.0):
    """-> d_logits [B,L,K], d_trans [K,K] for g_b = (d_ll|1) * scale."""
    require_cuda(logits, tags, seq_len, trans, alpha, logz, d_ll)
    B, L, K = logits.shape
    d_logits = torch.empty_like(logits)
    d_trans = torch.zeros_like(trans)
    check(lib().ner_crf_loglik_bwd(ptr(logits), ptr(_i32(tags)), ptr(_i32(seq_len)), ptr(trans), ptr(alpha), ptr(logz),
                                   ptr(d_ll), scale, ptr(d_logits), ptr(d_trans), B, L, K, stream()))
    return d_logits, d_trans


# --------------------------------------------------------------------------- fp32-accurate dense (split bf16)
def split_bf16(x2d, Dp=None):
    """f32 [M,D] -> (hi, lo) bf16 [M,Dp] with hi + lo ~= x to 2^-17."""
    require_cuda(x2d)
    assert x2d.dtype == torch.float32 and x2d.dim() == 2 and x2d.stride(1) == 1
    M, D = x2d.shape
    Dp = Dp or (D + 7) // 8 * 8
    hi = torch.empty((M, Dp), dtype=torch.bfloat16, device=x2d.device)
    lo = torch.empty((M, Dp), dtype=torch.bfloat16, device=x2d.device)
    check(lib().ner_split_bf16(ptr(x2d), ptr(hi), ptr(lo), M, D, Dp, x2d.stride(0), stream()))
    return hi, lo


def gemm_split_f32(a_hi, a_lo, w_hi, w_lo, bias=None, residual=None, relu=False, out=None):
    """out f32 [M,N] = [relu](A·W^T + bias [+ residual]) at ~fp32 accuracy: A_hi·W_hi + A_hi·W_lo + A_lo·W_hi
    as three tcgen05 launches chained through the f32 residual epilogue."""
    M, N = a_hi.shape[0], w_hi.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=a_hi.device)
    gemm_bf16(a_hi, w_hi, bias, residual=residual, epilogue=EPI_RES_F32 if residual is not None else EPI_F32, out=out)
    gemm_bf16(a_hi, w_lo, None, residual=out, epilogue=EPI_RES_F32, out=out)
    gemm_bf16(a_lo, w_hi, None, residual=out, epilogue=EPI_RES_RELU_F32 if relu else EPI_RES_F32, out=out)
    return out


def attention_f32(q, k, v, seq_len, B, L, num_heads, head_dim, scale=1.0, bias_u=None, bias_v=None, rel_table=None,
                  want_f32=True, want_split=False):
    """fp32 attention (+ TENER relative term).  q/k/v: f32 2-D views [B*L, >= heads*head_dim] (row stride = stride(0))."""
    require_cuda(seq_len, bias_u, bias_v, rel_table)
    for t in (q, k, v):
        assert t.is_cuda and t.dtype == torch.float32 and t.dim() == 2 and t.stride(1) == 1 and t.shape[0] == B * L
    HD = num_heads * head_dim
    of = torch.empty((B * L, HD), dtype=torch.float32, device=q.device) if want_f32 else None
    hi = torch.empty((B * L, HD), dtype=torch.bfloat16, device=q.device) if want_split else None
    lo = torch.empty((B * L, HD), dtype=torch.bfloat16, device=q.device) if want_split else None
    check(lib().ner_attention_f32(q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0),
                                  ptr(bias_u), ptr(bias_v), ptr(rel_table), ptr(_i32(seq_len)), scale, ptr(of), ptr(hi), ptr(lo),
                                  B, L, num_heads, head_dim, stream()))
    return of, hi, lo


def attention_f32_bwd(q, k, v, seq_len, B, L, num_heads, head_dim, d_out, scale=1.0, bias_u=None, bias_v=None, rel_table=None):
    """Backward of attention_f32 -> (dQ, dK, dV [B*L, heads*head_dim] f32, d_bias_u, d_bias_v [heads, head_dim] | None)."""
    require_cuda(seq_len, bias_u, bias_v, rel_table, d_out)
    for t in (q, k, v, d_out):
        assert t.is_cuda and t.dtype == torch.float32 and t.dim() == 2 and t.stride(1) == 1 and t.shape[0] == B * L
    HD = num_heads * head_dim
    dq = torch.empty((B * L, HD), dtype=torch.float32, device=q.device)
    dk = torch.zeros((B * L, HD), dtype=torch.float32, device=q.device)
    dv = torch.zeros((B * L, HD), dtype=torch.float32, device=q.device)
    du = torch.zeros((num_heads, head_dim), dtype=torch.float32, device=q.device) if bias_u is not None else None
    dvb = torch.zeros((num_heads, head_dim), dtype=torch.float32, device=q.device) if bias_v is not None else None
    check(lib().ner_attention_f32_bwd(q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0),
                                      ptr(bias_u), ptr(bias_v), ptr(rel_table), ptr(_i32(seq_len)), scale, d_out.data_ptr(),
                                      d_out.stride(0), ptr(dq), HD, ptr(dk), HD, ptr(dv), HD, ptr(du), ptr(dvb), B, L, num_heads,
                                      head_dim, stream()))
    return dq, dk, dv, du, dvb


def relu(x, inplace=False):
    require_cuda(x)
    assert x.dtype == torch.float32 and x.is_contiguous()
    y = x if inplace else torch.empty_like(x)
    check(lib().ner_relu_f32(ptr(x), ptr(y), x.numel(), stream()))
    return y


def relu_bwd(act, dact):
    require_cuda(act, dact)
    assert act.dtype == torch.float32 and dact.dtype == torch.float32 and act.numel() == dact.numel()
    out = torch.empty_like(dact)
    check(lib().ner_relu_bwd_f32(ptr(act), ptr(dact), ptr(out), act.numel(), stream()))
    return out


def reduce_max_time(x):
    """tf.reduce_max(x[B,L,C], axis=1) -> [B,C] f32."""
    require_cuda(x)
    assert x.dtype == torch.float32 and x.dim() == 3
    B, L, C = x.shape
    y = torch.empty((B, C), dtype=torch.float32, device=x.device)
    check(lib().ner_reduce_max_time(ptr(x), ptr(y), B, L, C, stream()))
    return y


def reduce_max_time_bwd(x, y, dy, dx, scale=1.0):
    """dx[b,t,c] += scale * dy[b,c] / ties where x == max (TF's reduce_max gradient)."""
    require_cuda(x, y, dy, dx)
    B, L, C = x.shape
    assert dy.dtype == torch.float32 and tuple(dy.shape) == (B, C) and dx.shape == x.shape
    check(lib().ner_reduce_max_time_bwd(ptr(x), ptr(y), ptr(dy), ptr(dx), B, L, C, float(scale), stream()))
    return dx


def softmax_xent(logits, labels, scale=1.0, want_grad=False):
    """sparse softmax cross entropy per row [B, N<=32] -> loss [B] (and scale * (softmax - onehot))."""
    require_cuda(logits, labels)
    assert logits.dtype == torch.float32 and labels.dtype == torch.int32 and logits.dim() == 2
    B, N = logits.shape
    loss = torch.empty((B,), dtype=torch.float32, device=logits.device)
    dz = torch.empty_like(logits) if want_grad else None
    check(lib().ner_softmax_xent(ptr(logits), ptr(labels), ptr(loss), ptr(dz), B, N, float(scale), stream()))
    return (loss, dz) if want_grad else loss


# --------------------------------------------------------------------------- training-side kernels
def _require_rows(x2d):
    """CUDA f32 2-D tensor whose rows are contiguous (a column slice of a wider buffer is fine)."""
    if not x2d.is_cuda:
        raise _lib.NerB200Error("ner_b200 kernels take CUDA tensors (got a CPU tensor); there is no CPU path")
    assert x2d.dtype == torch.float32 and x2d.dim() == 2 and x2d.stride(1) == 1


def transpose_cast_bf16(x2d, Mp=None):
    """f32 [M,N] -> bf16 [N,Mp]: K-major operand of a weight-gradient GEMM (reduction over the M rows)."""
    _require_rows(x2d)
    M, N = x2d.shape
    Mp = Mp or (M + 7) // 8 * 8
    out = torch.empty((N, Mp), dtype=torch.bfloat16, device=x2d.device)
    check(lib().ner_transpose_cast_bf16(ptr(x2d), ptr(out), M, N, Mp, x2d.stride(0), stream()))
    return out


def wgrad_gemm(x2d, dy2d, out=None):
    """dW [K,N] f32 = x^T [K,M] · dy [M,N] on the tensor cores (bf16 operands, fp32 accumulate)."""
    M = x2d.shape[0]
    Mp = (M + 7) // 8 * 8
    xt = transpose_cast_bf16(x2d, Mp)        # [K, Mp]
    dyt = transpose_cast_bf16(dy2d, Mp)      # [N, Mp]
    N = dyt.shape[0]
    if N % 32 != 0:                          # pad the output columns to the GEMM's 32-column granule
        Np = (N + 31) // 32 * 32
        dyt = torch.nn.functional.pad(dyt, (0, 0, 0, Np - N))
    res = gemm_bf16(xt, dyt.contiguous(), None, epilogue=EPI_F32)
    res = res[:, :N] if res.shape[1] != N else res
    if out is not None:
        out.add_(res)
        return out
    return res.contiguous()


def colsum_add(x2d, out, scale=1.0):
    _require_rows(x2d)
    require_cuda(out)
    M, N = x2d.shape
    check(lib().ner_colsum_add(ptr(x2d), ptr(out), M, N, x2d.stride(0), scale, stream()))
    return out


def dense_small_n_bwd(x2d, w, dy, dW, db=None, want_dx=True):
    require_cuda(x2d, w, dy, dW, db)
    assert x2d.dtype == torch.float32
    M, F = x2d.shape
    N = w.shape[1]
    dx = torch.empty((M, F), dtype=torch.float32, device=x2d.device) if want_dx else None
    check(lib().ner_dense_small_n_bwd(ptr(x2d), ptr(w), ptr(dy), ptr(dW), ptr(db), ptr(dx), M, F, N, stream()))
    return dx


def dropout(x, keep_prob, seed, inplace=False):
    """tf.layers.dropout forward (and backward: same call on the gradient with the same seed); f32 or bf16."""
    require_cuda(x)
    assert x.dtype in (torch.float32, torch.bfloat16) and x.is_contiguous()
    y = x if inplace else torch.empty_like(x)
    fn = lib().ner_dropout if x.dtype == torch.float32 else lib().ner_dropout_bf16
    check(fn(ptr(x), ptr(y), x.numel(), float(keep_prob), int(seed) & 0xFFFFFFFFFFFFFFFF, stream()))
    return y


_sumsq_scratch = {}


def sumsq_add(g, out):
    require_cuda(g, out)
    key = (g.device.index, stream())
    scratch = _sumsq_scratch.get(key)
    if scratch is None:
        scratch = _sumsq_scratch[key] = torch.empty(int(lib().ner_sumsq_scratch_floats()), dtype=torch.float32, device=g.device)
    check(lib().ner_sumsq_add(ptr(g), g.numel(), ptr(out), ptr(scratch), stream()))
    return out


def adam_step(p, g, m, v, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, mode=1, clip=0.0, gnorm_sq=None,
              grad_scale=1.0):
    """mode 0: AdamWeightDecayOptimizer (bert), mode 1: tf.train.AdamOptimizer; flat f32 buffers."""
    require_cuda(p, g, m, v, gnorm_sq)
    n = p.numel()
    assert g.numel() == n and m.numel() == n and v.numel() == n
    check(lib().ner_adam_step(ptr(p), ptr(g), ptr(m), ptr(v), n, lr, beta1, beta2, eps, weight_decay, mode, clip,
                              ptr(gnorm_sq), grad_scale, stream()))


# --------------------------------------------------------------------------- encoder backward
def layernorm_bwd(y, gamma, d_out, d_gamma, d_beta, residual=None, eps=1e-12, want_f32=True, want_bf16=True, keep_prob=1.0,
                  seed=0, d_bias=None):
    """d_bias (optional, [H] f32, accumulated into): column sums of the masked dense-branch gradient = the bias gradient of
    the dense layer whose output this LayerNorm normalises."""
    require_cuda(y, gamma, d_out, d_gamma, d_beta, residual, d_bias)
    M, H = y.shape
    dz32 = torch.empty((M, H), dtype=torch.float32, device=y.device) if want_f32 else None
    dz16 = torch.empty((M, H), dtype=torch.bfloat16, device=y.device) if want_bf16 else None
    check(lib().ner_layernorm_dropout_bwd_bias(ptr(y), 1 if y.dtype == torch.bfloat16 else 0, ptr(residual), ptr(gamma),
                                               ptr(d_out), ptr(dz32), ptr(dz16), ptr(d_gamma), ptr(d_beta), ptr(d_bias), M, H, eps,
                                               float(keep_prob), int(seed) & 0xFFFFFFFFFFFFFFFF, stream()))
    return dz32, dz16


def transpose_bf16(x, Mp=None):
    require_cuda(x)
    assert x.dtype == torch.bfloat16 and x.dim() == 2
    M, N = x.shape
    Mp = Mp or (M + 7) // 8 * 8
    out = torch.empty((N, Mp), dtype=torch.bfloat16, device=x.device)
    check(lib().ner_transpose_bf16(ptr(x), ptr(out), M, N, Mp, stream()))
    return out


def wgrad_gemm_bf16(x16, dy16, out_f32=None):
    """dW [K,N] f32 (+)= x^T · dy for bf16 activations x [M,K], dy [M,N] (tensor cores, fp32 accumulate)."""
    M = x16.shape[0]
    Mp = (M + 7) // 8 * 8
    xt, dyt = transpose_bf16(x16, Mp), transpose_bf16(dy16, Mp)
    if out_f32 is None:
        return gemm_bf16(xt, dyt, None, epilogue=EPI_F32)
    return gemm_bf16(xt, dyt, None, residual=out_f32, epilogue=EPI_RES_F32, out=out_f32)


def colsum_bf16_add(x16, out):
    require_cuda(x16, out)
    M, N = x16.shape
    check(lib().ner_colsum_bf16_add(ptr(x16), ptr(out), M, N, stream()))
    return out


def gelu_bf16(pre, erf=False):
    require_cuda(pre)
    out = torch.empty_like(pre)
    check(lib().ner_gelu_bf16(ptr(pre), ptr(out), pre.numel(), 1 if erf else 0, stream()))
    return out


def gelu_f32(x, erf=False, inplace=False):
    require_cuda(x)
    assert x.dtype == torch.float32 and x.is_contiguous()
    y = x if inplace else torch.empty_like(x)
    check(lib().ner_gelu_f32(ptr(x), ptr(y), x.numel(), 1 if erf else 0, stream()))
    return y


def gelu_bwd_bias_bf16(pre, dact, d_bias, erf=False):
    """d_pre = d_act * gelu'(pre) and d_bias += column sums of d_pre, one pass (pre / dact bf16 [M, N])."""
    require_cuda(pre, dact, d_bias)
    M, N = pre.shape
    out = torch.empty_like(pre)
    check(lib().ner_gelu_bwd_bias_bf16(ptr(pre), ptr(dact), ptr(out), ptr(d_bias), M, N, 1 if erf else 0, stream()))
    return out


def gelu_bwd_bf16(pre, dact, erf=False):
    require_cuda(pre, dact)
    out = torch.empty_like(pre)
    check(lib().ner_gelu_bwd_bf16(ptr(pre), ptr(dact), ptr(out), pre.numel(), 1 if erf else 0, stream()))
    return out


def bert_embed_bwd(dx, ids, seg, d_word, d_type, d_pos):
    require_cuda(dx, ids, seg, d_word, d_type, d_pos)
    B, L = ids.shape
    H = dx.shape[-1]
    check(lib().ner_bert_embed_bwd(ptr(dx), ptr(_i32(ids)), ptr(None if seg is None else _i32(seg)), ptr(d_word), ptr(d_type),
                                   ptr(d_pos), B, L, H, d_word.shape[0], d_type.shape[0], stream()))


def gather_rows(src2d, idx, n):
    """dst row r = src row idx[r] for r < n (padded -> packed layout)."""
    require_cuda(src2d, idx)
    assert idx.dtype == torch.int32 and src2d.dim() == 2
    dst = torch.empty((n, src2d.shape[1]), dtype=src2d.dtype, device=src2d.device)
    check(lib().ner_gather_rows(ptr(src2d), ptr(idx), ptr(dst), n, src2d.shape[1] * src2d.element_size(), stream()))
    return dst


def scatter_rows(src2d, idx, rows):
    """dst [rows, C] zeros with dst row idx[r] = src row r (packed -> padded layout)."""
    require_cuda(src2d, idx)
    assert idx.dtype == torch.int32 and src2d.dim() == 2
    dst = torch.zeros((rows, src2d.shape[1]), dtype=src2d.dtype, device=src2d.device)
    check(lib().ner_scatter_rows(ptr(src2d), ptr(idx), ptr(dst), src2d.shape[0], src2d.shape[1] * src2d.element_size(), stream()))
    return dst


def bert_attention_bwd(qkv, mask, ctx, dctx, B, L, num_heads, head_dim=64, scale=None, mask_add=-10000.0, keep_prob=1.0, seed=0,
                       cu_seqlens=None):
    require_cuda(qkv, mask, ctx, dctx, cu_seqlens)
    assert qkv.dtype == torch.bfloat16 and ctx.dtype == torch.bfloat16 and dctx.dtype == torch.bfloat16
    dqkv = torch.empty_like(qkv)
    if scale is None:
        scale = 1.0 / (head_dim ** 0.5)
    if cu_seqlens is not None:
        check(lib().ner_bert_attention_bwd_packed(ptr(qkv), ptr(cu_seqlens), ptr(ctx), ptr(dctx), ptr(dqkv), B, L, num_heads, head_dim,
                                                  scale, float(keep_prob), int(seed) & 0xFFFFFFFFFFFFFFFF, stream()))
        return dqkv
    check(lib().ner_bert_attention_bwd(ptr(qkv), ptr(_i32(mask)), ptr(ctx), ptr(dctx), ptr(dqkv), B, L, num_heads, head_dim,
                                       scale, mask_add, float(keep_prob), int(seed) & 0xFFFFFFFFFFFFFFFF, stream()))
    return dqkv


# --------------------------------------------------------------------------- entity spans (serving tail)
def tag_classes(idx2tag):
    """idx2tag -> (uint8 class table [K] for ner_extract_spans, list of entity types)."""
    K = max(idx2tag) + 1
    types, table = [], [0] * K
    for i, tag in idx2tag.items():
        head = tag.split('-')[0]
        kind = 1 if head == 'B' else 2 if head == 'I' else 0
        t = 0
        if kind and '-' in tag:
            name = tag.split('-')[1]
            if name not in types:
                types.append(name)
            t = types.index(name)
        table[i] = kind | (4 if tag[:1] in ('B', 'I') else 0) | (t << 3)
    assert len(types) <= 32
    return torch.tensor(table, dtype=torch.uint8), types


def extract_spans(pred_ids, tag_class, cap=None):
    """pred_ids [B,L] int32 (device), tag_class uint8 [K] (device) -> (spans int32 [B,cap], counts int32 [B])."""
    require_cuda(pred_ids, tag_class)
    B, L = pred_ids.shape
    cap = cap or L            # 'B B B ...': every position can be a span of its own
    spans = torch.empty((B, cap), dtype=torch.int32, device=pred_ids.device)
    counts = torch.empty((B,), dtype=torch.int32, device=pred_ids.device)
    check(lib().ner_extract_spans(ptr(_i32(pred_ids)), ptr(tag_class), ptr(spans), ptr(counts), B, L, tag_class.numel(), cap, stream()))
    return spans, counts


def wgrad_group(problems, rows):
    """problems: list of (x bf16 [rows, ld_x], dy bf16 [rows, ld_dy], dy_col0, dw f32 [k_in, n_out]) -> dw += x[:, :k_in]^T dy[:, col0:col0+n_out],
    all in one launch (ner_wgrad_group_bf16)."""
    from ._lib import WgradProblem
    arr = (WgradProblem * len(problems))()
    for q, (x, dy, col0, dw) in zip(arr, problems):
        require_cuda(x, dy, dw)
        assert x.dtype == torch.bfloat16 and dy.dtype == torch.bfloat16 and dw.dtype == torch.float32
        q.x_bf16, q.ld_x, q.dy_bf16, q.ld_dy, q.dy_col0 = ptr(x), x.shape[1], ptr(dy), dy.shape[1], int(col0)
        q.dw, q.k_in, q.n_out = ptr(dw), dw.shape[0], dw.shape[1]
    check(lib().ner_wgrad_group_bf16(arr, len(problems), int(rows), stream()))


class PackGroup(object):
    """A fixed set of (f32 kernel [K,N], bf16 [N,K] pack view, bf16 [K,N] cast view) triples re-packed by ONE launch
    (ner_pack_weights_group_bf16).  The table lives on the device; `run()` after every optimizer step."""

    def __init__(self, triples):
        import ctypes
        from ._lib import PackEntry
        arr = (PackEntry * len(triples))()
        starts = [0]
        self.keep = triples
        for q, (src, nk, kn) in zip(arr, triples):
            require_cuda(src)
            for t in (nk, kn):          # destinations may be row / column blocks of a fused operand: unit stride along the row only
                if t is not None and not (t.is_cuda and t.dtype == torch.bfloat16 and t.stride(1) == 1):
                    raise NerB200Error("PackGroup destinations must be CUDA bf16 matrices with contiguous rows")
            K, N = src.shape
            assert src.dtype == torch.float32 and src.is_contiguous()
            q.src, q.K, q.N = ptr(src), K, N
            q.dst_nk_bf16, q.ld_nk = (ptr(nk), nk.stride(0)) if nk is not None else (0, 0)
            q.dst_kn_bf16, q.ld_kn = (ptr(kn), kn.stride(0)) if kn is not None else (0, 0)
            starts.append(starts[-1] + ((K + 63) // 64) * ((N + 63) // 64))
        dev = triples[0][0].device
        raw = bytes(memoryview(arr))
        self.table = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev)
        self.starts = torch.tensor(starts, dtype=torch.int32, device=dev)
        self.count, self.tiles = len(triples), starts[-1]
        self.signature = tuple(ptr(t[0]) for t in triples)

    def run(self):
        check(lib().ner_pack_weights_group_bf16(ptr(self.table), ptr(self.starts), self.count, self.tiles, stream()))
