"""BertModel forward on the sm_100a kernels (bert_base.bert.modeling.BertModel as driven by
reference tools/layer.py:63-81; variable names per SURVEY.md §8 a9 / Appendix A.3).

Per layer: fused-QKV tcgen05 GEMM -> attention kernel -> tcgen05 GEMM (+bias +residual, fp32)
-> LayerNorm (fp32 + bf16 copies) -> tcgen05 GEMM (+bias, GELU) -> tcgen05 GEMM (+bias
+residual) -> LayerNorm.  The residual stream stays fp32; GEMM operands are bf16.
"""
import json
import os

import torch

from . import ops, variables
from .config import BERT_BASE_CHINESE


def load_bert_config(pretrain_dir):
    cfg = dict(BERT_BASE_CHINESE)
    path = os.path.join(pretrain_dir or "", "bert_config.json")
    if pretrain_dir and os.path.exists(path):
        with open(path) as f:
            cfg.update(json.load(f))
    return cfg


def create_bert_variables(cfg, store, scope="bert"):
    """Create (or fetch) every BertModel variable with google-research/bert's initializers."""
    H, I = cfg["hidden_size"], cfg["intermediate_size"]
    tn = variables.truncated_normal(cfg.get("initializer_range", 0.02))
    gv = store.get_variable
    gv(f"{scope}/embeddings/word_embeddings", (cfg["vocab_size"], H), tn)
    gv(f"{scope}/embeddings/token_type_embeddings", (cfg["type_vocab_size"], H), tn)
    gv(f"{scope}/embeddings/position_embeddings", (cfg["max_position_embeddings"], H), tn)
    gv(f"{scope}/embeddings/LayerNorm/beta", (H,), variables.zeros)
    gv(f"{scope}/embeddings/LayerNorm/gamma", (H,), variables.ones)
    for l in range(cfg["num_hidden_layers"]):
        p = f"{scope}/encoder/layer_{l}"
        for n in ("query", "key", "value"):
            gv(f"{p}/attention/self/{n}/kernel", (H, H), tn)
            gv(f"{p}/attention/self/{n}/bias", (H,), variables.zeros)
        gv(f"{p}/attention/output/dense/kernel", (H, H), tn)
        gv(f"{p}/attention/output/dense/bias", (H,), variables.zeros)
        gv(f"{p}/attention/output/LayerNorm/beta", (H,), variables.zeros)
        gv(f"{p}/attention/output/LayerNorm/gamma", (H,), variables.ones)
        gv(f"{p}/intermediate/dense/kernel", (H, I), tn)
        gv(f"{p}/intermediate/dense/bias", (I,), variables.zeros)
        gv(f"{p}/output/dense/kernel", (I, H), tn)
        gv(f"{p}/output/dense/bias", (H,), variables.zeros)
        gv(f"{p}/output/LayerNorm/beta", (H,), variables.zeros)
        gv(f"{p}/output/LayerNorm/gamma", (H,), variables.ones)
    # pooler exists in the checkpoint but is unused by the reference (only sequence_output)
    gv(f"{scope}/pooler/dense/kernel", (H, H), tn)
    gv(f"{scope}/pooler/dense/bias", (H,), variables.zeros)


def _packed(store, cfg, scope):
    """bf16 [N,K] packs of every dense kernel (+ fused QKV), rebuilt when the store changes."""
    def build():
        v = store.vars
        out = []
        for l in range(cfg["num_hidden_layers"]):
            p = f"{scope}/encoder/layer_{l}"
            wqkv = torch.cat([v[f"{p}/attention/self/{n}/kernel"] for n in ("query", "key", "value")], dim=1).contiguous()
            bqkv = torch.cat([v[f"{p}/attention/self/{n}/bias"] for n in ("query", "key", "value")]).contiguous()
            out.append(dict(
                wqkv=ops.pack_weight_bf16(wqkv), bqkv=bqkv,
                wo=ops.pack_weight_bf16(v[f"{p}/attention/output/dense/kernel"]), bo=v[f"{p}/attention/output/dense/bias"],
                g1=v[f"{p}/attention/output/LayerNorm/gamma"], b1=v[f"{p}/attention/output/LayerNorm/beta"],
                wi=ops.pack_weight_bf16(v[f"{p}/intermediate/dense/kernel"]), bi=v[f"{p}/intermediate/dense/bias"],
                wd=ops.pack_weight_bf16(v[f"{p}/output/dense/kernel"]), bd=v[f"{p}/output/dense/bias"],
                g2=v[f"{p}/output/LayerNorm/gamma"], b2=v[f"{p}/output/LayerNorm/beta"]))
        return out
    return store.cached(("bert_pack", scope), build)


class PackInfo:
    """Sequence-packing plan of one batch: token rows of all sequences back to back, no padding."""

    def __init__(self, cu_seqlens, tok_src, total, B, L):
        self.cu_seqlens, self.tok_src, self.total, self.B, self.L = cu_seqlens, tok_src, int(total), B, L


def make_pack(input_mask, total_tokens=None):
    """Plan from a prefix mask [B,L].  `total_tokens` (host int) avoids a device sync."""
    B, L = input_mask.shape
    if total_tokens is None:
        total_tokens = getattr(input_mask, "total_tokens", None)
    if total_tokens is None:
        total_tokens = int(input_mask.sum().item())      # device sync; engine.Estimator passes the host count
    cu, tok_src = ops.seq_pack_plan(input_mask)
    return PackInfo(cu, tok_src, total_tokens, B, L)


def bert_forward(input_ids, input_mask, segment_ids, cfg, store=None, scope="bert", gelu="tanh", pack=None):
    """-> (sequence_output f32 [rows,H], bf16 copy [rows,H]); rows = B*L, or pack.total in packed mode."""
    store = store or variables.default_store()
    create_bert_variables(cfg, store, scope)
    B, L = input_ids.shape
    H, NH = cfg["hidden_size"], cfg["num_attention_heads"]
    v = store.vars
    layers = _packed(store, cfg, scope)
    x32, x16 = ops.bert_embed_ln(v[f"{scope}/embeddings/word_embeddings"], v[f"{scope}/embeddings/token_type_embeddings"],
                                 v[f"{scope}/embeddings/position_embeddings"], v[f"{scope}/embeddings/LayerNorm/gamma"],
                                 v[f"{scope}/embeddings/LayerNorm/beta"], input_ids, segment_ids, eps=1e-12,
                                 tok_src=pack.tok_src if pack else None, n_packed=pack.total if pack else 0)
    epi_gelu = ops.EPI_GELU_ERF_BF16 if gelu == "erf" else ops.EPI_GELU_TANH_BF16
    cu = pack.cu_seqlens if pack else None
    for w in layers:
        qkv = ops.gemm_bf16(x16, w["wqkv"], w["bqkv"], epilogue=ops.EPI_BF16)
        ctx = ops.bert_attention(qkv, input_mask, B, L, NH, H // NH, cu_seqlens=cu)
        y = ops.gemm_bf16(ctx, w["wo"], w["bo"], residual=x32, epilogue=ops.EPI_RES_F32)
        x32, x16 = ops.layernorm(y, w["g1"], w["b1"], eps=1e-12)
        inter = ops.gemm_bf16(x16, w["wi"], w["bi"], epilogue=epi_gelu)
        y = ops.gemm_bf16(inter, w["wd"], w["bd"], residual=x32, epilogue=ops.EPI_RES_F32)
        x32, x16 = ops.layernorm(y, w["g2"], w["b2"], eps=1e-12)
    return x32, x16
