"""ctypes binding of libner_b200.so (the C-ABI in include/ner_b200.h).

The library is the product path: if it is missing or cannot be loaded this module raises —
there is no CPU or PyTorch fallback.  Tensors are passed as raw device pointers; torch is only
the allocator / stream container.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libner_b200.so")

_c = ctypes
_vp, _i = _c.c_void_p, _c.c_int

# name -> (restype, argtypes); must list every symbol include/ner_b200.h declares.
SIGNATURES = {
    "ner_strerror": (_c.c_char_p, [_i]),
    "ner_abi_version": (_i, []),
    "ner_build_info": (_c.c_char_p, []),
    "ner_crf_viterbi": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "ner_crf_loglik_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "ner_crf_loglik_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _c.c_float, _vp, _vp, _i, _i, _i, _vp]),
    "ner_gemm_bf16": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "ner_pack_weight_bf16": (_i, [_vp, _vp, _i, _i, _vp]),
    "ner_cast_bf16": (_i, [_vp, _vp, _c.c_size_t, _vp]),
    "ner_dense_small_n": (_i, [_vp, _i, _vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "ner_seq_pack_plan": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "ner_bert_embed_ln": (_i, [_vp] * 9 + [_i] * 6 + [_c.c_float, _vp, _i, _vp]),
    "ner_layernorm": (_i, [_vp, _i] + [_vp] * 5 + [_i, _i, _c.c_float, _vp]),
    "ner_layernorm_dropout": (_i, [_vp, _i] + [_vp] * 5 + [_i, _i, _c.c_float, _c.c_float, _c.c_uint64, _vp]),
    "ner_bert_attention": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _c.c_float, _c.c_float, _vp, _i, _c.c_float, _c.c_uint64, _vp]),
    "ner_bilstm_recurrence": (_i, [_vp] * 5 + [_i, _i, _i, _i, _c.c_float, _vp, _vp, _vp, _vp, _c.c_float, _c.c_uint64, _vp]),
    "ner_bilstm_recurrence_bwd": (_i, [_vp] * 7 + [_i, _i, _i, _i, _c.c_float, _c.c_uint64, _vp]),
    "ner_transpose_cast_bf16": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "ner_colsum_add": (_i, [_vp, _vp, _i, _i, _i, _c.c_float, _vp]),
    "ner_dense_small_n_bwd": (_i, [_vp] * 6 + [_i, _i, _i, _vp]),
    "ner_dropout": (_i, [_vp, _vp, _c.c_size_t, _c.c_float, _c.c_uint64, _vp]),
    "ner_relu_bwd_f32": (_i, [_vp, _vp, _vp, _c.c_size_t, _vp]),
    "ner_relu_f32": (_i, [_vp, _vp, _c.c_size_t, _vp]),
    "ner_reduce_max_time": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "ner_reduce_max_time_bwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _c.c_float, _vp]),
    "ner_softmax_xent": (_i, [_vp, _vp, _vp, _vp, _i, _i, _c.c_float, _vp]),
    "ner_attention_f32_bwd": (_i, [_vp, _i, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _c.c_float, _vp, _i, _vp, _i, _vp, _i, _vp, _i,
                                   _vp, _vp, _i, _i, _i, _i, _vp]),
    "ner_dropout_bf16": (_i, [_vp, _vp, _c.c_size_t, _c.c_float, _c.c_uint64, _vp]),
    "ner_sumsq_add": (_i, [_vp, _c.c_size_t, _vp, _vp, _vp]),
    "ner_sumsq_scratch_floats": (_c.c_size_t, []),
    "ner_layernorm_bwd": (_i, [_vp, _i] + [_vp] * 7 + [_i, _i, _c.c_float, _vp]),
    "ner_layernorm_dropout_bwd": (_i, [_vp, _i] + [_vp] * 7 + [_i, _i, _c.c_float, _c.c_float, _c.c_uint64, _vp]),
    "ner_layernorm_dropout_bwd_bias": (_i, [_vp, _i] + [_vp] * 8 + [_i, _i, _c.c_float, _c.c_float, _c.c_uint64, _vp]),
    "ner_transpose_bf16": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "ner_colsum_bf16_add": (_i, [_vp, _vp, _i, _i, _vp]),
    "ner_gelu_bf16": (_i, [_vp, _vp, _c.c_size_t, _i, _vp]),
    "ner_gelu_f32": (_i, [_vp, _vp, _c.c_size_t, _i, _vp]),
    "ner_gelu_bwd_bf16": (_i, [_vp, _vp, _vp, _c.c_size_t, _i, _vp]),
    "ner_gelu_bwd_bias_bf16": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "ner_bert_embed_bwd": (_i, [_vp] * 6 + [_i] * 5 + [_vp]),
    "ner_bert_attention_bwd": (_i, [_vp] * 5 + [_i] * 4 + [_c.c_float, _c.c_float, _c.c_float, _c.c_uint64, _vp]),
    "ner_bert_attention_bwd_packed": (_i, [_vp] * 5 + [_i] * 4 + [_c.c_float, _c.c_float, _c.c_uint64, _vp]),
    "ner_gather_rows": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "ner_scatter_rows": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "ner_adam_step": (_i, [_vp] * 4 + [_c.c_size_t] + [_c.c_float] * 5 + [_i, _c.c_float, _vp, _c.c_float, _vp]),
    "ner_softlexicon_pool_fwd": (_i, [_vp] * 4 + [_i] * 6 + [_vp]),
    "ner_embedding_lookup": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "ner_cast_pad_bf16": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "ner_split_bf16": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "ner_attention_f32": (_i, [_vp, _i, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _c.c_float, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "ner_softlexicon_pool_bwd": (_i, [_vp] * 4 + [_i] * 5 + [_vp]),
}



class BertConfig(_c.Structure):
    _fields_ = [("hidden_size", _i), ("num_heads", _i), ("intermediate_size", _i), ("num_layers", _i),
                ("vocab_size", _i), ("type_vocab_size", _i), ("max_position", _i), ("ln_eps", _c.c_float),
                ("gelu_erf", _i), ("gemm_tile", _i)]


class BertLayerWeights(_c.Structure):
    _fields_ = [(n, _vp) for n in ("wqkv", "bqkv", "wo", "bo", "ln1_gamma", "ln1_beta", "wi", "bi", "wd", "bd",
                                   "ln2_gamma", "ln2_beta")]


class BertLayerGrads(_c.Structure):
    _fields_ = [(n, _vp) for n in ("wqkv_kn", "wo_kn", "wi_kn", "wd_kn", "d_wq", "d_wk", "d_wv", "d_bq", "d_bk", "d_bv",
                                   "d_wo", "d_bo", "d_ln1_gamma", "d_ln1_beta", "d_wi", "d_bi", "d_wd", "d_bd",
                                   "d_ln2_gamma", "d_ln2_beta")]


SIGNATURES["ner_bert_train_saved_bytes"] = (_c.c_size_t, [_c.POINTER(BertConfig), _i])
SIGNATURES["ner_bert_train_scratch_bytes"] = (_c.c_size_t, [_c.POINTER(BertConfig), _i])
SIGNATURES["ner_bert_encoder_train_fwd"] = (_i, [_c.POINTER(BertConfig)] + [_vp] * 5 + [_c.POINTER(BertLayerWeights)] + [_vp] * 3
                                            + [_i, _i, _c.c_float, _c.c_float, _c.c_uint64, _vp, _vp, _vp, _c.c_size_t, _vp])
SIGNATURES["ner_bert_encoder_train_bwd"] = (_i, [_c.POINTER(BertConfig), _vp, _c.POINTER(BertLayerWeights), _c.POINTER(BertLayerGrads)]
                                            + [_vp] * 5 + [_vp] * 3 + [_i, _i, _c.c_float, _c.c_float, _c.c_uint64, _vp, _vp,
                                                                       _c.c_size_t, _vp, _c.c_size_t, _vp])
SIGNATURES["ner_bert_train_packed_saved_bytes"] = (_c.c_size_t, [_c.POINTER(BertConfig), _i])
SIGNATURES["ner_bert_train_packed_scratch_bytes"] = (_c.c_size_t, [_c.POINTER(BertConfig), _i, _i])
SIGNATURES["ner_bert_encoder_train_fwd_packed"] = (_i, [_c.POINTER(BertConfig)] + [_vp] * 5 + [_c.POINTER(BertLayerWeights)] + [_vp] * 2
                                                   + [_i, _i, _vp, _vp, _i, _c.c_float, _c.c_float, _c.c_uint64, _vp, _vp, _vp,
                                                      _c.c_size_t, _vp])
SIGNATURES["ner_bert_encoder_train_bwd_packed"] = (_i, [_c.POINTER(BertConfig), _vp, _c.POINTER(BertLayerWeights),
                                                        _c.POINTER(BertLayerGrads)] + [_vp] * 5 + [_vp] * 2
                                                   + [_i, _i, _vp, _vp, _i, _c.c_float, _c.c_float, _c.c_uint64, _vp, _vp,
                                                      _c.c_size_t, _vp, _c.c_size_t, _vp])
SIGNATURES["ner_bert_bilstm_crf_predict_workspace_bytes"] = (_c.c_size_t, [_c.POINTER(BertConfig), _i, _i, _i, _i, _i])
SIGNATURES["ner_bert_bilstm_crf_predict"] = (_i, [_c.POINTER(BertConfig)] + [_vp] * 5 + [_c.POINTER(BertLayerWeights)] + [_vp] * 4
                                             + [_i, _i] + [_vp] * 3 + [_i] + [_vp] * 4 + [_i, _i, _i, _vp, _vp, _c.c_size_t, _vp])
SIGNATURES["ner_bert_embed_sum"] = (_i, [_vp] * 6 + [_i] * 6 + [_vp])
SIGNATURES["ner_axpy_f32"] = (_i, [_vp, _vp, _c.c_size_t, _c.c_float, _vp])
SIGNATURES["ner_bert_encoder_workspace_bytes"] = (_c.c_size_t, [_c.POINTER(BertConfig), _i])
SIGNATURES["ner_bert_encoder_fwd"] = (_i, [_c.POINTER(BertConfig)] + [_vp] * 5 + [_c.POINTER(BertLayerWeights)] + [_vp] * 3
                                      + [_i, _i, _vp, _vp, _i, _vp, _vp, _vp, _c.c_size_t, _vp])

class PackEntry(_c.Structure):
    _fields_ = [("src", _vp), ("K", _i), ("N", _i), ("dst_nk_bf16", _vp), ("ld_nk", _i), ("dst_kn_bf16", _vp), ("ld_kn", _i)]


SIGNATURES["ner_pack_weights_group_bf16"] = (_i, [_vp, _vp, _i, _i, _vp])


class WgradProblem(_c.Structure):
    _fields_ = [("x_bf16", _vp), ("ld_x", _i), ("dy_bf16", _vp), ("ld_dy", _i), ("dy_col0", _i), ("dw", _vp), ("k_in", _i),
                ("n_out", _i)]


SIGNATURES["ner_wgrad_group_bf16"] = (_i, [_c.POINTER(WgradProblem), _i, _i, _vp])
SIGNATURES["ner_bert_train_bwd_set_layer_events"] = (_i, [_vp, _i])
SIGNATURES["ner_extract_spans"] = (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp])
SIGNATURES["ner_lexicon_create"] = (_vp, [_vp, _vp, _vp, _i])
SIGNATURES["ner_lexicon_destroy"] = (None, [_vp])
SIGNATURES["ner_lexicon_num_nodes"] = (_c.c_int64, [_vp])
SIGNATURES["ner_lexicon_build"] = (_i, [_vp, _vp, _vp, _i, _vp, _vp, _i, _i, _vp, _vp, _i])

_lib = None


class NerB200Error(RuntimeError):
    pass


def lib():
    """Load (once) and return the ctypes handle; raises if the extension is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NerB200Error(
                f"{LIB_PATH} not found: build it with `python -m chinesener_b200.build` "
                "(there is no CPU fallback for the sm_100a kernels)")
        h = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(h, name)
            fn.restype = res
            fn.argtypes = args
        _lib = h
    return _lib


LAUNCHES = 0     # kernel launches issued through the C-ABI (every successful call launches one kernel)
_HOOK = None     # optional profiling hook: fn(name) -> context manager, set by bench.py


def check(status):
    global LAUNCHES
    LAUNCHES += 1
    if status != 0:
        raise NerB200Error(lib().ner_strerror(status).decode())


def ptr(t):
    return 0 if t is None else t.data_ptr()


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream():
    """cudaStream_t of torch's current stream.  The raw getter avoids building a torch.cuda.Stream object per
    call: with a non-default stream current (predict_iter(streams>1)) that wrapper cost ~0.4 ms per call on the
    GPU box's host profile — more than the whole step's device time."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise NerB200Error("ner_b200 kernels take CUDA tensors (got a CPU tensor); there is no CPU path")
        if t is not None and not t.is_contiguous():
            raise NerB200Error("ner_b200 kernels take contiguous row-major tensors")
