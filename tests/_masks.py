"""numpy restatement of the library's counter-based dropout decisions (chinesener_b200/csrc/common.cuh:
hash3 / keep_threshold), so tests can rebuild the exact masks a kernel used from (seed, coordinates)."""
import numpy as np

_M = np.uint64(0xFFFFFFFF)


def hash3(a, b, c):
    a, b, c = (np.asarray(v, dtype=np.uint64) & _M for v in (a, b, c))
    x = ((a * np.uint64(0x9E3779B1)) & _M) ^ (((b + np.uint64(0x7F4A7C15)) * np.uint64(0x85EBCA77)) & _M) \
        ^ (((c + np.uint64(0x165667B1)) * np.uint64(0xC2B2AE3D)) & _M)
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x7FEB352D)) & _M
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x846CA68B)) & _M
    x ^= x >> np.uint64(16)
    return x


def keep_threshold(keep):
    return np.uint64(min(float(np.float32(keep) * np.float32(4294967296.0)), 4294967295.0))


def elementwise_keep(n, keep, seed):
    """ner_dropout / ner_dropout_bf16: element i of a flat tensor."""
    i = np.arange(n, dtype=np.uint64)
    lo, hi = np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF)
    return hash3(lo, hi ^ (i >> np.uint64(32)), i & _M) < keep_threshold(keep)


def attention_keep(B, NH, L, keep, seed):
    """ner_bert_attention(_bwd): [B, NH, L(q), L(k)] boolean keep mask."""
    lo, hi = np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF)
    bh = np.arange(B * NH, dtype=np.uint64).reshape(B, NH, 1, 1)
    q = np.arange(L, dtype=np.uint64).reshape(1, 1, L, 1)
    k = np.arange(L, dtype=np.uint64).reshape(1, 1, 1, L)
    sa = lo ^ ((bh * np.uint64(0x9E3779B1)) & _M)
    return hash3(sa, q, k ^ hi) < keep_threshold(keep)
