"""The C-ABI library loads on a CPU-only box and exports every symbol include/ner_b200.h declares."""
import ctypes
import os
import re

from chinesener_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "ner_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ner_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "build with python -m chinesener_b200.build"
    h = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 5
    for name in declared:
        assert hasattr(h, name), f"{name} declared in ner_b200.h but not exported"


def test_binding_table_matches_header():
    assert sorted(_lib.SIGNATURES) == _declared_symbols()


def test_strerror_and_version():
    h = _lib.lib()
    assert h.ner_abi_version() >= 1
    assert h.ner_strerror(0) == b"ok"
    assert b"invalid" in h.ner_strerror(-1)
    assert b"CUDA error" in h.ner_strerror(-1001)


def test_invalid_args_rejected_without_gpu():
    h = _lib.lib()
    # null pointers / bad sizes are rejected before any CUDA call
    assert h.ner_crf_viterbi(None, None, None, None, None, 4, 8, 10, None) == -1
    assert h.ner_crf_viterbi(None, None, None, None, None, 0, 8, 10, None) == 0  # empty batch is a no-op
    assert h.ner_crf_viterbi(1, 1, 1, 1, None, 4, 8, 33, None) == -2  # K > 32 unsupported
    assert h.ner_crf_loglik_fwd(None, None, None, None, None, None, None, 4, 8, 10, 0, None) == -1
    assert h.ner_gemm_bf16(None, None, None, None, None, 128, 128, 64, 1, 0, None) == -1
