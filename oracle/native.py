"""ctypes binding + build recipe of oracle/crf_c.c — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The C restatement follows oracle/crf.py line by line (same add order, same tie rule); it is the fast form of the checker
for the roofline-sized CRF launches (all 262 144 rows in seconds).  `build()` compiles it with gcc into
oracle/liboracle_crf.so (git-ignored, travels to the GPU box with the snapshot); `available()` says whether the library
can be loaded — callers fall back to the numpy restatement on a row sample when it cannot.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "crf_c.c")
LIB = os.path.join(HERE, "liboracle_crf.so")
_lib = None


def build(force=False, verbose=False):
    """gcc -O2 -ffp-contract=off (float32 sums must round as numpy's do) [-fopenmp] -shared -fPIC.  -> path of the .so"""
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    base = ["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"]
    for cmd in (base[:3] + ["-fopenmp"] + base[3:], base):        # OpenMP when libgomp is there, single-threaded otherwise
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose:
            print(" ".join(cmd), "->", r.returncode, r.stderr.strip()[:400])
        if r.returncode == 0:
            return LIB
    raise RuntimeError("oracle/crf_c.c did not compile: " + r.stderr[:400])


def lib():
    global _lib
    if _lib is None:
        h = ctypes.CDLL(LIB)
        vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
        h.oracle_crf_decode_f32.restype = ctypes.c_int
        h.oracle_crf_decode_f32.argtypes = [vp, vp, vp, vp, vp, i64, i32, i32]
        h.oracle_crf_loglik_f64.restype = ctypes.c_int
        h.oracle_crf_loglik_f64.argtypes = [vp, vp, vp, vp, vp, i64, i32, i32]
        h.oracle_crf_abi.restype = ctypes.c_int
        assert h.oracle_crf_abi() == 1
        _lib = h
    return _lib


def available():
    try:
        lib()
        return True
    except (OSError, AssertionError):
        return False


def _c(a, dtype):
    return np.ascontiguousarray(np.asarray(a), dtype=dtype)


def crf_decode(potentials, transition_params, sequence_length):
    """oracle.crf.crf_decode(dtype=float32) on every row: (tags [B,T] int32, best [B] float32)."""
    x, tr, ln = _c(potentials, np.float32), _c(transition_params, np.float32), _c(sequence_length, np.int32)
    B, T, K = x.shape
    tags, best = np.empty((B, T), np.int32), np.empty((B,), np.float32)
    rc = lib().oracle_crf_decode_f32(x.ctypes.data, tr.ctypes.data, ln.ctypes.data, tags.ctypes.data, best.ctypes.data, B, T, K)
    if rc:
        raise RuntimeError(f"oracle_crf_decode_f32 -> {rc}")
    return tags, best


def crf_log_likelihood(inputs, tag_indices, sequence_lengths, transition_params):
    """oracle.crf.crf_log_likelihood(dtype=float64) on float32 inputs: ll [B] float64."""
    x, tr = _c(inputs, np.float32), _c(transition_params, np.float32)
    y, ln = _c(tag_indices, np.int32), _c(sequence_lengths, np.int32)
    B, T, K = x.shape
    ll = np.empty((B,), np.float64)
    rc = lib().oracle_crf_loglik_f64(x.ctypes.data, y.ctypes.data, ln.ctypes.data, tr.ctypes.data, ll.ctypes.data, B, T, K)
    if rc:
        raise RuntimeError(f"oracle_crf_loglik_f64 -> {rc}")
    return ll


if __name__ == "__main__":
    print(build(force=True, verbose=True))
