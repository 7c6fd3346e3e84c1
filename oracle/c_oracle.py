"""ctypes loader for oracle/libhades_oracle.so (test infrastructure; see hades_ref.c header).
Arrays are numpy uint64 of shape (..., 4): BlsScalar.0 limbs (Montgomery form)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "libhades_oracle.so"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libhades_oracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.oracle_init()
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _c(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return a


def permute(states, threads=1):
    s = _c(states).copy().reshape(-1, 5, 4)
    if threads > 1:
        lib().oracle_permute_mt(_p(s), ctypes.c_size_t(s.shape[0]), ctypes.c_int(threads))
    else:
        lib().oracle_permute(_p(s), ctypes.c_size_t(s.shape[0]))
    return s


def digest(tag, inp, in_len, out_len=1, threads=1, out=None):
    """out: optional preallocated (n, out_len, 4) uint64 array (lets a caller time the C call alone)."""
    tag = _c(tag).reshape(4)
    inp = _c(inp).reshape(-1, in_len, 4)
    n = inp.shape[0]
    if out is None:
        out = np.zeros((n, out_len, 4), dtype=np.uint64)
    assert out.dtype == np.uint64 and out.flags.c_contiguous and out.shape == (n, out_len, 4)
    if threads > 1:
        lib().oracle_digest_mt(_p(tag), _p(inp), ctypes.c_size_t(n), ctypes.c_size_t(in_len), _p(out),
                               ctypes.c_size_t(out_len), ctypes.c_int(threads))
    else:
        lib().oracle_digest(_p(tag), _p(inp), ctypes.c_size_t(n), ctypes.c_size_t(in_len), _p(out),
                            ctypes.c_size_t(out_len))
    return out


def digest_padded(tag, inp, in_len, pad):
    tag = _c(tag).reshape(4)
    pad = _c(pad).reshape(4)
    inp = _c(inp).reshape(-1, in_len, 4)
    n = inp.shape[0]
    out = np.zeros((n, 4), dtype=np.uint64)
    lib().oracle_digest_padded(_p(tag), _p(inp), ctypes.c_size_t(n), ctypes.c_size_t(in_len), _p(pad),
                               _p(out))
    return out


def encrypt(tag, msg, L, secret_uv, nonce):
    tag = _c(tag).reshape(4)
    msg = _c(msg).reshape(-1, L, 4)
    n = msg.shape[0]
    secret_uv = _c(secret_uv).reshape(n, 2, 4)
    nonce = _c(nonce).reshape(n, 4)
    cipher = np.zeros((n, L + 1, 4), dtype=np.uint64)
    lib().oracle_encrypt(_p(tag), _p(msg), ctypes.c_size_t(n), ctypes.c_size_t(L), _p(secret_uv),
                         _p(nonce), _p(cipher))
    return cipher


def decrypt(tag, cipher, L, secret_uv, nonce):
    tag = _c(tag).reshape(4)
    cipher = _c(cipher).reshape(-1, L + 1, 4)
    n = cipher.shape[0]
    secret_uv = _c(secret_uv).reshape(n, 2, 4)
    nonce = _c(nonce).reshape(n, 4)
    msg = np.zeros((n, L, 4), dtype=np.uint64)
    ok = np.zeros(n, dtype=np.uint8)
    lib().oracle_decrypt(_p(tag), _p(cipher), ctypes.c_size_t(n), ctypes.c_size_t(L), _p(secret_uv),
                         _p(nonce), _p(msg), _p(ok))
    return msg, ok
