"""ctypes access to the plain-C oracle (oracle/nts_oracle.c). TEST INFRASTRUCTURE: only tests/, smoke() and the
cpu-baseline legs of bench.py may use it."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_build", "libnts_oracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "port"])
        _lib = C.CDLL(SO)
        _lib.nts_oracle_segment_gather_sum.argtypes = [C.c_void_p] * 5 + [C.c_uint32, C.c_uint32, C.c_uint32]
        _lib.nts_oracle_segment_gather_sum.restype = None
        _lib.nts_oracle_gather_msg_to_dst.argtypes = [C.c_void_p] * 3 + [C.c_uint32, C.c_uint32]
        _lib.nts_oracle_gather_msg_to_dst.restype = None
        _lib.nts_oracle_edge_softmax.argtypes = [C.c_void_p] * 3 + [C.c_uint32, C.c_uint32]
        _lib.nts_oracle_edge_softmax.restype = None
        _lib.nts_oracle_norm_degree.argtypes = [C.c_void_p] * 5 + [C.c_size_t]
        _lib.nts_oracle_norm_degree.restype = None
    return _lib


def segment_gather_sum(offsets, indices, w, X, base=0, out=None):
    offsets = np.ascontiguousarray(offsets, dtype=np.uint32)
    indices = np.ascontiguousarray(indices, dtype=np.uint32)
    X = np.ascontiguousarray(X, dtype=np.float32)
    n_rows = offsets.shape[0] - 1
    if out is None:
        out = np.zeros((n_rows, X.shape[1]), dtype=np.float32)
    wp = None
    if w is not None:
        w = np.ascontiguousarray(w, dtype=np.float32)
        wp = w.ctypes.data
    lib().nts_oracle_segment_gather_sum(offsets.ctypes.data, indices.ctypes.data, wp, X.ctypes.data,
                                        out.ctypes.data, base, n_rows, X.shape[1])
    return out


def edge_softmax(column_offset, m):
    column_offset = np.ascontiguousarray(column_offset, dtype=np.uint32)
    m = np.ascontiguousarray(m, dtype=np.float32)
    a = np.zeros_like(m)
    lib().nts_oracle_edge_softmax(column_offset.ctypes.data, m.ctypes.data, a.ctypes.data,
                                  column_offset.shape[0] - 1, m.shape[1])
    return a
