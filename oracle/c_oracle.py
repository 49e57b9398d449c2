"""ctypes wrapper of oracle/scan_oracle.c (TEST INFRASTRUCTURE / CPU BASELINE, see oracle/__init__.py)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "liboracle_scan.so")
_lib = None


def available() -> bool:
    return os.path.exists(_PATH)


def _load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_PATH)
        _lib.oracle_ivfpq_topk.restype = C.c_int
        _lib.oracle_ivfpq_topk.argtypes = [C.c_void_p] * 6 + [C.c_int64] + [C.c_int] * 4 + [C.c_void_p] * 2 + [C.c_int]
    return _lib


def ivfpq_topk(storage, lut, is_empty, cell_start_g, cell_size_g, n_probe_list, k, nthreads=0):
    """Same contract as oracle.ivfpq_oracle.ivfpq_topk; returns (values, address, threads_used)."""
    lib = _load()
    storage = np.ascontiguousarray(storage, np.uint8)
    lut = np.ascontiguousarray(lut, np.float32)
    is_empty = np.ascontiguousarray(is_empty, np.uint8)
    cs = np.ascontiguousarray(cell_start_g, np.int64)
    cz = np.ascontiguousarray(cell_size_g, np.int64)
    npl = np.ascontiguousarray(n_probe_list, np.int64)
    M4, cap, four = storage.shape
    M = M4 * 4
    nq, n_probe = cs.shape
    assert lut.shape == (M, nq, 256)
    vals = np.empty((nq, k), np.float32)
    addr = np.empty((nq, k), np.int64)
    used = lib.oracle_ivfpq_topk(storage.ctypes.data, lut.ctypes.data, is_empty.ctypes.data, cs.ctypes.data,
                                 cz.ctypes.data, npl.ctypes.data, cap, M, nq, n_probe, k,
                                 vals.ctypes.data, addr.ctypes.data, int(nthreads))
    return vals, addr, used
