"""ctypes wrapper over oracle/dtw_oracle.c.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED -- see the header of dtw_oracle.c: the reference's DTW is the
un-vendored third-party module DTW_Cython (abnet3/utils.py:14, :149-151;
requirements.txt:9) and no reference test pins a DTW result.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libabn_oracle.so")
_lib = None


def build(force=False):
    """Compile dtw_oracle.c with gcc (recipe: oracle/Makefile)."""
    src = os.path.join(_HERE, "dtw_oracle.c")
    if (not force and os.path.exists(_SO)
            and os.path.getmtime(_SO) >= os.path.getmtime(src)):
        return _SO
    subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_SO)
        lib.abn_oracle_dtw.restype = ctypes.c_int
        lib.abn_oracle_dtw.argtypes = [
            ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_void_p, ctypes.c_void_p]
        lib.abn_oracle_path_cost.restype = ctypes.c_double
        lib.abn_oracle_path_cost.argtypes = [
            ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        _lib = lib
    return _lib


def dtw(dist_array, return_acc=False, return_ties=False):
    """DTW on a float64 local-distance matrix.

    Returns ``(cost, path1, path2)`` (+ accumulated-cost matrix, + tie count);
    raises ``ValueError`` when ``dist_array`` holds a NaN / negative entry (the
    reference's ``assert np.all(d >= 0)``, abnet3/utils.py:59).
    """
    lib = _load()
    D = np.ascontiguousarray(dist_array, dtype=np.float64)
    n1, n2 = D.shape
    p1 = np.empty(n1 + n2 - 1, dtype=np.int32)
    p2 = np.empty(n1 + n2 - 1, dtype=np.int32)
    cost = ctypes.c_double(0.0)
    ties = ctypes.c_int32(0)
    acc = np.empty((n1, n2), dtype=np.float64) if return_acc else None
    L = lib.abn_oracle_dtw(
        D.ctypes.data, n1, n2, n2, p1.ctypes.data, p2.ctypes.data,
        ctypes.addressof(cost), acc.ctypes.data if return_acc else None,
        ctypes.addressof(ties))
    if L < 0:
        raise ValueError("invalid distance matrix (NaN or negative entry)")
    out = [cost.value, p1[:L].copy(), p2[:L].copy()]
    if return_acc:
        out.append(acc)
    if return_ties:
        out.append(int(ties.value))
    return tuple(out)


def path_cost(dist_array, path1, path2):
    """Cost of a given path under ``dist_array`` (NaN if not a DTW path)."""
    lib = _load()
    D = np.ascontiguousarray(dist_array, dtype=np.float64)
    p1 = np.ascontiguousarray(path1, dtype=np.int32)
    p2 = np.ascontiguousarray(path2, dtype=np.int32)
    return lib.abn_oracle_path_cost(D.ctypes.data, D.shape[0], D.shape[1],
                                    D.shape[1], p1.ctypes.data,
                                    p2.ctypes.data, len(p1))
