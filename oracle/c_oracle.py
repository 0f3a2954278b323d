"""ctypes front-end of ``oracle/grf_oracle.c`` (TEST INFRASTRUCTURE ONLY)."""

from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libgrf_oracle.so")
_lib = None

DRAW_TRACE, DRAW_PHILOX = 0, 1


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "grf_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(
            ["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", _SO, src, "-lm"]
        )
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.grf_oracle_build.restype = ctypes.c_void_p
        _lib.grf_oracle_build.argtypes = [
            ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_int64, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_double,
            ctypes.c_int32, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_int32, ctypes.c_int32,
        ]
        _lib.grf_oracle_nnz.restype = ctypes.c_int64
        _lib.grf_oracle_nnz.argtypes = [ctypes.c_void_p, ctypes.c_int32]
        _lib.grf_oracle_visits.restype = ctypes.c_int64
        _lib.grf_oracle_visits.argtypes = [ctypes.c_void_p]
        _lib.grf_oracle_copy.restype = None
        _lib.grf_oracle_copy.argtypes = [ctypes.c_void_p, ctypes.c_int32] + [ctypes.c_void_p] * 3
        _lib.grf_oracle_free.restype = None
        _lib.grf_oracle_free.argtypes = [ctypes.c_void_p]
        _lib.grf_oracle_philox.restype = None
        _lib.grf_oracle_philox.argtypes = [ctypes.c_void_p] * 3
    return _lib


def philox(ctr, key):
    c = np.ascontiguousarray(ctr, dtype=np.uint32)
    k = np.ascontiguousarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().grf_oracle_philox(c.ctypes.data, k.ctypes.data, out.ctypes.data)
    return out


def step_matrices(adj_csr, num_walks, p_halt, max_walk_length, draw_mode=DRAW_PHILOX, seed=42,
                  trace=None, load_mode=0, scale_mode=0, start_lo=0, start_hi=None, return_visits=False):
    """Step matrices M_l[start_lo:start_hi, :] as scipy CSR (float64, int32,
    sorted columns).  ``trace`` = (trace_u, trace_k) for ``DRAW_TRACE``."""
    a = adj_csr.tocsr()
    n = a.shape[0]
    start_hi = n if start_hi is None else start_hi
    indptr = np.ascontiguousarray(a.indptr, dtype=np.int32)
    indices = np.ascontiguousarray(a.indices, dtype=np.int32)
    data = np.ascontiguousarray(a.data, dtype=np.float64)
    tu = tk = None
    if draw_mode == DRAW_TRACE:
        tu = np.ascontiguousarray(trace[0], dtype=np.float64)
        tk = np.ascontiguousarray(trace[1], dtype=np.int32)
    L = max_walk_length
    h = lib().grf_oracle_build(
        n, indptr.ctypes.data, indices.ctypes.data, data.ctypes.data, start_lo, start_hi, num_walks, L,
        float(p_halt), draw_mode, int(seed), tu.ctypes.data if tu is not None else None,
        tk.ctypes.data if tk is not None else None, load_mode, scale_mode,
    )
    try:
        mats = []
        n_local = start_hi - start_lo
        for s in range(L):
            nnz = lib().grf_oracle_nnz(h, s)
            ip = np.zeros(n_local + 1, dtype=np.int64)
            ix = np.zeros(nnz, dtype=np.int32)
            dv = np.zeros(nnz, dtype=np.float64)
            lib().grf_oracle_copy(h, s, ip.ctypes.data, ix.ctypes.data, dv.ctypes.data)
            mats.append(sp.csr_matrix((dv, ix, ip), shape=(n_local, n)))
        visits = lib().grf_oracle_visits(h)
    finally:
        lib().grf_oracle_free(h)
    return (mats, visits) if return_visits else mats
