"""ctypes wrapper of oracle/libsfs_oracle.so (C restatement of the hot path).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(os.path.join(HERE, "libsfs_oracle.so"))
        _lib.oracle_scan.restype = C.c_int64
        _lib.oracle_max_threads.restype = C.c_int
    return _lib


def max_threads():
    return int(lib().oracle_max_threads())


def decode(G, S, W1, W2, ns1, ns2, nthreads=1):
    """G: genotype matrix in the B32 layout (flat uint32, ceil(S/32)*(W1+W2)*32 words)."""
    G = np.ascontiguousarray(G, dtype=np.uint32).reshape(-1)
    assert G.size >= (S + 31) // 32 * (W1 + W2) * 32
    cnt = np.zeros((S, 4), dtype=np.uint16)
    lib().oracle_decode_mt(G.ctypes.data_as(C.c_void_p), C.c_int64(S), C.c_int(W1), C.c_int(W2), C.c_int(ns1), C.c_int(ns2),
                           cnt.ctypes.data_as(C.c_void_p), C.c_int(nthreads))
    return cnt


def scan(cnt, pos, off, n1, n2, W=None, N=None, fold=True, bg="per_chrom", nthreads=0, include_flags=None):
    cnt = np.ascontiguousarray(cnt, dtype=np.uint16)
    pos = np.ascontiguousarray(pos, dtype=np.int32)
    off = np.ascontiguousarray(off, dtype=np.int64)
    S = len(pos)
    cap = S + 1
    o = dict(chrom=np.zeros(cap, np.int32), start=np.zeros(cap, np.int64), end=np.zeros(cap, np.int64), snp_count=np.zeros(cap, np.int32),
             T2D=np.zeros(cap), T1D_p1=np.zeros(cap), T1D_p2=np.zeros(cap), none=np.zeros(cap, np.uint8))
    inc = None if include_flags is None else np.ascontiguousarray(include_flags, dtype=np.uint8)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    n = lib().oracle_scan(p(cnt), p(pos), p(off), C.c_int(len(off) - 1), C.c_int(n1), C.c_int(n2), C.c_int(1 if fold else 0),
                          C.c_int64(W if W is not None else N), C.c_int(0 if W is not None else 1), C.c_int(1 if bg == "genome" else 0),
                          C.c_int(nthreads), p(inc) if inc is not None else None, C.c_int64(cap), p(o["chrom"]), p(o["start"]),
                          p(o["end"]), p(o["snp_count"]), p(o["T2D"]), p(o["T1D_p1"]), p(o["T1D_p2"]), p(o["none"]))
    if n < 0:
        raise KeyError("allele count exceeds 2n")
    out = {k: v[:n] for k, v in o.items()}
    out["T2D_none"] = (out["none"] & 1) != 0
    out["T1D_p1_none"] = (out["none"] & 2) != 0
    out["T1D_p2_none"] = (out["none"] & 4) != 0
    return out
