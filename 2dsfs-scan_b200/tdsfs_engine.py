"""Host-side engine between the reference's dict-based Python API and libtdsfs.so.

Everything numeric (spectra, window assignment, likelihoods) runs on the GPU through the C ABI (tdsfs_capi); this
module only converts the reference's `data_dict` (scripts/src/twoDSFS_class.py:132-134) to arrays, builds the
per-SNP filter flags, and turns struct-of-array results back into the reference's dict-of-dicts, applying the
quirk ledger of SURVEY.md section 9 (None, +inf, stale-carry) that is observable in the reference's outputs.
"""
from __future__ import annotations

import math

import numpy as np

import tdsfs_capi as T

_EPS10 = float(np.finfo(np.float64).eps * 10)

_conv = False


def _dictconv():
    """ctypes.PyDLL binding of lib/libtdsfs_dictconv.so (optional accelerator of the dict -> array conversion)."""
    global _conv
    if _conv is False:
        import ctypes
        import os
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libtdsfs_dictconv.so")
        _conv = None
        if os.path.exists(path) and not os.environ.get("TDSFS_NO_DICTCONV"):
            try:
                f = ctypes.PyDLL(path).tdsfs_dict_to_arrays
                f.restype = ctypes.c_int
                f.argtypes = [ctypes.py_object, ctypes.py_object, ctypes.py_object, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                              ctypes.c_void_p, ctypes.py_object, ctypes.py_object]
                _conv = f
            except OSError:
                _conv = None
    return _conv


class SnpTable:
    """A data_dict in array form, sorted by (chromosome string, position) as every reference scanner sorts it
    (e.g. :828-835).  cnt[S,4] = (ref1, alt1, ref2, alt2) for the two requested populations; a population
    missing from a record counts (0, 0) (:190-191)."""

    __slots__ = ("chroms", "off", "pos", "cnt", "ann", "keys", "pops", "last_key_row", "n")

    @classmethod
    def from_dict(cls, data_dict, pop1, pop2):
        conv = _dictconv()
        if conv is not None and type(data_dict) is dict:
            return cls._from_dict_c(conv, data_dict, pop1, pop2)
        return cls._from_dict_py(data_dict, pop1, pop2)

    @classmethod
    def _from_dict_c(cls, conv, data_dict, pop1, pop2):
        """The per-SNP loop in C (csrc/dictconv.c); same semantics and exceptions as _from_dict_py."""
        n = len(data_dict)
        pos = np.empty(n, dtype=np.int64)
        cnt = np.zeros((n, 4), dtype=np.int64)
        cidx = np.empty(n, dtype=np.int32)
        aidx = np.empty(n, dtype=np.int32)
        names, vocab = [], []
        conv(data_dict, pop1, pop2, pos.ctypes.data, cnt.ctypes.data, cidx.ctypes.data, aidx.ctypes.data, names, vocab)
        keys = list(data_dict.keys())
        rank = np.argsort(np.argsort(np.array(names, dtype=object), kind="stable"), kind="stable") if names else np.zeros(0, np.int64)
        cidx = rank[cidx] if n else cidx.astype(np.int64)
        order = np.lexsort((pos, cidx))
        self = cls()
        self.chroms = sorted(names)
        self.pos = pos[order]
        self.cnt = cnt[order]
        va = np.empty(len(vocab), dtype=object)
        va[:] = vocab
        self.ann = va[aidx][order] if n else np.array([], dtype=object)
        self.keys = [keys[i] for i in order.tolist()]
        self.off = np.concatenate([[0], np.cumsum(np.bincount(cidx, minlength=len(names)))]).astype(np.int64) if n else np.zeros(1, np.int64)
        self.pops = (pop1, pop2)
        self.n = n
        self.last_key_row = int(np.flatnonzero(order == n - 1)[0]) if n else -1
        return self

    @classmethod
    def _from_dict_py(cls, data_dict, pop1, pop2):
        n = len(data_dict)
        chrom = [None] * n
        pos = np.empty(n, dtype=np.int64)
        cnt = np.zeros((n, 4), dtype=np.int64)
        ann = [None] * n
        keys = list(data_dict.keys())
        for i, k in enumerate(keys):
            parts = k.split("-")
            if len(parts) != 2:
                # the reference unpacks `chr_id, pos = snp_id.split('-')` (:176)
                raise ValueError("too many values to unpack (expected 2)" if len(parts) > 2 else
                                 "not enough values to unpack (expected 2, got 1)")
            chrom[i] = parts[0]
            pos[i] = int(parts[1])
            info = data_dict[k]
            calls = info["calls"]
            c1 = calls.get(pop1, (0, 0))
            c2 = calls.get(pop2, (0, 0))
            cnt[i, 0], cnt[i, 1], cnt[i, 2], cnt[i, 3] = c1[0], c1[1], c2[0], c2[1]
            ann[i] = info.get("annotation")
        self = cls()
        names = sorted(set(chrom))
        cid = {c: j for j, c in enumerate(names)}
        cidx = np.fromiter((cid[c] for c in chrom), dtype=np.int64, count=n)
        order = np.lexsort((pos, cidx))  # stable: ties keep insertion order, like list.sort on (chrom, pos)
        self.chroms = names
        self.pos = pos[order]
        self.cnt = cnt[order]
        self.ann = np.array(ann, dtype=object)[order] if n else np.array([], dtype=object)
        self.keys = [keys[i] for i in order.tolist()]
        self.off = np.concatenate([[0], np.cumsum(np.bincount(cidx, minlength=len(names)))]).astype(np.int64) if n else np.zeros(1, np.int64)
        self.pops = (pop1, pop2)
        self.n = n
        # row (in sorted order) of the LAST-INSERTED key: needed by the T2D_scan quirk
        self.last_key_row = int(np.flatnonzero(order == n - 1)[0]) if n else -1
        return self

    def check_ranges(self):
        if self.n and (self.pos.min() < 0 or self.pos.max() > 2 ** 31 - 2):
            raise OverflowError("positions must fit in int32")
        if self.n and self.cnt.max() > 65535:
            raise OverflowError("allele counts must fit in uint16")


def filter_flags(table: SnpTable, start, end, variant_type):
    """snp_flags of the C ABI: bit0 = passes the spectrum filters (:179-187), bit1 = counted by count_snps (:298-301)."""
    if start is None and end is None and variant_type is None:
        return None
    inc = np.ones(table.n, dtype=bool)
    if start is not None:
        inc &= table.pos >= start
    if end is not None:
        inc &= table.pos <= end
    cntb = np.ones(table.n, dtype=bool)
    if variant_type is not None:
        m = np.fromiter((a == variant_type for a in table.ann), dtype=bool, count=table.n)
        inc &= m
        cntb = m
    return (inc.astype(np.uint8) | (cntb.astype(np.uint8) << 1)).astype(np.uint8)


class Engine:
    """One GPU handle + the conversions.  Not thread-safe (neither is the reference)."""

    def __init__(self, device=0):
        self.h = T.Handle(device)
        self._panel = None

    def close(self):
        self.h.close()

    # ---- loading
    def load(self, table: SnpTable, n1, n2, fold, flags=None, cnt=None):
        table.check_ranges()
        if self._panel != (n1, n2, bool(fold)):
            self.h.set_panel(n1, n2, fold)
            self._panel = (n1, n2, bool(fold))
        if getattr(table, "G", None) is not None and cnt is None:  # PackedPanel: genotype-level entry (K1 from 2-bit calls)
            self.h.load_genotypes(table.G, table.n, table.W1, table.W2, table.ns1, table.ns2,
                                  np.ascontiguousarray(table.pos, dtype=np.int32), table.off, fixups=table.fixups, flags=flags)
            return
        c = table.cnt if cnt is None else cnt
        self.h.load_counts(np.ascontiguousarray(c, dtype=np.uint16), np.ascontiguousarray(table.pos, dtype=np.int32), table.off, flags)

    def background(self, mode, chrom=0, lo=-1, hi=-1):
        try:
            self.h.background(mode, chrom, lo, hi)
        except T.TdsfsError as e:
            if e.code == T.ERR_RANGE:
                # the reference's calculate_1d_sfs does sfs_dict[alt_count] += 1 on a dict with keys 0..2n (:433)
                raise KeyError("allele count exceeds 2 * pop_size") from None
            raise

    # ---- spectra
    def spectra(self, table, n1, n2, fold, flags):
        """Integer spectra of the whole table: (2D [2n1+1, 2n2+1], raw 1D pop1, raw 1D pop2)."""
        self.load(table, n1, n2, fold, flags)
        self.background(T.BG_GENOME)
        return self.h.get_background(0)

    # ---- likelihood of explicit spectra (calculate_likelihood_1D / _2D)
    def likelihood(self, fg, bg, guarded=True):
        bins = sorted(fg.keys())[1:-1]
        x = [int(fg[k]) for k in bins]
        n = sum(x)
        if not bins and not guarded:
            # an empty interior (1D spectrum of a single diploid) never divides: scipy's logpmf of empty vectors is NaN
            return math.nan
        if n == 0:
            if guarded:
                return None
            raise ZeroDivisionError("division by zero")  # sims_scan.py:348 count/total_fg
        b = [bg[k] for k in bins]
        B = sum(b)
        if B == 0:
            if guarded:
                return None
            raise ZeroDivisionError("division by zero")  # sims_scan.py:371 / float division
        ba = np.asarray(b, dtype=np.float64)
        # scipy's domain gate (multinomial._process_parameters): NaN when p is not a probability vector
        p = ba / float(B)
        if abs(1.0 - p.sum()) > _EPS10 or np.any(p < 0) or np.any(p > 1):
            return math.nan
        val, none = self.h.likelihood(np.asarray(x, dtype=np.int64), ba, float(B))
        return None if none else val

    # ---- scans
    def scan(self, size, snp_mode=False):
        return self.h.scan(size, snp_mode=snp_mode)


def window_keys(table: SnpTable, res, live):
    """'{chrom} {start}-{end}' labels (:936) of the emitted windows."""
    ch, st, en = res["chrom"][live].tolist(), res["start"][live].tolist(), res["end"][live].tolist()
    names = table.chroms
    return [f"{names[c]} {s}-{e}" for c, s, e in zip(ch, st, en)]


def stat_lists(res, live):
    """T2D / T1D_pop1 / T1D_pop2 as Python floats with None where the reference returns None."""
    fl = res["flags"][live]
    out = []
    for name, bit in (("T2D", T.F_T2D_NONE), ("T1D_p1", T.F_T1D_P1_NONE), ("T1D_p2", T.F_T1D_P2_NONE)):
        vals = res[name][live].tolist()
        none = ((fl & bit) != 0).tolist()
        out.append([None if nn else v for v, nn in zip(vals, none)])
    return out


def dense2d_to_dict(h2):
    """row-major (i, j) insertion order, as the reference initialises it (:161-163)"""
    R1, R2 = h2.shape
    flat = h2.ravel().tolist()
    return {(i, j): flat[i * R2 + j] for i in range(R1) for j in range(R2)}


def dict_to_dense2d(d, n1, n2, what="background_2d_sfs"):
    """Background dict -> dense doubles; the scorer reads bg[k] for every interior key of the dense foreground (:659-661)."""
    R1, R2 = 2 * n1 + 1, 2 * n2 + 1
    out = np.zeros(R1 * R2, dtype=np.float64)
    try:
        for i in range(R1):
            base = i * R2
            for j in range(R2):
                if (i == 0 and j == 0) or (i == R1 - 1 and j == R2 - 1):
                    out[base + j] = d.get((i, j), 0)
                else:
                    out[base + j] = d[(i, j)]
    except KeyError as e:
        raise KeyError(e.args[0]) from None
    return out


def dict_to_folded1d(d, n):
    """bg[k] for the folded foreground's interior keys k = 1..n-1 (:511-513); bins 0 and n are never read."""
    out = np.zeros(n + 1, dtype=np.float64)
    for k in range(1, n):
        out[k] = d[k]
    return out
