"""Host-side packing of genotype calls into the 2-bit-per-call matrix the GPU consumes (K0).

Layout: SNP-major rows of uint32 words, 16 calls per word (call i in bits 2i..2i+1), population-1 block then
population-2 block, each padded with zeros to a whole word.  Codes: 0 = 0/0, 1 = 0/1, 3 = 1/1, 2 = missing,
so that  alt = popcount(block) - #missing  and  ref = 2*(samples - #missing) - alt.
Replaces the per-sample character counting of make_data_dict_vcf (scripts/src/twoDSFS_class.py:118-130)."""
from __future__ import annotations

import numpy as np

CODE_HOMREF, CODE_HET, CODE_MISSING, CODE_HOMALT = 0, 1, 2, 3


def words_for(n_samples: int) -> int:
    return max(1, (n_samples + 15) // 16)


def _pack_block(codes: np.ndarray) -> np.ndarray:
    S, ns = codes.shape
    W = words_for(ns)
    padded = np.zeros((S, W * 16), dtype=np.uint32)
    padded[:, :ns] = codes
    shifts = (2 * np.arange(16, dtype=np.uint32))[None, None, :]
    return (padded.reshape(S, W, 16) << shifts).sum(axis=2, dtype=np.uint32)


def pack_codes(codes1: np.ndarray, codes2: np.ndarray):
    """codes[S, ns] of 2-bit codes -> (G[S, W1+W2] uint32, W1, W2)."""
    b1, b2 = _pack_block(np.asarray(codes1)), _pack_block(np.asarray(codes2))
    return np.ascontiguousarray(np.concatenate([b1, b2], axis=1)), b1.shape[1], b2.shape[1]
