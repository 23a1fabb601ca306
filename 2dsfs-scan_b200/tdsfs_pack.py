"""Host-side packing of genotype calls into the 2-bit-per-call matrix the GPU consumes (K0).

Per SNP: uint32 words of 16 calls (call i in bits 2i..2i+1), population-1 words then population-2 words, each
population padded with zeros to a whole word.  Codes: 0 = 0/0, 1 = 0/1, 3 = 1/1, 2 = missing, so that
alt = popcount(block) - #missing  and  ref = 2*(samples - #missing) - alt.
Memory layout "B32" (block-transposed): SNPs are grouped in blocks of 32; word w of SNP s is stored at uint32 index
((s // 32) * RW + w) * 32 + s % 32, RW = W1 + W2.  A warp therefore reads word w of 32 consecutive SNPs as one
128-byte line, and in shared memory as one conflict-free access.  The last block is zero padded.
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


def to_b32(rows: np.ndarray) -> np.ndarray:
    """row-major words [S, RW] -> B32 buffer (flat uint32, ceil(S/32)*RW*32 words)."""
    S, RW = rows.shape
    nb = (S + 31) // 32
    pad = np.zeros((nb * 32, RW), dtype=np.uint32)
    pad[:S] = rows
    return np.ascontiguousarray(pad.reshape(nb, 32, RW).transpose(0, 2, 1)).reshape(-1)


def from_b32(buf: np.ndarray, S: int, RW: int) -> np.ndarray:
    """B32 buffer -> row-major words [S, RW]."""
    nb = (S + 31) // 32
    return np.ascontiguousarray(np.asarray(buf, dtype=np.uint32).reshape(nb, RW, 32).transpose(0, 2, 1)).reshape(nb * 32, RW)[:S]


def pack_codes(codes1: np.ndarray, codes2: np.ndarray):
    """codes[S, ns] of 2-bit codes -> (G in B32 layout (flat uint32), W1, W2)."""
    b1, b2 = _pack_block(np.asarray(codes1)), _pack_block(np.asarray(codes2))
    return to_b32(np.concatenate([b1, b2], axis=1)), b1.shape[1], b2.shape[1]
