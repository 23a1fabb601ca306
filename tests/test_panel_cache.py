"""On-disk packed-panel cache (tdsfs_pack.save_panel / load_panel / cached_pack_vcf): replaces the reference's pickle + bz2
cache of the parsed dict (scripts/twoDSFS.py:505-510, scripts/src/twoDSFS_class.py:1918-1919)."""
import os
import shutil

import numpy as np
import pytest

from helpers import GOLDEN


def _same(a, b):
    assert (a.n, a.W1, a.W2, a.ns1, a.ns2, a.last_key_row, a.n_records, a.n_skipped) == \
           (b.n, b.W1, b.W2, b.ns1, b.ns2, b.last_key_row, b.n_records, b.n_skipped)
    assert list(a.chroms) == list(b.chroms) and tuple(a.pops) == tuple(b.pops)
    assert np.array_equal(np.asarray(a.G), np.asarray(b.G)) and np.asarray(b.G).dtype == np.uint32
    assert np.array_equal(a.pos, b.pos) and np.array_equal(a.off, b.off)
    assert list(a.ann) == list(b.ann)
    if a.fixups is None:
        assert b.fixups is None
    else:
        assert np.array_equal(a.fixups, b.fixups)
    assert np.array_equal(a.counts(), b.counts())


@pytest.mark.parametrize("name", ["ecb_subset", "ingest_small"])
@pytest.mark.parametrize("mmap", [True, False])
def test_round_trip(tmp_path, name, mmap):
    from tdsfs_pack import PackedPanel, pack_vcf
    vcf, popmap = os.path.join(GOLDEN, f"{name}.vcf.gz"), os.path.join(GOLDEN, f"{name}.popmap.txt")
    pops = ("uv", "bv")
    P = pack_vcf(vcf, popmap, *pops)
    path = str(tmp_path / "panel.tdsfspk")
    P.save(path, sources=(vcf, popmap))
    Q = PackedPanel.load(path, mmap=mmap)
    _same(P, Q)
    if mmap and P.n:
        assert isinstance(Q.G, np.memmap) and Q.G.offset % 4096 == 0  # mapped, page aligned: no copy of the matrix


def test_cached_pack_reuses_and_invalidates(tmp_path):
    import tdsfs_pack as TP
    vcf, popmap = str(tmp_path / "x.vcf.gz"), str(tmp_path / "x.popmap.txt")
    shutil.copy(os.path.join(GOLDEN, "ecb_subset.vcf.gz"), vcf)
    shutil.copy(os.path.join(GOLDEN, "ecb_subset.popmap.txt"), popmap)
    P = TP.cached_pack_vcf(vcf, popmap, "uv", "bv")
    cache = vcf + ".uv.bv.tdsfspk"
    assert os.path.exists(cache)
    calls = []
    real = TP.pack_vcf
    TP.pack_vcf = lambda *a, **k: (calls.append(a), real(*a, **k))[1]
    try:
        Q = TP.cached_pack_vcf(vcf, popmap, "uv", "bv")
        assert not calls, "an up-to-date cache must not re-parse the VCF"
        _same(P, Q)
        st = os.stat(popmap)
        os.utime(popmap, ns=(st.st_atime_ns, st.st_mtime_ns + 10 ** 9))  # a touched source invalidates the cache
        R = TP.cached_pack_vcf(vcf, popmap, "uv", "bv")
        assert len(calls) == 1
        _same(P, R)
    finally:
        TP.pack_vcf = real


def test_rejects_foreign_file(tmp_path):
    from tdsfs_pack import load_panel
    p = tmp_path / "junk"
    p.write_bytes(b"not a cache at all")
    with pytest.raises(ValueError):
        load_panel(str(p))
