"""CPU-only tests of the host logic of the drop-in layer: VCF ingest gates, dict <-> array conversion, filters."""
import json
import os

import numpy as np

from helpers import GOLDEN


def test_parse_vcf_matches_reference_ingest():
    import twoDSFS_class as K
    exp = json.load(open(os.path.join(GOLDEN, "ingest_small.json")))["data_dict"]
    d = K.parse_vcf_to_dict(os.path.join(GOLDEN, "ingest_small.vcf.gz"), os.path.join(GOLDEN, "ingest_small.popmap.txt"))
    got = [[k, list(v["segregating"]), v["context"], {p: list(c) for p, c in v["calls"].items()}, v["annotation"]] for k, v in d.items()]
    assert got == exp


def test_parse_ecb_subset_counts():
    import twoDSFS_class as K
    exp = json.load(open(os.path.join(GOLDEN, "ecb_subset.json")))
    d = K.parse_vcf_to_dict(os.path.join(GOLDEN, "ecb_subset.vcf.gz"), os.path.join(GOLDEN, "ecb_subset.popmap.txt"))
    assert [[k, list(v["calls"]["uv"]), list(v["calls"]["bv"])] for k, v in d.items()] == exp["counts"]


def test_snp_table_sorting_and_flags():
    from tdsfs_engine import SnpTable, filter_flags
    d = {"chr2-50": {"calls": {"a": (3, 1), "b": (2, 2)}, "annotation": "x"},
         "chr10-7": {"calls": {"a": (4, 0)}, "annotation": "y"},
         "chr2-5": {"calls": {"b": (1, 3)}, "annotation": "x"},
         "chr10-70": {"calls": {"a": (0, 4), "b": (4, 0)}, "annotation": "x"}}
    t = SnpTable.from_dict(d, "a", "b")
    assert t.chroms == ["chr10", "chr2"]  # string order, as the reference sorts
    assert t.keys == ["chr10-7", "chr10-70", "chr2-5", "chr2-50"]
    assert t.off.tolist() == [0, 2, 4] and t.pos.tolist() == [7, 70, 5, 50]
    assert t.cnt.tolist() == [[4, 0, 0, 0], [0, 4, 4, 0], [0, 0, 1, 3], [3, 1, 2, 2]]
    assert t.last_key_row == 1
    assert filter_flags(t, None, None, None) is None
    assert filter_flags(t, 6, 60, "x").tolist() == [0, 2, 2, 3]
    assert filter_flags(t, None, 60, None).tolist() == [3, 2, 3, 3]


def test_b32_layout_roundtrip():
    from tdsfs_pack import from_b32, to_b32
    rng = np.random.default_rng(0)
    rows = rng.integers(0, 2 ** 32, size=(70, 5), dtype=np.uint64).astype(np.uint32)
    buf = to_b32(rows)
    assert buf.size == 3 * 5 * 32
    assert buf[(40 // 32 * 5 + 3) * 32 + 40 % 32] == rows[40, 3]
    assert np.array_equal(from_b32(buf, 70, 5), rows)
