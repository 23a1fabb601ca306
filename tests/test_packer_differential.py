"""K0 (csrc/vcf_pack.cpp) against the oracle's restatement of the reference ingest (make_data_dict_vcf, reference
scripts/src/twoDSFS_class.py:36-138) on seeded random VCFs that exercise its quirks: positional population list when header
samples are missing from the popmap, FILTER / REF / ALT gates, FORMAT with GT anywhere, haploid / polyploid / half / multi-
allelic calls (fix-ups), records with fewer or more sample columns than the header, duplicate keys, unsorted chromosomes,
a last line without terminator -- written as plain gzip and as BGZF (bgzip's blocked gzip, inflated member-parallel)."""
import gzip
import os
import struct
import zlib

import numpy as np
import pytest

import sfs_oracle as O


def write_bgzf(data, path, block=65280):
    with open(path, "wb") as f:
        for o in list(range(0, len(data), block)) + [None]:
            chunk = b"" if o is None else data[o:o + block]
            c = zlib.compressobj(6, zlib.DEFLATED, -15)
            comp = c.compress(chunk) + c.flush()
            f.write(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, 12 + 6 + len(comp) + 8 - 1))
            f.write(comp + struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))


def random_vcf(rng, n_rec, n_samples):
    names = [f"s{i}" for i in range(n_samples)]
    in_map = rng.random(n_samples) < 0.85                       # header samples absent from the popmap shift the labels
    pops = rng.choice(["uv", "bv", "other"], size=n_samples, p=[0.45, 0.4, 0.15])
    popmap = "".join(f"{n}\t{p}\n" for n, p, m in zip(names, pops, in_map) if m) + "ghost\tuv\n"
    lines = ["##fileformat=VCFv4.2\n", "##contig=<ID=c2>\n", "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(names) + "\n"]
    calls = ["0|0", "0|1", "1|0", "1|1", "0/0", "0/1", "1/1", "./.", ".|.", ".", "0", "1", "./1", "0|.", "0|1|1", "0/2", "2|1", "1/1/1/0", ""]
    pc = np.array([30, 14, 14, 10, 4, 4, 4, 4, 2, 2, 2, 2, 2, 2, 1, 1, 1, 0.5, 0.5])
    pc = pc / pc.sum()
    for _ in range(n_rec):
        chrom = str(rng.choice(["c2", "c10", "c1", "scaf_9"]))
        pos = int(rng.integers(1, 400))                          # collisions -> duplicate keys, the last record wins
        filt = str(rng.choice(["PASS", ".", "q10", "LowQual"], p=[0.7, 0.15, 0.1, 0.05]))
        ref = str(rng.choice(["A", "c", "G", "t", "N", "AT"], p=[0.3, 0.25, 0.2, 0.15, 0.05, 0.05]))
        alt = str(rng.choice(["C", "g", "T", "a", "*", "CA", "A,C"], p=[0.3, 0.25, 0.15, 0.15, 0.05, 0.05, 0.05]))
        info = str(rng.choice(["DP=10", "ANN=T|missense_variant|MODERATE", "X|syn", "|", "A|B|C|D"]))
        fmt = str(rng.choice(["GT", "GT:DP", "DP:GT", "AD:DP:GT:GQ"], p=[0.5, 0.2, 0.2, 0.1]))
        gti = fmt.split(":").index("GT")
        ncol = n_samples if rng.random() < 0.85 else int(rng.integers(1, n_samples + 3))   # zip() truncates either way
        cols = []
        for _j in range(ncol):
            sub = [str(int(rng.integers(0, 40))) for _ in fmt.split(":")]
            sub[gti] = str(rng.choice(calls, p=pc))
            cols.append(":".join(sub))
        lines.append("\t".join([chrom, str(pos), ".", ref, alt, ".", filt, info, fmt] + cols) + "\n")
    text = "".join(lines)
    if rng.random() < 0.5:
        text = text[:-1]                                         # last line without its newline
    return text, popmap


@pytest.mark.parametrize("seed", range(8))
@pytest.mark.parametrize("container", ["gzip", "bgzf"])
def test_packer_matches_oracle_ingest(tmp_path, seed, container):
    from tdsfs_pack import pack_vcf
    rng = np.random.default_rng(500 + seed)
    n_samples = int(rng.choice([3, 40, 70, 130]))
    text, popmap = random_vcf(rng, int(rng.choice([60, 400, 1200])), n_samples)
    vcf, pm = str(tmp_path / "r.vcf.gz"), str(tmp_path / "popmap.txt")
    open(pm, "w").write(popmap)
    if container == "gzip":
        with gzip.open(vcf, "wb") as f:
            f.write(text.encode())
    else:
        write_bgzf(text.encode(), vcf, block=int(rng.choice([300, 4096, 65280])))
        assert gzip.open(vcf, "rb").read() == text.encode()
    d = O.make_data_dict_vcf(vcf, pm)
    # counts mode: make_data_dict_vcf itself (all populations, key text, REF/ALT, annotation), dict and orders identical
    from tdsfs_pack import vcf_to_data_dict
    dd = vcf_to_data_dict(vcf, pm, nthreads=int(rng.choice([1, 3, 8])))
    assert dd == d and list(dd) == list(d)
    assert all(list(dd[k]["calls"]) == list(d[k]["calls"]) for k in d)
    assert all(type(v) is int for k in list(d)[:50] for c in dd[k]["calls"].values() for v in c)
    P = pack_vcf(vcf, pm, "uv", "bv", nthreads=int(rng.choice([1, 2, 5])))
    keys = [f"{P.chroms[c]}-{p}" for c in range(len(P.chroms)) for p in P.pos[P.off[c]:P.off[c + 1]].tolist()]
    assert keys == sorted(d, key=lambda k: (k.split("-")[0], int(k.split("-")[1])))
    cnt = P.counts()
    for i, k in enumerate(keys):
        u, b = d[k]["calls"].get("uv", (0, 0)), d[k]["calls"].get("bv", (0, 0))
        assert cnt[i].tolist() == [u[0], u[1], b[0], b[1]], (k, cnt[i], u, b)
        assert P.ann[i] == d[k]["annotation"], k
    assert P.n_records - P.n_skipped >= P.n                     # duplicates collapse


def test_counts_mode_second_header_and_ragged_records(tmp_path):
    """A header line after records extends the positional population list (:78-85); short records leave later populations
    out of their calls dict (:118)."""
    from tdsfs_pack import vcf_to_data_dict
    pm = str(tmp_path / "popmap.txt")
    open(pm, "w").write("a\tuv\nb\tbv\nc\tzz\nd\tuv\n")
    head = "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t"
    text = (head + "a\tb\n" + "c1\t5\t.\tA\tC\t.\tPASS\tx|syn\tGT\t0|1\t1|1\n" + "c1\t9\t.\tg\tt\t.\t.\t.\tGT\t0|1\n" +
            head + "c\td\n" + "c1\t07\t.\tA\tC\t.\tPASS\t.\tGT:DP\t0|1:3\t1|1:2\t0/0:1\t./1:9\n" +
            "c1\t5\t.\tA\tG\t.\tPASS\t.\tGT\t1|1\t0|0\t0|1")
    vcf = str(tmp_path / "h.vcf.gz")
    with gzip.open(vcf, "wt") as f:
        f.write(text)
    d = O.make_data_dict_vcf(vcf, pm)
    dd = vcf_to_data_dict(vcf, pm)
    assert dd == d and list(dd) == list(d) and [list(v["calls"]) for v in dd.values()] == [list(v["calls"]) for v in d.values()]
    assert dd["c1-07"]["calls"] == {"uv": (1, 2), "bv": (0, 2), "zz": (2, 0)} and dd["c1-5"]["segregating"] == ("A", "G")


def test_packer_errors(tmp_path):
    from tdsfs_pack import pack_vcf, vcf_to_data_dict
    pm = str(tmp_path / "popmap.txt")
    open(pm, "w").write("a\tuv\nb\tbv\n")
    head = "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\ta\tb\n"
    for body, exc in (("c\t5\t.\tA\tC\t.\tPASS\t.\tDP\t3\t4\n", ValueError),            # no GT in FORMAT (:115)
                      ("c\t5\t.\tA\tC\t.\tPASS\t.\tGT\n", ValueError),                   # no sample columns: FORMAT is 'GT\\n' (:88, :115)
                      ("c\t5\t.\tA\tC\t.\tPASS\t.\tDP:GT\t3:0|1\t4\n", IndexError)):    # sample without the GT sub-field (:123)
        vcf = str(tmp_path / "e.vcf.gz")
        with gzip.open(vcf, "wt") as f:
            f.write(head + body)
        with pytest.raises(exc):
            O.make_data_dict_vcf(vcf, pm)
        with pytest.raises(exc):
            pack_vcf(vcf, pm, "uv", "bv")
        with pytest.raises(exc):
            vcf_to_data_dict(vcf, pm)
    vcf = str(tmp_path / "short.vcf.gz")
    with gzip.open(vcf, "wt") as f:
        f.write(head + "c\t5\t.\tA\tC\t.\tPASS\n")                                   # cols[7] (:92)
    with pytest.raises(IndexError):
        O.make_data_dict_vcf(vcf, pm)
    with pytest.raises(IndexError):
        vcf_to_data_dict(vcf, pm)
    with pytest.raises(FileNotFoundError):
        vcf_to_data_dict(str(tmp_path / "missing.vcf.gz"), pm)
    with pytest.raises(FileNotFoundError):
        pack_vcf(str(tmp_path / "missing.vcf.gz"), pm, "uv", "bv")
