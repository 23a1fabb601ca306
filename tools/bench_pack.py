#!/usr/bin/env python
"""K0 (host VCF packer, csrc/vcf_pack.cpp) throughput on a synthetic gzip VCF: records x samples genotype calls per second,
beside the dict-building ingest of the drop-in API (the reference's make_data_dict_vcf algorithm) on a slice of the same file.
usage: python tools/bench_pack.py [records] [samples_per_pop] [threads]"""
import gzip
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "2dsfs-scan_b200"))


def write_vcf(path, popmap, R, ns, seed=1, missing=0.02):
    rng = np.random.default_rng(seed)
    names = [f"A{i}" for i in range(ns)] + [f"B{i}" for i in range(ns)]
    with open(popmap, "w") as f:
        for n in names:
            f.write(f"{n}\t{'p1' if n[0] == 'A' else 'p2'}\n")
    gts = np.array(["0|0", "0|1", "1|0", "1|1", ".|."])
    with gzip.open(path, "wt", compresslevel=1) as f:
        f.write("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(names) + "\n")
        pos = np.cumsum(rng.geometric(1 / 50, size=R))
        for r in range(R):
            p = np.exp(rng.uniform(np.log(1e-3), 0))
            a = (rng.random(2 * ns) < p).astype(int) + 2 * (rng.random(2 * ns) < p).astype(int)
            a[rng.random(2 * ns) < missing] = 4
            f.write(f"chr{1 + r * 4 // R}\t{pos[r]}\t.\tA\tC\t.\tPASS\tAC=1|syn|x\tGT\t" + "\t".join(gts[a]) + "\n")


def write_bgzf(data, path, block=65280):
    import struct
    import zlib
    with open(path, "wb") as f:
        for o in list(range(0, len(data), block)) + [None]:
            chunk = b"" if o is None else data[o:o + block]
            c = zlib.compressobj(6, zlib.DEFLATED, -15)
            comp = c.compress(chunk) + c.flush()
            f.write(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, 12 + 6 + len(comp) + 8 - 1))
            f.write(comp + struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))


def main():
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    ns = int(sys.argv[2]) if len(sys.argv) > 2 else 500
    nt = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    vcf, popmap = f"/tmp/vcfbench/s{R}_{ns}.vcf.gz", f"/tmp/vcfbench/s{R}_{ns}.popmap"
    os.makedirs("/tmp/vcfbench", exist_ok=True)
    if not os.path.exists(vcf):
        write_vcf(vcf, popmap, R, ns)
    import tdsfs_pack
    bgz = vcf.replace(".vcf.gz", ".bgzf.vcf.gz")
    if not os.path.exists(bgz):
        write_bgzf(gzip.open(vcf, "rb").read(), bgz)
    text_mb = sum(len(b) for b in iter(lambda f=gzip.open(vcf, "rb"): f.read(1 << 24), b"")) / 1e6
    for name, path in (("gzip", vcf), ("bgzf", bgz)):
        best = 1e9
        for _ in range(5):  # the first calls of a process pay for cold malloc arenas / page faults
            t0 = time.perf_counter()
            P = tdsfs_pack.pack_vcf(path, popmap, "p1", "p2", nt)
            best = min(best, time.perf_counter() - t0)
        calls = P.n * 2 * ns
        print(f"packer ({name}): {P.n} SNPs x {2 * ns} samples in {best:.3f} s = {calls / best / 1e6:.1f} M calls/s, "
              f"{text_mb / best:.0f} MB/s of VCF text, threads={nt or len(os.sched_getaffinity(0))}")
    # the dict-building ingest (reference algorithm) on the first records
    import twoDSFS_class as K
    small = "/tmp/vcfbench/small.vcf.gz"
    n_small = min(R, 1000)
    with gzip.open(vcf, "rt") as f, gzip.open(small, "wt") as g:
        for i, line in enumerate(f):
            if i >= n_small + 2:
                break
            g.write(line)
    t0 = time.perf_counter()
    d = K.parse_vcf_to_dict(small, popmap)
    dt2 = time.perf_counter() - t0
    print(f"dict ingest (reference algorithm, Python): {len(d)} SNPs in {dt2:.3f} s = {len(d) * 2 * ns / dt2 / 1e6:.2f} M calls/s")


if __name__ == "__main__":
    main()
