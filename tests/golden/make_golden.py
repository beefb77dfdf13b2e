#!/usr/bin/env python
"""Generate tests/golden/golden.json by running the UNMODIFIED reference (oracle/_ref, built by
`make -C oracle ref` from /root/reference) at one thread.  Run in the build container only; the output is
committed so that neither the CPU tests nor the GPU box need /root/reference.

    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import struct
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402

DATA = os.path.join(HERE, "data")


def md5(b):
    return hashlib.md5(bytes(b)).hexdigest()


def f32hex(x):
    return struct.pack("<f", x).hex()


def synth_reads(seed, n_reads, read_len, genome_len, err=0.0, with_n=False):
    """SURVEY.md §8d synthetic reads: uniform random genome, uniform start, random strand."""
    rng = np.random.default_rng(seed)
    g = rng.integers(0, 4, genome_len, dtype=np.uint8)
    comp = np.array([3, 2, 1, 0], dtype=np.uint8)  # ACGT -> TGCA
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    starts = rng.integers(0, genome_len - read_len, n_reads)
    strands = rng.integers(0, 2, n_reads)
    out = []
    for s, st in zip(starts, strands):
        r = g[s:s + read_len]
        if st:
            r = comp[r[::-1]]
        r = lut[r].copy()
        if err:
            m = rng.random(read_len) < err
            r[m] = lut[rng.integers(0, 4, int(m.sum()))]
        if with_n and rng.random() < 0.05:
            r[rng.integers(0, read_len)] = ord("N")
        out.append(r.tobytes().decode())
    return out


def write_fasta(path, reads):
    with open(path, "w") as fh:
        for i, r in enumerate(reads):
            fh.write(">r%d\n%s\n" % (i, r))


def run_case(name, cls, k, sizes, path, bigcount=None, n_median=40, abund=True, save=True):
    print("  case", name, file=sys.stderr)
    ref = ol.Ref(cls, k, sizes)
    if bigcount is not None:
        ref.set_use_bigcount(bigcount)
    reads, kmers = ref.consume_seqfile(path)
    rec = {
        "name": name, "cls": cls, "k": k, "sizes": [int(s) for s in sizes], "file": os.path.basename(path),
        "bigcount": bigcount, "reads": reads, "kmers": kmers, "n_unique": ref.n_unique_kmers(),
        "n_occupied": ref.n_occupied(), "table_md5": [md5(ref.table(i)) for i in range(len(sizes))],
    }
    if save:
        with tempfile.TemporaryDirectory() as td:
            p = os.path.join(td, "t.ct")
            ref.save(p)
            rec["file_md5"] = md5(open(p, "rb").read())
            rec["file_size"] = os.path.getsize(p)
    seqs = ol.ref_parse_clean(path)
    med = []
    for s in seqs[:n_median]:
        if len(s) < k:
            med.append(None)
            continue
        m, a, d = ref.median(s)
        med.append([m, f32hex(a), f32hex(d), int(ref.median_at_least(s, 2)), int(ref.median_at_least(s, 5))])
    rec["medians"] = med
    if seqs and len(seqs[0]) >= k:
        rec["counts0"] = ref.kmer_counts(seqs[0]).tolist()
        rec["hashes0"] = [int(x) for x in ref.kmer_hashes(seqs[0])]
    if abund:
        track_cls = "Nodegraph" if cls.endswith("graph") else "Nodetable"
        tracking = ol.Ref(track_cls, k, sizes)
        dist = ref.abundance_distribution(path, tracking)
        rec["abund"] = {str(i): int(v) for i, v in enumerate(dist) if v}
        rec["tracking_unique"] = tracking.n_unique_kmers()
    return rec


def main():
    L = ol.ref_lib()
    out = {"primes": {}, "hash_twobit": [], "hash_murmur": [], "cases": [], "parse": {}}
    # primes
    for n, x in [(4, 100000000), (2, 1000), (3, 100000), (3, 10000), (4, 1e5), (1, 1), (4, 2), (3, 7), (2, 3),
                 (4, 1000000000), (4, 2000000000), (2, 10000000), (1, 4 ** 4 + 1), (4, 8e9)]:
        p = (ol.C.c_uint64 * 8)()
        got = L.ref_primes(n, int(x), p)
        out["primes"]["%d,%d" % (n, int(x))] = [int(p[i]) for i in range(got)]
    # hashes on random k-mers
    rng = np.random.default_rng(7)
    for k in [1, 2, 4, 5, 12, 15, 16, 17, 20, 21, 31, 32]:
        for _ in range(6):
            s = "".join("ACGT"[i] for i in rng.integers(0, 4, k))
            c, f, r = ol.C.c_uint64(), ol.C.c_uint64(), ol.C.c_uint64()
            L.ref_hash_twobit(s.encode(), k, ol.C.byref(c), ol.C.byref(f), ol.C.byref(r))
            out["hash_twobit"].append([s, c.value, f.value, r.value])
    pal = ["ACGT", "AATT", "ACGCGT", "GAATTC", "ATATATATATATATATATAT", "ACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT"]
    for k in [1, 4, 7, 8, 9, 15, 16, 17, 20, 31, 32, 33, 40, 47, 48, 49, 64, 65, 100, 127]:
        for _ in range(4):
            s = "".join("ACGT"[i] for i in rng.integers(0, 4, k))
            c, f, r = ol.C.c_uint64(), ol.C.c_uint64(), ol.C.c_uint64()
            L.ref_hash_murmur(s.encode(), k, ol.C.byref(c), ol.C.byref(f), ol.C.byref(r))
            out["hash_murmur"].append([s, c.value, f.value, r.value])
    for s in pal:
        c, f, r = ol.C.c_uint64(), ol.C.c_uint64(), ol.C.c_uint64()
        L.ref_hash_murmur(s.encode(), len(s), ol.C.byref(c), ol.C.byref(f), ol.C.byref(r))
        out["hash_murmur"].append([s, c.value, f.value, r.value])

    def primes(n, x):
        return out["primes"].setdefault("%d,%d" % (n, int(x)), ol.primes_near_x(n, int(x)))

    # synthetic inputs (regenerated identically by the tests from the same seeds)
    syn = os.path.join(DATA, "synth-2k-150.fa")
    write_fasta(syn, synth_reads(42, 2000, 150, 10000))
    syn_err = os.path.join(DATA, "synth-err-n.fa")
    write_fasta(syn_err, synth_reads(43, 1500, 100, 20000, err=0.01, with_n=True))
    ragged = os.path.join(DATA, "ragged.fa")
    rng2 = np.random.default_rng(5)
    rr = []
    for i in range(400):
        ln = int(rng2.integers(0, 90))
        rr.append("".join("ACGTNacgtn"[j] for j in rng2.integers(0, 10, ln)) if ln else "A")
    write_fasta(ragged, rr)
    lowc = os.path.join(DATA, "lowcomplexity.fa")
    lc = ["A" * 120, "A" * 80, "AC" * 60, "ACG" * 40, "T" * 150] * 120 + synth_reads(44, 300, 100, 3000)
    write_fasta(lowc, lc)

    r20 = os.path.join(DATA, "random-20-a.fa")
    ab2 = os.path.join(DATA, "test-abund-read-2.fa")
    c25 = os.path.join(DATA, "25k.fq.gz")
    cases = [
        ("r20-cg-1e5", "Countgraph", 20, primes(3, 1e5), r20, None),
        ("r20-cg-1e4", "Countgraph", 20, primes(3, 1e4), r20, None),
        ("r20-ng-1e5", "Nodegraph", 20, primes(3, 1e5), r20, None),
        ("r20-ng-1e4", "Nodegraph", 20, primes(3, 1e4), r20, None),
        ("r20-cg-k12", "Countgraph", 12, primes(4, 1e5), r20, None),
        ("r20-scg-1e4", "SmallCountgraph", 20, primes(3, 1e4), r20, None),
        ("r20-ct-k20", "Counttable", 20, primes(3, 1e4), r20, None),
        ("r20-nt-k33", "Nodetable", 33, primes(4, 1e4), r20, None),
        ("r20-sct-k40", "SmallCounttable", 40, primes(2, 1e4), r20, None),
        ("ab2-cg-big", "Countgraph", 17, primes(4, 1e5), ab2, True),
        ("ab2-cg-nobig", "Countgraph", 17, primes(4, 1e5), ab2, False),
        ("ab2-cg-tiny-big", "Countgraph", 12, [101, 103], ab2, True),
        ("ab2-scg", "SmallCountgraph", 17, primes(4, 1e5), ab2, None),
        ("syn-cg-k20", "Countgraph", 20, primes(4, 2e4), syn, True),
        ("syn-cg-k32", "Countgraph", 32, primes(4, 5e4), syn, True),
        ("syn-ng-k32", "Nodegraph", 32, primes(4, 1e5), syn, None),
        ("syn-scg-k31", "SmallCountgraph", 31, primes(4, 3e4), syn, None),
        ("syn-ct-k40", "Counttable", 40, primes(4, 3e4), syn, True),
        ("syn-nt-k21", "Nodetable", 21, primes(2, 3e4), syn, None),
        ("synerr-cg-k20", "Countgraph", 20, primes(4, 1e5), syn_err, True),
        ("synerr-sct-k25", "SmallCounttable", 25, primes(3, 5e4), syn_err, None),
        ("ragged-cg-k15", "Countgraph", 15, primes(3, 5e3), ragged, True),
        ("ragged-ng-k4", "Nodegraph", 4, primes(2, 200), ragged, None),
        ("ragged-ct-k17", "Counttable", 17, primes(3, 5e3), ragged, True),
        ("lowc-cg-k20-big", "Countgraph", 20, primes(4, 1e4), lowc, True),
        ("lowc-cg-k20-nobig", "Countgraph", 20, primes(4, 1e4), lowc, False),
        ("lowc-cg-k8-big", "Countgraph", 8, primes(3, 500), lowc, True),
        ("lowc-scg-k20", "SmallCountgraph", 20, primes(4, 1e4), lowc, None),
        ("lowc-ct-k33-big", "Counttable", 33, primes(2, 300), lowc, True),
        ("C1-25k-cg", "Countgraph", 20, primes(4, 1e8), c25, True),
        ("25k-ng-k32", "Nodegraph", 32, primes(4, 1e8), c25, None),
        ("25k-scg-k31", "SmallCountgraph", 31, primes(4, 1e8), c25, None),
        ("25k-ct-k40", "Counttable", 40, primes(4, 1e8), c25, True),
    ]
    for name, cls, k, sizes, path, big in cases:
        heavy = name.startswith(("C1", "25k"))
        out["cases"].append(run_case(name, cls, k, sizes, path, bigcount=big, n_median=10 if heavy else 40))

    # what the reference parser + cleaner yields for each fixture (md5 of '\n'.join(cleaned reads))
    for fn in sorted(os.listdir(DATA)):
        p = os.path.join(DATA, fn)
        if fn.endswith((".ht", ".ct", ".gz.ht")) or "version" in fn:
            continue
        try:
            seqs = ol.ref_parse_clean(p)
            out["parse"][fn] = {"n": len(seqs), "bases": sum(map(len, seqs)), "md5": md5("\n".join(seqs).encode())}
        except ol.RefError as e:
            out["parse"][fn] = {"error": str(e)}
    with open(os.path.join(HERE, "golden.json"), "w") as fh:
        json.dump(out, fh, indent=0, sort_keys=True)
    print("wrote golden.json: %d cases" % len(out["cases"]))


if __name__ == "__main__":
    main()
