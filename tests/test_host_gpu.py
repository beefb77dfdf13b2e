"""The liboxli-compatible host layer end to end on the GPU, written like the reference's own table tests
(tests/test_countgraph.py, test_nodegraph.py, test_tabletype.py, test_counttable.py): Python class ->
C++ Hashtable/GpuStorage -> C ABI -> kernels; outputs compared with goldens from the compiled reference."""
import hashlib
import os
import threading

import numpy as np
import pytest

import oracle_lib as ol
from common import reads_of, md5

pytestmark = pytest.mark.gpu


def _kh():
    import khmer_b200
    return khmer_b200


def _case(golden, name):
    return next(c for c in golden["cases"] if c["name"] == name)


@pytest.mark.parametrize("name", ["r20-cg-1e5", "r20-ng-1e4", "r20-scg-1e4", "r20-ct-k20", "r20-nt-k33", "r20-sct-k40",
                                  "ab2-cg-big", "ab2-cg-tiny-big", "syn-cg-k20", "syn-ct-k40", "synerr-cg-k20",
                                  "ragged-cg-k15", "lowc-cg-k20-big", "lowc-cg-k8-big", "lowc-ct-k33-big", "C1-25k-cg",
                                  "25k-ng-k32", "25k-scg-k31", "25k-ct-k40"])
def test_consume_seqfile_save_matches_reference(golden, datadir, tmp_path, name):
    """consume_seqfile(filename) then save(): counters, statistics and the saved file are the reference's,
    byte for byte (incl. the bigcount trailer, whose order is the std::unordered_map's)."""
    kh = _kh()
    rec = _case(golden, name)
    t = getattr(kh, rec["cls"])(rec["k"], 1, 1, primes=rec["sizes"])
    if rec["bigcount"] is not None:
        t.set_use_bigcount(rec["bigcount"])
    reads, kmers = t.consume_seqfile(os.path.join(datadir, rec["file"]))
    assert (reads, kmers) == (rec["reads"], rec["kmers"])
    assert t.n_unique_kmers() == rec["n_unique"] and t.n_occupied() == rec["n_occupied"]
    assert t.hashsizes() == rec["sizes"] and t.n_tables() == len(rec["sizes"]) and t.ksize() == rec["k"]
    raw = t.get_raw_tables()
    assert [md5(bytes(v)) for v in raw] == rec["table_md5"]
    p = str(tmp_path / "out.ct")
    t.save(p)
    assert os.path.getsize(p) == rec["file_size"]
    assert md5(open(p, "rb").read()) == rec["file_md5"]
    # load round trip (classmethod load, graphs.pyx:303-307)
    t2 = getattr(kh, rec["cls"]).load(p)
    assert t2.ksize() == rec["k"] and t2.hashsizes() == rec["sizes"] and t2.n_occupied() == rec["n_occupied"]
    assert [md5(bytes(v)) for v in t2.get_raw_tables()] == rec["table_md5"]
    if rec["bigcount"] is not None:
        assert t2.get_use_bigcount() == rec["bigcount"]
    p2 = str(tmp_path / "again.ct")
    t2.save(p2)
    a, b = open(p, "rb").read(), open(p2, "rb").read()
    if rec["cls"] in ("Countgraph", "Counttable"):
        # the bigcount trailer is written in unordered_map iteration order, which a load (re-insertion in file
        # order) legitimately permutes — in the reference too; compare it as a set
        body = 4 + 2 + 1 + 4 + 1 + 8 + sum(8 + s for s in rec["sizes"])
        assert a[:body + 8] == b[:body + 8]
        ents = lambda d: sorted(d[body + 8 + 10 * i: body + 18 + 10 * i] for i in range((len(d) - body - 8) // 10))
        assert ents(a) == ents(b)
        if ol.have_ref():
            r = ol.Ref(rec["cls"], 1, [1])
            r.load(p)
            p3 = str(tmp_path / "ref_again.ct")
            r.save(p3)
            assert open(p3, "rb").read() == b      # same permutation as the reference's own load -> save
    else:
        assert a == b
    # queries through the loaded table
    seqs = [ol.clean(s) for s in reads_of(rec["file"])[:len(rec["medians"])]]
    for s, m in zip(seqs, rec["medians"]):
        if m is None:
            with pytest.raises(ValueError):
                t2.get_median_count(s)
            continue
        med, avg, sd = t2.get_median_count(s)
        assert med == m[0]
        assert np.float32(avg).tobytes() == bytes.fromhex(m[1]) and np.float32(sd).tobytes() == bytes.fromhex(m[2])
        assert t2.median_at_least(s, 2) == bool(m[3]) and t2.median_at_least(s, 5) == bool(m[4])
    if "counts0" in rec:
        assert t2.get_kmer_counts(seqs[0]) == rec["counts0"]
        assert t2.get_kmer_hashes(seqs[0]) == rec["hashes0"]


@pytest.mark.parametrize("name", ["r20-cg-k12", "ab2-cg-big", "ab2-cg-nobig", "syn-scg-k31", "lowc-cg-k20-big", "25k-ct-k40"])
def test_abundance_distribution_matches_reference(golden, datadir, name):
    kh = _kh()
    rec = _case(golden, name)
    path = os.path.join(datadir, rec["file"])
    t = getattr(kh, rec["cls"])(rec["k"], 1, 1, primes=rec["sizes"])
    if rec["bigcount"] is not None:
        t.set_use_bigcount(rec["bigcount"])
    t.consume_seqfile(path)
    tracking = (kh.Nodegraph if rec["cls"].endswith("graph") else kh.Nodetable)(rec["k"], 1, 1, primes=rec["sizes"])
    dist = t.abundance_distribution(path, tracking)
    assert len(dist) == 65535                      # graphs.pyx:293-296
    assert {str(i): v for i, v in enumerate(dist) if v} == rec["abund"]


def test_gz_save_load(golden, datadir, tmp_path):
    kh = _kh()
    rec = _case(golden, "ab2-cg-big")
    t = kh.Countgraph(rec["k"], 1, 1, primes=rec["sizes"])
    t.set_use_bigcount(True)
    t.consume_seqfile(os.path.join(datadir, rec["file"]))
    p = str(tmp_path / "t.ct.gz")
    t.save(p)
    import gzip
    plain = str(tmp_path / "t.ct")
    t.save(plain)
    assert gzip.open(p, "rb").read() == open(plain, "rb").read()
    t2 = kh.Countgraph.load(p)
    assert t2.get(t2.hash("GGTTGACGGGGCTCAGG")) == t.get(t.hash("GGTTGACGGGGCTCAGG")) == 1001


def test_reference_fixture_files(datadir, tmp_path):
    """goodversion-k12.ht(.gz) load; badversion files and truncation at every byte raise OSError
    (tests/test_countgraph.py:695-714,1118-1158, tests/test_nodegraph.py)."""
    kh = _kh()
    ng = kh.Nodegraph.load(os.path.join(datadir, "goodversion-k12.ht"))
    assert ng.ksize() == 12
    with pytest.raises(OSError):
        kh.Countgraph.load(os.path.join(datadir, "goodversion-k12.ht"))      # wrong type
    for fn, cls in (("badversion-k12.ct", kh.Countgraph), ("badversion-k12.ht", kh.Nodegraph)):
        with pytest.raises(OSError) as e:
            cls.load(os.path.join(datadir, fn))
        assert "Does not start with signature for a oxli file" in str(e.value)   # pre-OXLI files, as the reference reports them
    with pytest.raises(OSError):
        kh.Countgraph.load(str(tmp_path / "missing.ct"))
    t = kh.Countgraph(5, 1, 1, primes=[11, 13])
    t.set_use_bigcount(True)
    for _ in range(300):
        t.count("AAAAC")
    p = str(tmp_path / "small.ct")
    t.save(p)
    data = open(p, "rb").read()
    assert kh.Countgraph.load(p).get("AAAAC") == 300
    for cut in range(len(data)):
        q = str(tmp_path / "trunc.ct")
        with open(q, "wb") as fh:
            fh.write(data[:cut])
        with pytest.raises(OSError):
            kh.Countgraph.load(q)
    for cls, ext in ((kh.SmallCountgraph, "sct"), (kh.Nodegraph, "pt")):
        s = cls(5, 1, 1, primes=[11, 13])
        s.count("AAAAC")
        p = str(tmp_path / ("small." + ext))
        s.save(p)
        data = open(p, "rb").read()
        assert cls.load(p).get("AAAAC") == 1
        for cut in range(len(data)):
            q = str(tmp_path / "trunc.bin")
            with open(q, "wb") as fh:
                fh.write(data[:cut])
            with pytest.raises(OSError):
                cls.load(q)


def test_single_kmer_api_like_reference_tests():
    kh = _kh()
    # tests/test_countgraph.py:123-146 collision suite with explicit sizes
    GG = "G" * 12
    t = kh.Countgraph(12, 1, 1, primes=[1000003, 1009837])
    assert t.hash(GG) == 11184810 and t.hash("AAACGTATGACT") == 184777
    # hash(GG) % 1000003 == hash(collision_1), hash(GG) % 1009837 == hash(collision_2): test_collision_1/2/3
    o = ol.Oracle("Countgraph", 12, [1000003, 1009837])
    for kmer in ("AAACGTATGACT", "AAATACCGAGCG", "AAACGTATCGAG"):
        t.count(kmer)
        o.add(kmer)
    assert t.get("AAACGTATGACT") == 1 and t.get(GG) == o.get(GG) == 1      # both of GG's bins are taken
    assert t.add(GG) is o.add(GG) is False and t.add("ACGTACGTTTTT") is True
    assert t.get(t.hash(GG)) == 2 and t.reverse_hash(t.hash(GG)) == "C" * 12
    with pytest.raises(ValueError):
        t.get("ACGT")                      # wrong length
    with pytest.raises(TypeError):
        t.get(3.5)
    with pytest.raises(ValueError):
        t.consume("ACGT")                  # shorter than k
    assert t.consume("A" * 20) == 9
    # bigcount (tests/test_countgraph.py:890-1036)
    b = kh.Countgraph(4, 4 ** 4, 4)
    b.set_use_bigcount(True)
    for _ in range(1000):
        b.count("GGTT")
    assert b.get("GGTT") == 1000
    nb = kh.Countgraph(4, 4 ** 4, 4)
    for _ in range(500):
        nb.count("GGTT")
    assert nb.get("GGTT") == 255
    with pytest.raises(ValueError):
        kh.Nodegraph(4, 100, 2).set_use_bigcount(True)
    # nibble clamp (tests/test_nibblestorage.py:69-95)
    s = kh.SmallCounttable(4, 1, 1, primes=[5, 7])
    for _ in range(20):
        s.add(3)
    assert s.get(3) == 15 and s.get(4) == 0 and s.get(2) == 0
    assert bytes(s.get_raw_tables()[0]) == bytes([0, 0x0F, 0])
    # Murmur tables (tests/test_counttable.py:42-60)
    c = kh.Counttable(4, 1e5, 3)
    assert c.hash("AAAC") == 11898086063751343884 and c.hash("AAAG") == 10548630838975263317
    with pytest.raises(ValueError):
        c.reverse_hash(12345)
    assert c.get_kmer_hashes("AAACG") == [c.hash("AAAC"), c.hash("AACG")]
    # median (tests/test_countgraph.py:285-330)
    m = kh.Countgraph(4, 1e5, 2)
    m.consume("AAAAAA")
    assert m.get_median_count("AAAAAA") == (3, 3.0, 0.0)
    assert m.median_at_least("AAAAAA", 3) and not m.median_at_least("AAAAAA", 4)
    with pytest.raises(ValueError):
        m.get_median_count("AAA")
    assert m.get_min_count("AAAAAC") == 0 and m.get_max_count("AAAAAC") == 3
    assert m.get_kmers("AAAAC") == ["AAAA", "AAAC"]


def test_occupancy_known_answers(datadir):
    # tests/test_nodegraph.py:222-260, tests/test_countgraph.py:633-666
    kh = _kh()
    path = os.path.join(datadir, "random-20-a.fa")
    for cls in (kh.Nodegraph, kh.Countgraph):
        t = cls(20, 100000, 3)
        t.consume_seqfile(path)
        assert (t.n_occupied(), t.n_unique_kmers()) == (3884, 3960)
        t = cls(20, 10000, 3)
        t.consume_seqfile(kh.ReadParser(path))
        assert (t.n_occupied(), t.n_unique_kmers()) == (3269, 3916)
    for ext in (".gz", ".bz2"):
        t = kh.Nodegraph(20, 100000, 3)
        assert t.consume_seqfile(path + ext) == (99, 3960)
        assert t.n_occupied() == 3884


def test_nodegraph_update(datadir):
    kh = _kh()
    a, b = kh.Nodegraph(20, 1e4, 3), kh.Nodegraph(20, 1e4, 3)
    a.consume_seqfile(os.path.join(datadir, "random-20-a.fa"))
    b.consume_seqfile(os.path.join(datadir, "synth-2k-150.fa"))
    oa, ob = ol.Oracle("Nodegraph", 20, a.hashsizes()), ol.Oracle("Nodegraph", 20, a.hashsizes())
    oa.consume_reads(reads_of("random-20-a.fa"))
    ob.consume_reads(reads_of("synth-2k-150.fa"))
    a.update(b)
    oa.update(ob)
    assert a.n_occupied() == oa.n_occupied()
    assert [bytes(v) for v in a.get_raw_tables()] == [oa.table(i).tobytes() for i in range(3)]
    with pytest.raises(ValueError):
        a.update(kh.Nodegraph(20, 1e5, 3))


def test_threads_share_one_parser(golden, datadir):
    """scripts/load-into-counting.py:145-158: T threads call consume_seqfile on one table with one parser.
    Counter bytes do not depend on the interleaving."""
    kh = _kh()
    rec = _case(golden, "C1-25k-cg")
    t = kh.Countgraph(rec["k"], 1, 1, primes=rec["sizes"])
    os.environ["KMGPU_FEED_BASES"] = "200000"
    try:
        parser = kh.ReadParser(os.path.join(datadir, rec["file"]))
        out = []
        ths = [threading.Thread(target=lambda: out.append(t.consume_seqfile(parser))) for _ in range(4)]
        [th.start() for th in ths]
        [th.join() for th in ths]
    finally:
        del os.environ["KMGPU_FEED_BASES"]
    assert sum(o[0] for o in out) == rec["reads"] and sum(o[1] for o in out) == rec["kmers"]
    assert [md5(bytes(v)) for v in t.get_raw_tables()] == rec["table_md5"]
    assert t.n_occupied() == rec["n_occupied"]


def test_banding_and_mask_against_live_reference(datadir, tmp_path):
    if not ol.have_ref():
        pytest.skip("compiled reference not shipped")
    kh = _kh()
    path = os.path.join(datadir, "synth-err-n.fa")
    sizes = ol.primes_near_x(3, 40000)
    for band in range(3):
        g = kh.Countgraph(21, 1, 1, primes=sizes)
        r = ol.Ref("Countgraph", 21, sizes)
        assert g.consume_seqfile_banding(path, 3, band) == r.consume_seqfile_banding(path, 3, band)
        assert [bytes(v) for v in g.get_raw_tables()] == [r.table(i).tobytes() for i in range(3)]
        assert g.n_unique_kmers() == r.n_unique_kmers()
    gm = kh.Nodegraph(21, 1, 1, primes=sizes)
    rm = ol.Ref("Nodegraph", 21, sizes)
    gm.consume_seqfile(os.path.join(datadir, "random-20-a.fa"))
    rm.consume_seqfile(os.path.join(datadir, "random-20-a.fa"))
    mixed = str(tmp_path / "mixed.fa")
    with open(mixed, "w") as fh:
        fh.write(open(os.path.join(datadir, "random-20-a.fa")).read())
        fh.write(open(path).read())
    for consume_masked in (False, True):
        g = kh.Countgraph(21, 1, 1, primes=sizes)
        r = ol.Ref("Countgraph", 21, sizes)
        assert g.consume_seqfile_with_mask(mixed, gm, 0, consume_masked) == r.consume_seqfile_with_mask(mixed, rm, 0, consume_masked)
        assert [bytes(v) for v in g.get_raw_tables()] == [r.table(i).tobytes() for i in range(3)]


def _trim_expected(counts, k, seqlen, keep):
    """Hashtable::trim_on_abundance / trim_below_abundance (src/oxli/hashtable.cc:504-560) over a count vector."""
    if len(counts) == 0:
        return 0
    if len(counts) == 1 or not keep(counts[0]):
        return 0
    i = k
    for c in counts[1:]:
        if not keep(c):
            return i
        i += 1
    return seqlen


def test_trim_functions_against_oracle_counts():
    kh = _kh()
    rng = np.random.default_rng(17)
    g = kh.Countgraph(8, 1e5, 3)
    o = ol.Oracle("Countgraph", 8, g.hashsizes())
    for s in ("ACGTACGTAAGGTTCCACGTACGTAAGGTTCC", "ACGTACGTAAGGTTCC", "GGATTACAGGATTACATTT"):
        g.consume(s)
        o.consume(s)
    queries = ["ACGTACGTAAGGTTCCTTTTTTTTTTTTACGTACGT", "TTTTTTTTTTTTACGTACGTAAGG", "GGATTACAGGATTACATTTACGTACGTAAGGTTCC",
               "ACGTACGT", "ACGTACGTA"] + ["".join("ACGT"[i] for i in rng.integers(0, 4, 40)) for _ in range(5)]
    for q in queries:
        c = o.kmer_counts(q).tolist()
        for thr in (1, 2, 3):
            pos = _trim_expected(c, 8, len(q), lambda v: v >= thr)
            assert g.trim_on_abundance(q, thr) == (q[:pos], pos)
            pos = _trim_expected(c, 8, len(q), lambda v: v <= thr)
            assert g.trim_below_abundance(q, thr) == (q[:pos], pos)
    seq = "ACGTACGTAAGGTTCCTTTTTTTTTTTTACGTACGT"
    assert g.find_spectral_error_positions(seq, 1) == [16]      # hashtable.cc:565-612 walked by hand


@pytest.mark.parametrize("cls,fn,k", [("Nodegraph", "random-20-a.fa", 20), ("Countgraph", "synth-err-n.fa", 21),
                                      ("SmallCountgraph", "lowcomplexity.fa", 12), ("Nodegraph", "25k.fq.gz", 20)])
def test_consume_seqfile_and_tag_against_live_reference(datadir, tmp_path, cls, fn, k):
    """load-graph.py's default path (oxli/functions.py:57-66 -> Hashgraph::consume_seqfile_and_tag, src/oxli/hashgraph.cc:200-320):
    counters and is-new bits from the device, the tag scan on the host; tag count, n_consumed (= new k-mers) and the saved
    tagset are the compiled reference's, byte for byte.  (Inputs without reads shorter than k: for those the reference inserts an
    uninitialised hash value as a tag, hashgraph.cc:207,263-266 — undefined behaviour this layer does not reproduce.)"""
    if not ol.have_ref() or not hasattr(ol.ref_lib(), "ref_n_tags"):
        pytest.skip("compiled reference with the tagging wrappers not shipped")
    kh = _kh()
    sizes = ol.primes_near_x(4, 2e5 if fn != "25k.fq.gz" else 4e6)
    path = os.path.join(datadir, fn)
    g = getattr(kh, cls)(k, 1, 1, primes=sizes)
    r = ol.Ref(cls, k, sizes)
    assert g.consume_seqfile_and_tag(path) == r.consume_seqfile_and_tag(path)
    assert g.n_tags() == r.n_tags() and g.n_tags() > 0
    assert g.n_unique_kmers() == r.n_unique_kmers() and g.n_occupied() == r.n_occupied()
    a, b = str(tmp_path / "a.tagset"), str(tmp_path / "b.tagset")
    g.save_tagset(a)
    r.save_tagset(b)
    assert open(a, "rb").read() == open(b, "rb").read()
    g2 = getattr(kh, cls)(k, 1, 1, primes=sizes)
    g2.load_tagset(a)
    assert g2.n_tags() == g.n_tags() and g2._get_tag_density() == 40
    assert [bytes(v) for v in g.get_raw_tables()] == [r.table(i).tobytes() for i in range(4)]


def test_normalize_batch_reference_md5(datadir):
    """scripts/normalize-by-median.py -k 21 -C 20 -M 1e7 on simple-genome-reads.fa (tests/test_script_output.py:51-59) through
    the host layer's batch entry: the kept records written as the script writes them have the reference's md5."""
    kh = _kh()
    recs, name = [], None
    for ln in open(os.path.join(datadir, "simple-genome-reads.fa")):
        ln = ln.rstrip("\n")
        if ln.startswith(">"):
            name = ln[1:]
        else:
            recs.append((name, ln))
    g = kh.Countgraph(21, 1e7 / 4, 4)
    keep, kmers = g.normalize_batch([s.upper().replace("N", "A") for _, s in recs], 20)
    out = "".join(">%s\n%s\n" % (n, s) for (n, s), k in zip(recs, keep) if k)
    assert hashlib.md5(out.encode()).hexdigest() == "942e9024c25a8d85033d755d86aba4a3"
    assert kmers == sum(len(s) - 20 for (_, s), k in zip(recs, keep) if k)


@pytest.mark.parametrize("cutoff,variable,z,want", [(2, False, 20, "9495801b282ff6b08961b685d12a954c"),
                                                    (4, False, 20, "65596253b87ed8d5aeb14dc8cf5a7406"),
                                                    (4, True, 15, "393805ac92e8bed31a374de9ee89ead8")])
def test_trim_low_abund_reference_md5(datadir, cutoff, variable, z, want):
    """scripts/trim-low-abund.py -k 21 -M 1e7 -C cutoff [-V -Z z] on simple-genome-reads.fa (tests/test_script_output.py:118-182):
    the script's two passes (scripts/trim-low-abund.py:196-271, khmer/trimming.py:36-66) over this backend's get_median_count,
    median_at_least, trim_on_abundance and consume write the records with the reference's md5 (the oracle reproduces all six
    pinned md5s in tests/test_oracle.py)."""
    kh = _kh()
    recs, name = [], None
    for ln in open(os.path.join(datadir, "simple-genome-reads.fa")).read().split("\n"):
        if ln.startswith(">"):
            name = ln[1:]
        elif ln:
            recs.append((name, ln))
    g = kh.Countgraph(21, 1e7 / 4, 4)
    k = 21

    def trim_record(name, seq, cleaned):
        if variable and not g.median_at_least(cleaned, z):
            return (name, seq)
        _, t = g.trim_on_abundance(cleaned, cutoff)
        if t < k:
            return None
        return (name, seq if t == len(seq) else seq[:t])

    out, saved = [], []
    for name, seq in recs:
        cleaned = seq.upper().replace("N", "A")
        if g.get_median_count(cleaned)[0] >= z:
            r = trim_record(name, seq, cleaned)
            if r:
                out.append(r)
        else:
            g.consume(cleaned)
            saved.append((name, seq))
    for name, seq in saved:
        cleaned = seq.upper().replace("N", "A")
        if not variable or g.median_at_least(cleaned, z):
            r = trim_record(name, seq, cleaned)
            if r:
                out.append(r)
        else:
            out.append((name, seq))
    text = "".join(">%s\n%s\n" % r for r in out)
    assert hashlib.md5(text.encode()).hexdigest() == want


def test_ascii_feed_still_matches(golden, datadir):
    """the host feed packs reads to 2 bits on the parser threads by default; KMGPU_FEED_PACKED=0 ships ASCII and packs on the device"""
    import subprocess, sys
    env = dict(os.environ, KMGPU_FEED_PACKED="0")
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", os.path.abspath(__file__), "-k",
                        "test_consume_seqfile_save_matches_reference and (C1 or syn-ct or ragged)"], env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
