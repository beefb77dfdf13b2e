"""Host layer pieces that need no GPU: hash helpers, prime finder, the FASTA/FASTQ reader — against the
reference's known answers and the parser goldens recorded from the compiled reference."""
import hashlib
import os

import pytest

import khmer_b200 as kh


def test_hash_known_answers():
    # tests/test_functions.py:51-99,148-171
    assert kh.forward_hash("AAAA", 4) == 0 and kh.forward_hash("TTTT", 4) == 0
    assert kh.forward_hash("CCCC", 4) == 170 and kh.forward_hash("GGGG", 4) == 170
    assert [kh.forward_hash_no_rc(s, 4) for s in ("AAAA", "TTTT", "CCCC", "GGGG")] == [0, 85, 170, 255]
    assert [kh.reverse_hash(h, 4) for h in (0, 85, 170, 255)] == ["AAAA", "TTTT", "CCCC", "GGGG"]
    assert kh.forward_hash("GGTTGACGGGGCTCAGGGGGCGGCTGACTCCG", 32) == 13607885392109549066
    with pytest.raises(ValueError):
        kh.forward_hash("A" * 33, 33)
    assert kh.hash_murmur3("AAAA") == 526240128537019279 == kh.hash_murmur3("TTTT")
    assert kh.hash_murmur3("CCCC") == 14391997331386449225 == kh.hash_murmur3("GGGG")
    assert kh.hash_no_rc_murmur3("AAAA") == 5231866503566620412
    assert kh.hash_no_rc_murmur3("TTTT") == 5753003579327329651
    assert kh.hash_no_rc_murmur3("CCCC") == 3789793362494378039
    assert kh.hash_no_rc_murmur3("GGGG") == 17519752047064575358
    assert kh.reverse_complement("AAAC") == "GTTT"


def test_hashes_against_reference_golden(golden):
    for s, c, f, r in golden["hash_twobit"]:
        assert kh.forward_hash(s, len(s)) == c and kh.forward_hash_no_rc(s, len(s)) == f
    for s, c, f, r in golden["hash_murmur"]:
        assert kh.hash_murmur3(s) == c and kh.hash_no_rc_murmur3(s) == f


def test_primes(golden):
    for key, want in golden["primes"].items():
        n, x = map(int, key.split(","))
        assert kh.get_n_primes_near_x(n, x) == want


def test_band_interval():
    assert kh.compute_band_interval(4, 0) == (0, (2 ** 64 - 1) // 4)
    with pytest.raises(ValueError):
        kh.compute_band_interval(4, 5)


def test_reader_matches_reference_parser(golden, datadir):
    for fn, want in golden["parse"].items():
        p = os.path.join(datadir, fn)
        if "error" in want:
            with pytest.raises(OSError):
                list(kh.ReadParser(p))
            continue
        parser = kh.ReadParser(p)
        seqs = [r.cleaned_seq for r in parser]
        assert len(seqs) == want["n"] == parser.num_reads, fn
        assert hashlib.md5("\n".join(seqs).encode()).hexdigest() == want["md5"], fn


def test_reader_errors(tmp_path, datadir):
    with pytest.raises(OSError):
        kh.ReadParser(str(tmp_path / "nope.fa"))
    with pytest.raises(OSError) as e:
        kh.ReadParser(os.path.join(datadir, "test-empty.fa"))
    assert "does not contain any sequences" in str(e.value)
    bad = tmp_path / "bad.fa"
    bad.write_text("this is not\nfasta\n")
    with pytest.raises(OSError):
        kh.ReadParser(str(bad))
    empty_seq = tmp_path / "e.fa"
    empty_seq.write_text(">a\nACGT\n>b\n>c\nAC\n")
    p = kh.ReadParser(str(empty_seq))
    assert next(p).sequence == "ACGT"
    with pytest.raises(ValueError):          # InvalidRead: "Sequence is empty" (read_parsers.cc:347-349)
        next(p)
    fq = tmp_path / "q.fq"
    fq.write_text("@a\nACGT\n+\nIII\n")
    with pytest.raises(ValueError):          # "Sequence and quality lengths differ"
        list(kh.ReadParser(str(fq)))
    multi = tmp_path / "m.fa"
    multi.write_text(">a desc\nACGT\nTTGA\n\n>b\r\nAC\r\nGT")
    assert [r.sequence for r in kh.ReadParser(str(multi))] == ["ACGTTTGA", "ACGT"]


def test_api_surface():
    """Every method of khmer/_oxli/graphs.pyx the four scripts and the table tests use exists."""
    import khmer_b200._oxli as ox
    want = ["ksize", "hash", "reverse_hash", "add", "count", "get", "consume", "get_kmers", "get_kmer_hashes",
            "get_kmer_counts", "get_min_count", "get_max_count", "get_median_count", "median_at_least",
            "n_unique_kmers", "n_occupied", "n_tables", "hashsizes", "set_use_bigcount", "get_use_bigcount", "save",
            "load", "consume_seqfile", "consume_seqfile_with_mask", "consume_seqfile_banding",
            "consume_seqfile_banding_with_mask", "abundance_distribution", "trim_on_abundance",
            "trim_below_abundance", "find_spectral_error_positions", "get_raw_tables"]
    for cls in ("Countgraph", "SmallCountgraph", "Nodegraph", "Counttable", "SmallCounttable", "Nodetable"):
        for m in want:
            assert hasattr(getattr(ox, cls), m), (cls, m)
    assert hasattr(ox.Nodegraph, "update")


def test_parallel_batch_parser_equals_serial(tmp_path, datadir):
    """read_batch on plain files cuts the mapping at record boundaries and parses slices concurrently; the result
    must be the serial parser's, record for record (subprocesses: thread count and threshold are read once)."""
    import subprocess
    import sys
    import numpy as np
    rng = np.random.default_rng(3)
    fa = tmp_path / "multi.fa"
    with open(fa, "w") as fh:
        for i in range(3000):
            n = int(rng.integers(1, 400))
            s = "".join("ACGTNacgt"[j] for j in rng.integers(0, 9, n))
            fh.write(">r%d some description\n" % i)
            for o in range(0, n, 70):                       # multi-line records, some blank lines, a CRLF here and there
                fh.write(s[o:o + 70] + ("\r\n" if i % 97 == 0 else "\n"))
            if i % 50 == 0:
                fh.write("\n")
    fq = tmp_path / "reads.fq"
    with open(fq, "w") as fh:
        for i in range(4000):
            n = int(rng.integers(1, 200))
            s = "".join("ACGTN"[j] for j in rng.integers(0, 5, n))
            q = "".join(chr(int(c)) for c in rng.integers(33, 74, n))
            if i % 7 == 0:
                q = "@" + q[1:]                              # quality lines starting with '@' must not fool the cutter
            fh.write("@read%d\n%s\n+\n%s\n" % (i, s, q))
    code = ("import sys; sys.path.insert(0, %r); import khmer_b200 as kh, hashlib\n"
            "p = kh.ReadParser(sys.argv[1]); out = []\n"
            "while True:\n"
            "    b = p.read_batch(int(sys.argv[2]))\n"
            "    if not b: break\n"
            "    out += b\n"
            "print(len(out), p.num_reads, hashlib.md5(b'\\n'.join(out)).hexdigest())\n") % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for path in (str(fa), str(fq), os.path.join(datadir, "synth-2k-150.fa"), os.path.join(datadir, "random-20-a.fq")):
        outs = set()
        for threads, batch in (("1", "100000000"), ("4", "100000000"), ("7", "20000"), ("3", "5000")):
            env = dict(os.environ, KMGPU_PARSE_THREADS=threads, KMGPU_PARSE_MIN_BYTES="1000")
            r = subprocess.run([sys.executable, "-c", code, path, batch], env=env, capture_output=True, text=True)
            assert r.returncode == 0, r.stderr[-2000:]
            outs.add(r.stdout.strip())
        assert len(outs) == 1, (path, outs)
        want = [r.sequence for r in kh.ReadParser(path)]
        n, nr, digest = next(iter(outs)).split()
        assert int(n) == len(want) == int(nr)
        assert digest == hashlib.md5("\n".join(want).encode()).hexdigest()


def test_parallel_batch_parser_falls_back_on_irregular_input(tmp_path):
    import subprocess
    import sys
    bad = tmp_path / "bad.fa"
    with open(bad, "w") as fh:
        for i in range(2000):
            fh.write(">r%d\nACGTACGTAC\n" % i)
        fh.write(">empty\n>after\nACGT\n")
    code = ("import sys; sys.path.insert(0, %r); import khmer_b200 as kh\n"
            "p = kh.ReadParser(sys.argv[1]); n = 0\n"
            "try:\n"
            "    while True:\n"
            "        b = p.read_batch(1000000)\n"
            "        if not b: break\n"
            "        n += len(b)\n"
            "except ValueError as e:\n"
            "    print('ValueError', n, e)\n") % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, KMGPU_PARSE_THREADS="4", KMGPU_PARSE_MIN_BYTES="1000")
    r = subprocess.run([sys.executable, "-c", code, str(bad)], env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip() == "ValueError 2000 Sequence is empty"      # the good reads first, then the reference's error


def _write_bgzf(path, data, block=30000):
    """block-compressed gzip as bgzip writes it: members of <= 64 KB with a 'BC' extra subfield holding the member size, then
    the empty end-of-file member"""
    import struct
    import zlib
    with open(path, "wb") as fh:
        for o in list(range(0, len(data), block)) + [None]:
            chunk = b"" if o is None else data[o:o + block]
            c = zlib.compressobj(6, zlib.DEFLATED, -15)
            body = c.compress(chunk) + c.flush()
            bsize = 12 + 6 + len(body) + 8
            fh.write(b"\x1f\x8b\x08\x04" + b"\0\0\0\0" + b"\0\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize - 1))
            fh.write(body + struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))


def test_compressed_input_inflated_ahead_and_parsed_in_parallel(tmp_path):
    """gzip / bzip2 input is inflated by a thread of its own and the inflated window parsed like a plain file; BGZF members are
    inflated concurrently.  Records must equal the plain file's, whatever the batch size; a truncated stream must raise after
    the reads before the damage."""
    import bz2
    import gzip
    import subprocess
    import sys
    import numpy as np
    rng = np.random.default_rng(11)
    recs = []
    for i in range(30000):
        n = int(rng.integers(30, 200))
        s = "".join("ACGTN"[j] for j in rng.integers(0, 5, n))
        q = "".join(chr(int(c)) for c in rng.integers(33, 74, n))
        recs.append("@read%d\n%s\n+\n%s\n" % (i, s, q))
    data = "".join(recs).encode()
    plain = tmp_path / "r.fq"
    plain.write_bytes(data)
    (tmp_path / "r.fq.gz").write_bytes(gzip.compress(data, 4))
    (tmp_path / "r2.fq.gz").write_bytes(gzip.compress(data[:len(data) // 2], 4) + gzip.compress(data[len(data) // 2:], 4))   # two members
    (tmp_path / "r.fq.bz2").write_bytes(bz2.compress(data))
    _write_bgzf(tmp_path / "r.bgzf.fq.gz", data)
    assert gzip.decompress((tmp_path / "r.bgzf.fq.gz").read_bytes()) == data    # the hand-built file is valid gzip
    want = [r.sequence for r in kh.ReadParser(str(plain))]
    code = ("import sys; sys.path.insert(0, %r); import khmer_b200 as kh, hashlib\n"
            "out = []\n"
            "try:\n"
            "    p = kh.ReadParser(sys.argv[1])\n"
            "    while True:\n"
            "        b = p.read_batch(int(sys.argv[2]))\n"
            "        if not b: break\n"
            "        out += b\n"
            "    print(len(out), hashlib.md5(b'\\n'.join(out)).hexdigest())\n"
            "except (OSError, ValueError) as e:\n"
            "    print('error', len(out), type(e).__name__)\n") % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    digest = hashlib.md5("\n".join(want).encode()).hexdigest()
    for name in ("r.fq.gz", "r2.fq.gz", "r.fq.bz2", "r.bgzf.fq.gz"):
        for threads, batch in (("1", "100000000"), ("4", "100000000"), ("5", "300000"), ("3", "20000")):
            env = dict(os.environ, KMGPU_PARSE_THREADS=threads, KMGPU_PARSE_MIN_BYTES="1000")
            r = subprocess.run([sys.executable, "-c", code, str(tmp_path / name), batch], env=env, capture_output=True, text=True)
            assert r.returncode == 0, r.stderr[-2000:]
            assert r.stdout.split() == [str(len(want)), digest], (name, threads, batch, r.stdout)
        assert [r.sequence for r in kh.ReadParser(str(tmp_path / name))] == want     # one read at a time
    # damage: a gzip stream cut short, a BGZF member with a flipped byte
    gz = (tmp_path / "r.fq.gz").read_bytes()
    (tmp_path / "cut.fq.gz").write_bytes(gz[:len(gz) * 2 // 3])
    bg = bytearray((tmp_path / "r.bgzf.fq.gz").read_bytes())
    bg[len(bg) // 2] ^= 0x5A
    (tmp_path / "bad.bgzf.fq.gz").write_bytes(bytes(bg))
    for name in ("cut.fq.gz", "bad.bgzf.fq.gz"):
        env = dict(os.environ, KMGPU_PARSE_THREADS="4", KMGPU_PARSE_MIN_BYTES="1000")
        r = subprocess.run([sys.executable, "-c", code, str(tmp_path / name), "200000"], env=env, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        tag, n, exc = r.stdout.split()
        # (the damaged member lies in the first inflated block of the BGZF file: nothing of it is handed out)
        assert tag == "error" and int(n) < len(want) and (int(n) > 0 or name.startswith("bad")) and exc in ("OSError", "ValueError"), r.stdout


def test_packed_batches_equal_cleaned_twobit(tmp_path):
    """read_batch in pack mode (what the device feed uploads): 2 bits per base, first base in the top bits of each 64-bit word,
    A0 T1 C2 G3, every other byte cleaned to A — with and without the SIMD packer, for several thread counts and batch sizes"""
    import subprocess
    import sys
    import numpy as np
    rng = np.random.default_rng(5)
    alphabet = np.frombuffer(b"ACGTacgtNnRY.-*", dtype=np.uint8)
    seqs = []
    fa = tmp_path / "mix.fa"
    with open(fa, "wb") as fh:
        for i in range(6000):
            n = int(rng.integers(1, 400))
            s = alphabet[rng.integers(0, len(alphabet) if i % 3 == 0 else 4, n)].tobytes()
            seqs.append(s)
            fh.write(b">r%d\n" % i)
            for o in range(0, n, 97):
                fh.write(s[o:o + 97] + b"\n")
    code_of = np.zeros(256, dtype=np.uint64)
    for ch, v in ((b"T", 1), (b"t", 1), (b"C", 2), (b"c", 2), (b"G", 3), (b"g", 3)):
        code_of[ch[0]] = v
    allb = np.frombuffer(b"".join(seqs), dtype=np.uint8)
    codes = code_of[allb]
    pad = (-len(codes)) % 32
    codes = np.concatenate([codes, np.zeros(pad, dtype=np.uint64)]).reshape(-1, 32)
    want_words = (codes << (np.uint64(62) - np.uint64(2) * np.arange(32, dtype=np.uint64))).sum(axis=1, dtype=np.uint64)
    want_off = np.concatenate([[0], np.cumsum([len(s) for s in seqs])]).astype(np.uint64)
    code = ("import sys; sys.path.insert(0, %r); import khmer_b200 as kh, numpy as np, hashlib\n"
            "p = kh.ReadParser(sys.argv[1]); words = []; offs = [np.zeros(1, dtype=np.uint64)]; bases = 0\n"
            "while True:\n"
            "    n, nb, w, o = p.read_batch_packed(int(sys.argv[2]))\n"
            "    if not n: break\n"
            "    w = np.frombuffer(w, dtype=np.uint64); o = np.frombuffer(o, dtype=np.uint64)\n"
            "    # batches are independent streams: unpack to codes and re-concatenate\n"
            "    c = ((w[:, None] >> (np.uint64(62) - np.uint64(2) * np.arange(32, dtype=np.uint64))) & np.uint64(3)).reshape(-1)[:nb]\n"
            "    words.append(c.astype(np.uint8)); offs.append(o[1:] + np.uint64(bases)); bases += nb\n"
            "allc = np.concatenate(words)\n"
            "print(bases, hashlib.md5(allc.tobytes()).hexdigest(), hashlib.md5(np.concatenate(offs).tobytes()).hexdigest())\n"
            ) % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    want_codes = code_of[allb].astype(np.uint8)
    want = "%d %s %s" % (len(allb), hashlib.md5(want_codes.tobytes()).hexdigest(), hashlib.md5(want_off.tobytes()).hexdigest())
    assert int(want_words[0]) >> 62 == int(code_of[allb[0]])
    for simd in ("0", "1"):
        for threads, batch in (("1", "100000000"), ("4", "100000000"), ("5", "70000"), ("3", "9000")):
            env = dict(os.environ, KMGPU_PARSE_THREADS=threads, KMGPU_PARSE_MIN_BYTES="1000", KMGPU_NO_SIMD=simd)
            r = subprocess.run([sys.executable, "-c", code, str(fa), batch], env=env, capture_output=True, text=True)
            assert r.returncode == 0, r.stderr[-2000:]
            assert r.stdout.strip() == want, (simd, threads, batch)
