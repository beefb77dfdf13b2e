"""Parity at the scale bench.py runs: 1 M x 150 bp reads at 30x (131 M k-mers) into Countgraph k=20, primes(4, 1e8), bigcount on,
default chunking (one 150 M-position chunk per call, uploaded in parts: the same grouped path and bucket loads as the bench's chunks),
compared with the UNMODIFIED reference compiled in oracle/_ref at one thread: the four table images, n_occupied,
n_unique_kmers, and — after a second, 300x batch that drives counters through 255 on every chunk — the bigcount map the
reference writes into its .ct file.  -m gpu; needs oracle/_ref (shipped prebuilt to the GPU box)."""
import os
import struct

import numpy as np
import pytest

import oracle_lib as ol
from common import md5

pytestmark = pytest.mark.gpu

K = 20
READ_LEN = 150


def _batch(seed, n_reads, genome_len):
    rng = np.random.default_rng(seed)
    g = rng.integers(0, 4, genome_len, dtype=np.uint8)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    buf = np.empty(n_reads * READ_LEN, dtype=np.uint8)
    view = buf.reshape(n_reads, READ_LEN)
    step = 1 << 18
    for i0 in range(0, n_reads, step):
        m = min(step, n_reads - i0)
        starts = rng.integers(0, genome_len - READ_LEN, m)
        strands = rng.integers(0, 2, m).astype(bool)
        codes = g[starts[:, None] + np.arange(READ_LEN)[None, :]]
        codes[strands] = (3 - codes[strands])[:, ::-1]
        view[i0:i0 + m] = lut[codes]
    off = np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(READ_LEN)
    return buf, off


def _write_fasta(path, buf, n_reads):
    rows = buf.reshape(n_reads, READ_LEN)
    block = np.empty((n_reads, READ_LEN + 4), dtype=np.uint8)
    block[:, :3] = np.frombuffer(b">r\n", dtype=np.uint8)
    block[:, 3:3 + READ_LEN] = rows
    block[:, -1] = ord("\n")
    with open(path, "wb") as fh:
        fh.write(block.tobytes())


def _ct_bigcounts(path, sizes):
    """the (hash, count) trailer of a .ct file (doc/dev/binary-file-formats.rst; src/oxli/storage.cc:582-638)"""
    with open(path, "rb") as fh:
        fh.seek(4 + 1 + 1 + 1 + 4 + 1 + 8 + sum(8 + s for s in sizes))
        n, = struct.unpack("<Q", fh.read(8))
        raw = np.frombuffer(fh.read(n * 10), dtype=np.dtype([("h", "<u8"), ("c", "<u2")]))
    return dict(zip(raw["h"].tolist(), raw["c"].tolist()))


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref (compiled reference) not present")
def test_bench_scale_parity_grouped_path(tmp_path):
    """the same comparison with the fused grouping kernel (k_part + k_apply2) instead of bins[] + k_bucketize + k_apply"""
    import subprocess
    import sys
    env = dict(os.environ, KMGPU_PREFER_BINS="0")
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", os.path.abspath(__file__), "-k",
                        "test_bench_scale_parity_with_compiled_reference"], env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref (compiled reference) not present")
def test_bench_scale_parity_with_compiled_reference(tmp_path):
    from khmer_b200 import cabi
    n_reads = 1_000_000
    sizes = ol.primes_near_x(4, 1e8)
    g = cabi.Sketch(cabi.BYTE, cabi.TWOBIT, K, sizes)
    g.set_use_bigcount(True)
    r = ol.Ref("Countgraph", K, sizes)
    r.set_use_bigcount(True)

    def both(seed, genome_len, tag):
        buf, off = _batch(seed, n_reads, genome_len)
        fa = str(tmp_path / ("reads_%s.fa" % tag))
        _write_fasta(fa, buf, n_reads)
        assert g.consume_reads((buf, off), clean=True) == n_reads * (READ_LEN - K + 1)
        reads, kmers = r.consume_seqfile(fa, threads=1)
        assert (reads, kmers) == (n_reads, n_reads * (READ_LEN - K + 1))
        os.unlink(fa)
        assert g.stats() == (r.n_occupied(), r.n_unique_kmers()), tag
        for i in range(4):
            assert md5(g.table(i)) == md5(r.table(i)), "%s: table %d differs from the reference" % (tag, i)

    # 1. the bench's workload: 30x coverage of a 5 Mbp genome
    both(4242, n_reads * READ_LEN // 30, "30x")
    assert g.bigcounts()[0].size == 0
    # 2. 300x coverage of a 0.5 Mbp genome: every 20-mer of it is seen ~260 times, counters cross 255 inside both chunks
    both(77, n_reads * READ_LEN // 300, "300x")
    ct = str(tmp_path / "ref.ct")
    r.save(ct)
    want = _ct_bigcounts(ct, sizes)
    gk, gv = g.bigcounts()
    assert len(want) > 1000
    assert dict(zip(gk.tolist(), gv.tolist())) == want
    g.close()
